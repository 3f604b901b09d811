class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 1
    end_prod = 22
    md_print_freq = 1
    sample_freq = 1
    celldm = 22.0
    timestep = 0.001
    parse_vel = False
class SpinRotation:
    mol_type = 'methyl'
    identifier = 'H(3)C(2)N(1)'
    nmol = 4
    C_SR = [1.0, 1.0, 12.0]
    methyl_indeces = [0, 1, 2, 3, 4]
    global_momentum = True
