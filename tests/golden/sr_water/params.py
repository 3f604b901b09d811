class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 1
    end_prod = 29
    md_print_freq = 1
    sample_freq = 2
    celldm = 12.164004228700508
    timestep = 0.001
    parse_vel = False
class SpinRotation:
    mol_type = 'water'
    nmol = 8
    C_SR = [36.9, 33.5, 35.5]
