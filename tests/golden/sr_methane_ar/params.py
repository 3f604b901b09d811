class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 21
    start_prod = 1
    end_prod = 24
    md_print_freq = 1
    sample_freq = 1
    celldm = 111.23850891757796
    timestep = 0.001
    parse_vel = False
class SpinRotation:
    mol_type = 'methane'
    nmol = 4
    C_SR = [16.495, 16.495, -1.875]
