class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 30
    start_prod = 2
    end_prod = 35
    md_print_freq = 2
    sample_freq = 1
    celldm = 127.33630547834731
    timestep = 0.001
    parse_vel = False
class SpinRotation:
    mol_type = 'methane'
    nmol = 5
    C_SR = [16.495, 16.495, -1.875]
