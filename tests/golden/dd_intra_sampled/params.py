class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01', '02']
    nat = 24
    start_prod = 2
    end_prod = 46
    md_print_freq = 1
    sample_freq = 2
    celldm = 12.164004228700508
    timestep = 0.004
    parse_vel = False
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'intra'
    rmax = 15
