class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 3
    end_prod = 42
    md_print_freq = 5
    sample_freq = 1
    celldm = 12.164004228700508
    timestep = 0.002
    parse_vel = False
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'inter'
    rmax = 9.0
