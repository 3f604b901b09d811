class ParseDynamics:
    MD_format = 'QE'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 1
    end_prod = 40
    md_print_freq = 7
    sample_freq = 1
    celldm = 12.164004228700508
    timestep = 0.001
    parse_vel = False
    symbols = ['O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H']
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'all'
    rmax = 9.0
