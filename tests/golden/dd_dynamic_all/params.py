class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 18
    start_prod = 2
    end_prod = 26
    md_print_freq = 2
    sample_freq = 2
    celldm = 11.05173128763446
    timestep = 0.001
    parse_vel = False
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'all'
    rmax = 12.0
