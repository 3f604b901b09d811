class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 1
    end_prod = 30
    md_print_freq = 1
    sample_freq = 1
    celldm = 12.164004228700508
    timestep = 0.002
    parse_vel = False
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'all'
    rmax = 16.0
    analyte_label = 1
    analyte_species = '1H'
