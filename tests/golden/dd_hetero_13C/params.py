class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = 24
    start_prod = 1
    end_prod = 40
    md_print_freq = 2
    sample_freq = 1
    celldm = 22.0
    timestep = 0.002
    parse_vel = False
class Dipolar:
    identifier = 'H(3)C(2)N(1)'
    pair_type = 'intra'
    rmax = 8.0
    analyte_label = 1
    analyte_species = '13C'
