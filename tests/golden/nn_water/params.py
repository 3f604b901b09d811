class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01', '02']
    nat = 25
    start_prod = 1
    end_prod = 3
    md_print_freq = 4
    sample_freq = 1
    celldm = 12.164004228700508
    timestep = 0.002
    parse_vel = False
