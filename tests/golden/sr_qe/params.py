class ParseDynamics:
    MD_format = 'QE'
    traj_dir = './'
    trajs = ['01']
    nat = 18
    start_prod = 2
    end_prod = 26
    md_print_freq = 7
    sample_freq = 2
    celldm = 11.05173128763446
    timestep = 0.001
    parse_vel = False
    symbols = ['O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H', 'O', 'H', 'H']
class SpinRotation:
    mol_type = 'water'
    nmol = 6
    C_SR = [36.9, 33.5, 35.5]
