"""which_trajs / user_confirm with the reference's behaviour (helper.py:5-27): the dipolar
driver lists the trajectory directories and blocks on ``Continue? (Y/n)``."""
import os
import sys
import time


def user_confirm():
    start_input = time.time()
    while True:
        confirm = input('Continue? (Y/n)')   # EOFError propagates on a closed stdin, as in the reference
        if confirm.lower() in ('y', 'yes', ''):
            break
        elif confirm.lower() in ('n', 'no'):
            print("Stopping")
            sys.exit(2)
        else:
            print("Please enter yes(Y) or no(N)")
    return time.time() - start_input


def which_trajs(PD):
    if 'trajs' not in PD.__dict__.keys():
        print("List of trajectories not provided. All will be parsed...")
        PD.trajs = [t.name for t in os.scandir(PD.traj_dir) if t.name.isnumeric()]
    print("The following directories will be parsed as trajectories:")
    for traj in PD.trajs:
        print(PD.traj_dir + traj + '\n')
    return user_confirm()
