"""``python -m dynpy`` front-end with the reference's contract (dynpy.py:7-127): same getopt
table, same required-key checks, messages and exit codes; the relaxation modes run on the
CUDA path.  ``--DDrax`` / ``--SRrax`` (the spellings of BASELINE.json) are accepted as aliases."""
from __future__ import annotations

import getopt
import inspect
import sys


def usage():
    print("USAGE:" + "\n" +
          "-h   --help                                Displays this message" + "\n\n" +
          "-i   --inputs <params.py>                  Generate EFG calculation inputs for ADF and/or QE-PAW." + "\n\n" +
          "-e   --parse-efgs <params.py>              Parse EFG data from ADF or GIPAW outputs." + "\n\n" +
          "-q   --Qrelax <params.py>                  Compute quadrupolar relaxation rates and related parameters from EFG time series.\n\n" +
          "-s   --SRrelax <params.py>                 Calculate SR relaxation rates and related parameters from MD trajectory.\n" +
          "                                               Currently supported molecule types are 'water','acetonitrile', 'methane', and methyl groups\n\n" +
          "-d   --DDrelax <params.py>                 Compute dipolar relaxation rates and related parameters from MD trajectory.")


def read_input(input_arg, required):
    """dynpy.py:17-35"""
    try:
        file = input_arg.split('.py')[0].strip('./\\')
    except Exception:
        print("\nPlease provide, as an argument, an input file with required paramaters.\n")
        sys.exit(2)
    try:
        dynpy_params = __import__(file)
    except ImportError:
        print("\nInput parameter file must have .py file extension and have valid python syntax.\n")
        sys.exit(2)
    input_classes = inspect.getmembers(dynpy_params, inspect.isclass)
    all_input_vars = [varr for clas in input_classes for varr in list(clas[1].__dict__.keys()) if "__" not in varr]
    for clas, reqs in required.items():
        for req in reqs:
            if req not in all_input_vars:
                print("Error: Missing required input variable %s in class %s." % (req, clas))
                sys.exit(2)
    return dynpy_params


_PD_REQ = ['MD_format', 'traj_dir', 'timestep', 'celldm', 'trajs', 'md_print_freq', 'nat', 'start_prod', 'end_prod']


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if "" not in sys.path and "." not in sys.path:
        sys.path.insert(0, "")  # the parameter file is imported by bare name from the cwd
    try:
        opts, args = getopt.getopt(argv, "hi:e:q:s:d:", ["help", 'inputs=', 'parse-efgs=', 'Qrelax=', 'SRrelax=',
                                                         'DDrelax=', 'DDrax=', 'SRrax='])
    except getopt.GetoptError:
        usage()
        sys.exit(2)
    if opts == []:
        usage()
        sys.exit(2)
    for opt, arg in opts:
        if opt in ("-h", "--help"):
            usage()
        elif opt in ("-i", "--inputs", "-e", "--parse-efgs"):
            print("%s is outside the accelerated hot path (EFG input generation / output scraping): "
                  "run it with the reference DynPy." % opt)
            sys.exit(2)
        elif opt in ("-q", "--Qrelax"):
            dynpy_params = read_input(arg, {'Quadrupolar': ['data_set', 'analyte']})
            from . import dist, qrax
            dist.init_from_env()
            qrax.QR_module_main(dynpy_params.Quadrupolar)
        elif opt in ("-s", "--SRrelax", "--SRrax"):
            dynpy_params = read_input(arg, {'ParseDynamics': _PD_REQ, 'SpinRotation': ['mol_type', 'C_SR']})
            from . import SRparse, dist
            dist.init_from_env()
            SRparse.SR_module_main(dynpy_params.ParseDynamics, dynpy_params.SpinRotation)
        elif opt in ("-d", "--DDrelax", "--DDrax"):
            dynpy_params = read_input(arg, {'ParseDynamics': _PD_REQ, 'Dipolar': ['identifier']})
            from . import DDparse, dist
            dist.init_from_env()
            DDparse.DD_module_main(dynpy_params.ParseDynamics, dynpy_params.Dipolar)
    if 'dynpy_b200.dist' in sys.modules:
        sys.modules['dynpy_b200.dist'].shutdown()   # multi-rank runs: tear the process group down cleanly


if __name__ == "__main__":
    main()
