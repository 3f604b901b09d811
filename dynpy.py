"""``python -m dynpy -q|-d|-s <params.py>`` — drop-in entry point with the reference's CLI
(dynpy.py:42-124); the relaxation modes run on the B200 CUDA path in ``dynpy_b200``."""
from dynpy_b200.cli import main, read_input, usage  # noqa: F401

if __name__ == "__main__":
    main()
