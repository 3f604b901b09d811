"""dynpy_b200 — B200-native correlation-function hot path of DynPy.

Only what the hot path needs lives here:

``csrc/``      hand-written sm_100a CUDA kernels + the C-ABI (``include/dynpy_b200.h``)
``_lib``       ctypes binding of that C-ABI (fails loudly when the .so is missing)
``ops``        torch-tensor wrappers over the C-ABI (device pointers, streams)
``qrax`` / ``DDparse`` / ``SRparse`` / ``distances`` / ``ddrax`` / ``SRrax``
               host-side mirrors of the reference modules of the same names
``cli``        ``python -m dynpy`` front-end (getopt table of dynpy.py:42-124)
``synth``      seeded synthetic inputs of the BASELINE shapes
"""
__version__ = "0.1.0"
