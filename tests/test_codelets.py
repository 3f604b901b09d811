"""The generated fp64 DFT codelets (dynpy_b200/csrc/codelets.cuh, tools/gen_codelets.py) compiled for the
host and checked against numpy.fft — runs without a GPU; the same source is what the kernels inline."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dynpy_b200", "csrc")

HOST_SRC = r"""
#include "codelets.cuh"
using namespace dynpy;
#define CASE(ID, NAME, N) case ID: { cd v[N]; for (int j = 0; j < N; ++j) v[j] = make_double2(io[2*j], io[2*j+1]); \
    NAME(v); for (int j = 0; j < N; ++j) { io[2*j] = v[j].x; io[2*j+1] = v[j].y; } return N; }
extern "C" int run_codelet(int id, double* io) {
    switch (id) {
        CASE(0, fft2, 2) CASE(1, fft4, 4) CASE(2, fft8, 8) CASE(3, fft16, 16)
        CASE(4, fft2_h1, 2) CASE(5, fft4_h2, 4) CASE(6, fft8_h4, 8) CASE(7, fft16_h8, 16)
        CASE(8, fft5, 5) CASE(9, fft25, 25) CASE(10, fft25_odd, 25)
        CASE(11, fft32, 32) CASE(12, fft8_odd, 8) CASE(13, fft16_odd, 16) CASE(14, fft32_odd, 32)
        CASE(15, fft10, 10) CASE(16, fft10_odd, 10) CASE(17, fft12, 12) CASE(18, fft12_odd, 12) CASE(19, fft20, 20) CASE(20, fft20_odd, 20)
    }
    return -1;
}
"""
# id, n, non-zero leading inputs, kind
SPECS = [(0, 2, 2, ""), (1, 4, 4, ""), (2, 8, 8, ""), (3, 16, 16, ""), (4, 2, 1, ""), (5, 4, 2, ""), (6, 8, 4, ""),
         (7, 16, 8, ""), (8, 5, 5, ""), (9, 25, 25, ""), (10, 25, 25, "odd"), (11, 32, 32, ""), (12, 8, 8, "odd"),
         (13, 16, 16, "odd"), (14, 32, 32, "odd"), (15, 10, 10, ""), (16, 10, 10, "odd"), (17, 12, 12, ""),
         (18, 12, 12, "odd"), (19, 20, 20, ""), (20, 20, 20, "odd")]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not installed")
    d = tmp_path_factory.mktemp("codelets")
    src = d / "host_codelets.cpp"
    src.write_text(HOST_SRC)
    so = d / "libcodelets.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-I", cuda_inc, "-I", CSRC, "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.run_codelet.argtypes = [ctypes.c_int, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("cid,n,nz,kind", SPECS)
def test_codelet_matches_numpy(lib, cid, n, nz, kind):
    rng = np.random.default_rng(100 + cid)
    for _ in range(5):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        x[nz:] = 0
        io = np.zeros(2 * n)
        io[0::2], io[1::2] = x.real, x.imag
        io[2 * nz:] = np.nan  # structurally zero inputs must not be read
        assert lib.run_codelet(cid, io.ctypes.data) == n
        got = io[0::2] + 1j * io[1::2]
        ref = np.fft.fft(np.concatenate([x, np.zeros(n)]))[1::2] if kind == "odd" else np.fft.fft(x)
        assert np.max(np.abs(got - ref)) <= 4e-16 * np.sqrt(n) * np.max(np.abs(ref))


def test_generated_header_is_current(tmp_path):
    """codelets.cuh is exactly what tools/gen_codelets.py emits (no hand edits)."""
    before = open(os.path.join(CSRC, "codelets.cuh")).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_codelets.py")], check=True, capture_output=True)
    assert open(os.path.join(CSRC, "codelets.cuh")).read() == before
