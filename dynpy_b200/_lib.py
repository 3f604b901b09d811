"""ctypes binding of the C-ABI in include/dynpy_b200.h.

The shared library is built in-tree (dynpy_b200/csrc/libdynpy_b200.so) by
``__graft_entry__.build()`` / ``make -C dynpy_b200/csrc``.  There is NO fallback: if the
library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYNPY_B200_LIB") or os.path.join(_HERE, "csrc", "libdynpy_b200.so")

p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f64 = C.c_double

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    "dynpy_version": (i32, []),
    "dynpy_last_error": (C.c_char_p, []),
    "dynpy_device_sm_count": (i32, []),
    "dynpy_profile_enable": (i32, [i32]),
    "dynpy_profile_read": (i32, [p, p]),
    "dynpy_fft_length": (i64, [i64]),
    "dynpy_fft_workspace_bytes": (i64, [i64, i64]),
    "dynpy_wk_acf_sum_f64": (i32, [p, p, i64, i64, i64, i32, p, p, i64, p, i64, p]),
    "dynpy_ifft_acf_f64": (i32, [p, i32, i64, p, i64, f64, p, i64, p]),
    "dynpy_wk_each_workspace_bytes": (i64, [i64, i64]),
    "dynpy_wk_acf_each_f64": (i32, [p, p, i64, i64, i64, p, i64, p, i64, p]),
    "dynpy_qr_efg_to_spectrum_f64": (i32, [p, i64, i64, p, i64, p, i64, p]),
    "dynpy_wrap_f64": (i32, [p, p, i64, f64, p]),
    "dynpy_pdist_workspace_bytes": (i64, [i64]),
    "dynpy_pdist_ortho_f64": (i32, [p, p, p, i64, i64, f64, f64, f64, p, f64, p, p, p, p, p, p, p, p, p, i64, p]),
    "dynpy_cart_to_dipolar_f64": (i32, [p, p, p, p, i64, p, p]),
    "dynpy_supercell_f64": (i32, [p, i64, i64, i64, f64, p, p]),
    "dynpy_topology_check_f64": (i32, [p, i64, i64, p, f64, p, f64, f64, f64, f64, p, p]),
    "dynpy_topology_events_f64": (i32, [p, i64, i64, p, f64, p, f64, f64, f64, f64, p, i64, p, p]),
    "dynpy_topology_events_range_f64": (i32, [p, i64, i64, i64, i64, p, f64, p, f64, f64, f64, f64, p, i64, p, p]),
    "dynpy_dd_pair_count_f64": (i32, [p, i64, i64, p, p, i64, f64, f64, f64, f64, p, p, p]),
    "dynpy_dd_workspace_bytes": (i64, [i64, i64, i64]),
    "dynpy_dd_pairs_to_spectrum_f64": (i32, [p, i64, i64, p, p, i64, f64, f64, f64, p, i64, p, i64, p]),
    "dynpy_dd_ragged_workspace_bytes": (i64, [i64, i64, i64]),
    "dynpy_dd_pairs_ragged_to_acf_f64": (i32, [p, i64, i64, p, p, i64, f64, f64, f64, f64, p, i64, p, i64, p]),
    "dynpy_dd_pairs_masked_to_acf_f64": (i32, [p, i64, i64, p, p, i64, f64, f64, f64, f64, p, p, p, p, i64, p, i64, p]),
    "dynpy_dd_pairs_ragged_capped_f64": (i32, [p, i64, i64, p, p, i64, f64, f64, f64, f64, p, p, p, p, i64, i64, p, i64, p]),
    "dynpy_sr_angmom_f64": (i32, [p, p, i64, i64, p, i64, i32, p, p, i32, f64, f64, f64, f64, p, p, p, p]),
    "dynpy_simpson_f64": (i32, [p, p, i64, i32, i64, i32, p, p]),
    "dynpy_dft_workspace_bytes": (i64, [i64]),
    "dynpy_dft_f64": (i32, [p, p, i64, i64, p, i64, p, i64, p]),
    "dynpy_cutoff_bounds_f64": (i32, [p, i64, i32, i64, i32, f64, p, p]),
    "dynpy_parse_md_blocks_host": (i32, [C.c_char_p, i64, i32, i32, p, i64, f64, p, p, i32]),
    "dynpy_parse_md_blocks_hdr_host": (i32, [C.c_char_p, i64, i32, i32, p, i64, f64, p, p, p, i32]),
}


class DynpyError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DynpyError(
                "dynpy_b200 CUDA library not built: %s is missing. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C dynpy_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dynpy_last_error()
        raise DynpyError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
