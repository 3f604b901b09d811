"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference (/root/reference) in this
container.  Nothing in the product path may import this file.

The reference targets the pandas of its era; under pandas 3 two behaviours changed
(SURVEY.md §8c, Appendix A).  This module restores them *before* the reference is
imported, then hands control to the reference's own ``dynpy.main()``:

1. ``future.infer_string = False`` (parseMD.py:246 assigns floats into str columns);
2. ``GroupBy.apply`` passes the grouping columns to the callee again (pandas-2
   ``include_groups=True`` semantics; isotopes.py:151, qrax.py:162 rely on it).

It is only usable where /root/reference exists (this container, not the GPU box);
its outputs are committed as fixtures under tests/golden/ by oracle/make_golden.py.
"""
from __future__ import annotations

import os
import subprocess
import sys

REFERENCE = os.environ.get("DYNPY_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "dynpy.py"))


def install_shim() -> None:
    import pandas as pd
    from functools import wraps
    from pandas.core.groupby.groupby import GroupBy
    from pandas.core.groupby.generic import DataFrameGroupBy

    pd.set_option("future.infer_string", False)

    def _apply(self, func, *args, include_groups=True, **kwargs):
        f = wraps(func)(lambda g: func(g, *args, **kwargs)) if (args or kwargs) else func
        if isinstance(self, DataFrameGroupBy) and self._selection is None:
            data = self._selected_obj
        else:
            data = self._obj_with_exclusions
        return self._python_apply_general(f, data)

    GroupBy.apply = _apply


def import_reference():
    """Make the reference's flat modules importable in *this* process (leaf numerics:
    qrax, ddrax, SRrax, distances ...)."""
    install_shim()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)


def run_cli(flag: str, params_py: str, cwd: str, stdin: str = "y\n", timeout: int = 3600) -> str:
    """``python -m dynpy <flag> <params_py>`` of the reference, in a subprocess whose
    cwd holds the parameter file (dynpy.py:17-24 imports it by bare name)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import refshim; "
            "refshim.install_shim(); sys.argv=['dynpy', %r, %r]; import dynpy; dynpy.main()"
            % (REFERENCE, os.path.dirname(os.path.abspath(__file__)), flag, params_py))
    env = dict(os.environ)
    env["PYTHONPATH"] = ""
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_ref")
    r = subprocess.run([sys.executable, "-c", code], cwd=cwd, input=stdin, text=True,
                       capture_output=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference run failed (%d)\n%s\n%s" % (r.returncode, r.stdout[-4000:], r.stderr[-4000:]))
    return r.stdout
