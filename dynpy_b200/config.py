"""Run-time switches shared by the drivers."""
import os


def default_simpson_mode() -> str:
    """'cartwright' (scipy >= 1.11) or 'avg' (older scipy): behave like the scipy the reference
    would import in this environment; DYNPY_SIMPSON=avg|cartwright overrides."""
    env = os.environ.get("DYNPY_SIMPSON")
    if env in ("avg", "cartwright"):
        return env
    try:
        import scipy
        major, minor = (int(x) for x in scipy.__version__.split(".")[:2])
        return "cartwright" if (major, minor) >= (1, 11) else "avg"
    except Exception:
        return "cartwright"


# ---- optional per-stage wall clock of a driver run (DYNPY_STAGE_TIMES=<path.json>): device-synchronised
#      brackets around parse / H2D / topology / transforms / CSV, written by rank 0 when the driver returns
import contextlib
import json
import time

_STAGES = {}


@contextlib.contextmanager
def stage(name: str):
    path = os.environ.get("DYNPY_STAGE_TIMES")
    if not path:
        yield
        return
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        _STAGES[name] = _STAGES.get(name, 0.0) + time.perf_counter() - t0


def dump_stages(extra=None) -> None:
    path = os.environ.get("DYNPY_STAGE_TIMES")
    if not path:
        return
    rec = {"stages_s": dict(_STAGES)}
    if extra:
        rec.update(extra)
    with open(path, "w") as fh:
        json.dump(rec, fh)
