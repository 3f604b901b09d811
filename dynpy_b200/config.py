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
