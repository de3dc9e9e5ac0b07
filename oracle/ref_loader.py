"""TEST INFRASTRUCTURE ONLY.  Loads the compiled, unmodified reference (`oracle/_ref/redner*.so`, built by
oracle/build_ref.sh from /root/reference/src) as a Python module object *without* putting it on sys.path under the
name `redner` (the product ships its own module of that name).  Only tests/, bench.py's cpu_baseline / --impl
reference legs and __graft_entry__.smoke() may import this file."""
import ctypes
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_cached = None


def available() -> bool:
    return len(glob.glob(os.path.join(_REF_DIR, "redner*.so"))) > 0


def load():
    """Return the reference's pybind11 module (src/redner.cpp:20-272), CPU/Embree path."""
    global _cached
    if _cached is not None:
        return _cached
    cands = glob.glob(os.path.join(_REF_DIR, "redner*.so"))
    if not cands:
        raise ImportError("oracle/_ref not built; run `bash oracle/build_ref.sh` where /root/reference exists")
    # Embree depends on TBB; preload both so the loader finds them regardless of LD_LIBRARY_PATH.
    for lib in ("libtbbmalloc.so.2", "libtbb.so.2", "libembree3.so.3"):
        p = os.path.join(_REF_DIR, lib)
        if os.path.exists(p):
            ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    spec = importlib.util.spec_from_file_location("redner", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod
