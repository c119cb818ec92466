"""Import the UNMODIFIED reference (``/root/reference/AmpliPy.py``) with the pysam shim.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (``/root/reference`` does not
exist on the GPU box); used by ``tests/golden/make_golden.py`` to generate committed golden
vectors and by CPU tests (skipped when the reference is absent) to pin the C restatement in
``oracle/amplipy_oracle.c`` against the reference's own functions.
"""
import importlib.util
import os
import sys

REFERENCE_PATH = os.environ.get("AMPLIPY_REFERENCE", "/root/reference/AmpliPy.py")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pysam_shim")
_cached = None


def reference_available():
    return os.path.isfile(REFERENCE_PATH)


def load_reference():
    """Return the reference module object (its functions are the golden oracle)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(REFERENCE_PATH)
    saved = sys.modules.get("pysam")
    sys.path.insert(0, _SHIM_DIR)
    try:
        sys.modules.pop("pysam", None)
        spec = importlib.util.spec_from_file_location("amplipy_reference", REFERENCE_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(_SHIM_DIR)
        if saved is not None:
            sys.modules["pysam"] = saved
    _cached = mod
    return mod


def shim():
    """The pysam shim module the reference was loaded with."""
    return load_reference().pysam
