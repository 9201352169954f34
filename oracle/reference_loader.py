"""TEST INFRASTRUCTURE ONLY.  Import the reference's own matching code, unmodified, from /root/reference.

Only usable in the build container (the reference tree does not travel to the GPU box); callers must
check `available()` first.  Used by oracle/make_golden.py and by the CPU tests that pin
oracle/restated.py to the real thing.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("MV_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "evals", "utils", "correspondence.py"))


def load():
    """-> (evals.utils.correspondence, evals.utils.transformations) of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import faiss_shim

    faiss_shim.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    corr = importlib.import_module("evals.utils.correspondence")
    tr = importlib.import_module("evals.utils.transformations")
    return corr, tr
