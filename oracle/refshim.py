"""
Import shim for the REAL reference (container only; /root/reference does not exist
on the GPU box).  Used by tests/golden/make_golden.py and by the optional live
differential tests (skipped when the reference is absent).  Test infrastructure.

The reference imports matplotlib at module top (ivc:31, rbf:5, bc:4), which is not
installed; empty stub modules satisfy the import without touching the reference.
`VideoFrameCompressor.bloom_compressor` is never assigned by the reference
(SURVEY.md section 0.2), so `make_vfc()` injects it.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("RBF_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "improved_video_compressor.py"))


def load():
    """Returns (ivc, rbf, bc, fvc) reference modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import improved_video_compressor as ivc
    import rational_bloom_filter as rbf
    import bloom_compress as bc
    import fixed_video_compressor as fvc
    return ivc, rbf, bc, fvc


def make_vfc(ivc, **kw):
    vfc = ivc.VideoFrameCompressor(**kw)
    vfc.bloom_compressor = ivc.BloomFilterCompressor()
    return vfc
