"""GPU probe 2: why is K1 slower in situ than under ncu?  Pair patterns + L2 flush."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import new_bloom_filter_repo_b200 as pkg
from new_bloom_filter_repo_b200 import _cabi as cabi
import bench
L, ctx = cabi.lib(), cabi.ctx()
F = 60
frames = np.zeros((F, 2160, 3840, 3), np.uint8)
bench.fill_stream(frames, 3)
st = pkg.FrameStream(2160, 3840, 3, np.uint8, max_frames=F)
st.upload(frames)
scratch = C.c_void_p(); cabi.check(L.rbf_malloc(ctx, 512 << 20, C.byref(scratch)), ctx)
def run(prev, curr, variant, flush, reps=4):
    cabi.check(L.rbf_set_option(ctx, b"k1_variant", variant), ctx)
    out = []
    for _ in range(reps):
        if flush:
            cabi.check(L.rbf_memset(ctx, scratch, 1, 512 << 20), ctx); cabi.check(L.rbf_sync(ctx), ctx)
        st.encode(prev, curr, 3.0)
        out.append(st.stage_ms()["k1_threshold"])
    return out
P = 29
pats = {"consecutive": (np.arange(P), np.arange(P) + 1), "same_pair": (np.zeros(P, int), np.ones(P, int)),
        "disjoint": (2 * np.arange(P), 2 * np.arange(P) + 1), "reverse": (np.arange(P)[::-1] + 1, np.arange(P)[::-1])}
for name, (p, c) in pats.items():
    for variant in (0, 1):
        for flush in (0, 1):
            r = run(p.astype(np.uint32), c.astype(np.uint32), variant, flush)
            print("%-12s variant %d flush %d  k1 ms %s  us/pair %.2f" % (name, variant, flush, ["%.3f" % x for x in r], 1e3 * min(r) / P), flush=True)
