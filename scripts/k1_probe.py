"""GPU probe: K1 (threshold) stage time vs number of pairs / variant, through the C ABI timer."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import new_bloom_filter_repo_b200 as pkg
from new_bloom_filter_repo_b200 import _cabi as cabi
import bench
L, ctx = cabi.lib(), cabi.ctx()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 60
frames = np.zeros((F, 2160, 3840, 3), np.uint8)
bench.fill_stream(frames, 3)
st = pkg.FrameStream(2160, 3840, 3, np.uint8, max_frames=F)
st.upload(frames)
def t_k1(npairs, variant, reps=5):
    cabi.check(L.rbf_set_option(ctx, b"k1_variant", variant), ctx)
    cabi.check(L.rbf_set_option(ctx, b"k1_only", 1), ctx)
    idx = np.arange(npairs + 1, dtype=np.uint32)
    out = []
    for _ in range(reps):
        cabi.check(L.rbf_timer_start(ctx), ctx)
        st.encode(idx[:-1], idx[1:], 3.0)
        ms = C.c_double(); cabi.check(L.rbf_timer_stop_ms(ctx, C.byref(ms)), ctx)
        out.append(ms.value)
    cabi.check(L.rbf_set_option(ctx, b"k1_only", 0), ctx)
    return out
for variant in (0, 1):
    for npairs in (1, 2, 8, 29, F - 1):
        r = t_k1(npairs, variant)
        print("variant", variant, "pairs", npairs, "ms", ["%.3f" % x for x in r], "us/pair %.2f" % (1e3 * min(r) / npairs), flush=True)
# full pipeline stage times
cabi.check(L.rbf_set_option(ctx, b"k1_variant", 0), ctx)
for npairs in (8, 29, F - 1):
    idx = np.arange(npairs + 1, dtype=np.uint32)
    for _ in range(3):
        st.encode(idx[:-1], idx[1:], 3.0)
    print("pairs", npairs, {k: round(v, 3) for k, v in st.stage_ms().items()}, flush=True)
