"""Shared helpers for the tests: golden-fixture loading and the seeded synthetic inputs
(the same generators tests/golden/make_golden.py used, so inputs can be re-created)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mask_for(case) -> np.ndarray:
    rng = np.random.default_rng(case["seed"])
    return (rng.random(case["n"]) < case["p_gen"]).astype(np.uint8)


def synth_pair(h, w, seed, p_change, dtype=np.uint8):
    rng = np.random.default_rng(seed)
    hi = 256 if dtype == np.uint8 else 65536
    prev = rng.integers(0, hi, (h, w, 3)).astype(dtype)
    curr = prev.copy()
    ch = rng.random((h, w)) < p_change
    delta = 64 if dtype == np.uint8 else 16384
    curr[ch] = (curr[ch].astype(np.int64) + delta) % hi
    return prev, curr


def golden_pair(rec):
    dt = np.dtype(rec["dtype"]).type
    prev, curr = synth_pair(rec["h"], rec["w"], rec["seed"], rec["p_change"], dt)
    if dt == np.uint16:
        prev[0, 0, 0], curr[0, 0, 0] = 0, 40000
        prev[0, 1, 0], curr[0, 1, 0] = 65535, 0
        prev[0, 2, 0], curr[0, 2, 0] = 0, 32768
        prev[0, 3, 0], curr[0, 3, 0] = 32768, 0
        prev[0, 4, 0], curr[0, 4, 0] = 32767, 65535
    return prev, curr


def gray_pair(rec):
    """Input pair of a tests/golden/gray_kat.json case (same construction as make_golden.gen_gray)."""
    dt = np.dtype(rec["dtype"]).type
    h, w, seed, pc = rec["h"], rec["w"], rec["seed"], rec["p_change"]
    hi = 256 if dt == np.uint8 else 65536
    prev, curr = synth_pair(h, w, seed, pc, dt)
    if rec["threshold"] is None:
        rng = np.random.default_rng(seed)
        yy, xx = np.mgrid[0:h, 0:w]
        base = ((yy + xx) // 4 + 20).astype(np.int64) * (1 if dt == np.uint8 else 200)
        prev = np.stack([base + rng.integers(0, 3, (h, w)) for _ in range(3)], axis=-1).astype(dt)
        curr = prev.copy()
        ch = rng.random((h, w)) < pc
        curr[ch] = (curr[ch].astype(np.int64) + hi // 4) % hi
    curr[0, 0] = prev[0, 0]
    curr[0, 0, 0] = (int(prev[0, 0, 0]) + 9) % hi
    return prev, curr


def gray_cube():
    grid = np.array(sorted(set(list(range(0, 256, 5)) + [1, 2, 254, 255])), dtype=np.uint8)
    bb, gg, rr = np.meshgrid(grid, grid, grid, indexing="ij")
    return np.stack([bb, gg, rr], axis=-1).reshape(1, -1, 3)


def synth_stream(h, w, frames, seed, p_seq, dtype=np.uint8):
    """SURVEY.md 8(d) generator: frame0 = gradient + small noise; frame_t = frame_{t-1} with
    Bernoulli(p_t) pixels having all channels += 64 (mod range).  Returns uint array [frames,h,w,3]."""
    hi = 256 if dtype == np.uint8 else 65536
    delta = 64 if dtype == np.uint8 else 16384
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = ((yy * 3 + xx * 2) % (hi - 16)).astype(np.int64)
    f0 = np.stack([base, (base // 2 + 7) % hi, (base // 3 + 90) % hi], axis=-1)
    f0 = (f0 + rng.integers(0, 8, (h, w, 3))) % hi
    out = np.empty((frames, h, w, 3), dtype=dtype)
    out[0] = f0.astype(dtype)
    for t in range(1, frames):
        r = np.random.default_rng(seed + t)
        ch = r.random((h, w)) < p_seq[(t - 1) % len(p_seq)]
        nxt = out[t - 1].astype(np.int64)
        nxt[ch] = (nxt[ch] + delta) % hi
        out[t] = nxt.astype(dtype)
    return out
