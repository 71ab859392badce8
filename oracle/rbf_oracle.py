"""
ORACLE (test infrastructure only) -- CPU restatement of the reference's
rational-Bloom-filter hot path.

    !!  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    !!  --impl reference legs may import this module.  The product package
    !!  (new_bloom_filter_repo_b200/) never does: it fails loudly without CUDA.

Parity pinning: the reference (ross39/new_bloom_filter_repo @ 7e37ed8) holds no
golden vectors of its own ("parity unpinned" by the
reference's tests).  This oracle is therefore pinned by outputs of the reference
itself, generated in the build container by tests/golden/make_golden.py (which
imports /root/reference) and committed under tests/golden/.  tests/test_oracle_*.py
check every function here against those fixtures.

Third-party arithmetic: XXH64 is python-xxhash (requirements.txt:9,
`xxhash>=2.0.0`; the container has 3.7.0 bundling libxxhash 0.8.2), not part of
/root/reference.  `xxh64()` below restates the published XXH64 algorithm; it is
cross-checked against the wheel when the wheel is importable.

Every function cites the reference file:line it follows
(ivc = improved_video_compressor.py, rbf = rational_bloom_filter.py,
bc = bloom_compress.py).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# XXH64 (libxxhash 0.8.2 semantics) -- call sites ivc:77,78,94; rbf:27,115,116,134
# ----------------------------------------------------------------------------
_M = (1 << 64) - 1
P1 = 0x9E3779B185EBCA87
P2 = 0xC2B2AE3D27D4EB4F
P3 = 0x165667B19E3779F9
P4 = 0x85EBCA77C2B2AE63
P5 = 0x27D4EB2F165667C5


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & _M


def _round(acc: int, lane: int) -> int:
    acc = (acc + lane * P2) & _M
    return (_rotl(acc, 31) * P1) & _M


def _merge(h: int, v: int) -> int:
    h ^= _round(0, v)
    return (h * P1 + P4) & _M


def xxh64(data: bytes, seed: int = 0) -> int:
    """XXH64 digest as an int; equals xxhash.xxh64_intdigest(data, seed)."""
    seed &= _M
    n = len(data)
    i = 0
    if n >= 32:
        v1 = (seed + P1 + P2) & _M
        v2 = (seed + P2) & _M
        v3 = seed
        v4 = (seed - P1) & _M
        while i + 32 <= n:
            v1 = _round(v1, int.from_bytes(data[i:i + 8], "little"))
            v2 = _round(v2, int.from_bytes(data[i + 8:i + 16], "little"))
            v3 = _round(v3, int.from_bytes(data[i + 16:i + 24], "little"))
            v4 = _round(v4, int.from_bytes(data[i + 24:i + 32], "little"))
            i += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & _M
        h = _merge(h, v1)
        h = _merge(h, v2)
        h = _merge(h, v3)
        h = _merge(h, v4)
    else:
        h = (seed + P5) & _M
    h = (h + n) & _M
    while i + 8 <= n:
        h ^= _round(0, int.from_bytes(data[i:i + 8], "little"))
        h = (_rotl(h, 27) * P1 + P4) & _M
        i += 8
    if i + 4 <= n:
        h ^= (int.from_bytes(data[i:i + 4], "little") * P1) & _M
        h = (_rotl(h, 23) * P2 + P3) & _M
        i += 4
    while i < n:
        h ^= (data[i] * P5) & _M
        h = (_rotl(h, 11) * P1) & _M
        i += 1
    h ^= h >> 33
    h = (h * P2) & _M
    h ^= h >> 29
    h = (h * P3) & _M
    h ^= h >> 32
    return h


def _item_bytes(item) -> bytes:
    """`str(item)` then UTF-8, as python-xxhash does for str input (ivc:77, rbf:115)."""
    return str(item).encode("utf-8")


# ----------------------------------------------------------------------------
# Seed sets of the three RationalBloomFilter variants
# ----------------------------------------------------------------------------
IVC_SEEDS = (0x12345678, 0x87654321, 999)     # ivc:62-63, ivc:94
BC_SEEDS = (0, 1, 999)                        # bc:163-164, bc:195


def rbf_seeds(k_star: float) -> Tuple[int, int, int]:
    """rbf:100-101 (h1=0, h2=1) and rbf:134 (activation seed = ceil(k*))."""
    return (0, 1, math.ceil(k_star))


# ----------------------------------------------------------------------------
# Rational Bloom filter over integer indices -- ivc:39-138
# ----------------------------------------------------------------------------
class RationalBloomFilter:
    """Restatement of ivc.RationalBloomFilter (ivc:39-138); `seeds` selects the variant."""

    def __init__(self, size: int, k_star: float, seeds: Tuple[int, int, int] = IVC_SEEDS):
        self.size = size                                  # ivc:55
        self.k_star = k_star                              # ivc:56
        self.floor_k = math.floor(k_star)                 # ivc:57
        self.p_activation = k_star - self.floor_k         # ivc:58
        self.bit_array = np.zeros(size, dtype=np.uint8)   # ivc:59
        self.h1_seed, self.h2_seed, self.act_seed = seeds

    def _get_hash_indices(self, item, i: int) -> int:
        b = _item_bytes(item)
        h1 = xxh64(b, self.h1_seed)                       # ivc:77
        h2 = xxh64(b, self.h2_seed)                       # ivc:78
        return (h1 + i * h2) % self.size                  # ivc:81 (unbounded ints)

    def _determine_activation(self, item) -> bool:
        h = xxh64(_item_bytes(item), self.act_seed)       # ivc:94
        return h / (2 ** 64 - 1) < self.p_activation      # ivc:95-97

    def add_index(self, index) -> None:                   # ivc:99-114
        for i in range(self.floor_k):
            self.bit_array[self._get_hash_indices(index, i)] = 1
        if self._determine_activation(index):
            self.bit_array[self._get_hash_indices(index, self.floor_k)] = 1

    def check_index(self, index) -> bool:                 # ivc:116-138
        for i in range(self.floor_k):
            if self.bit_array[self._get_hash_indices(index, i)] == 0:
                return False
        if self._determine_activation(index):
            if self.bit_array[self._get_hash_indices(index, self.floor_k)] == 0:
                return False
        return True

    # string-keyed aliases of rbf.RationalBloomFilter (rbf:139-182)
    add = add_index
    contains = check_index


class StandardBloomFilter:
    """rbf.StandardBloomFilter (rbf:9-71): k independent hashes xxh64(str(item), seed=i) % m."""

    def __init__(self, m: int, k: int):
        self.size = m
        self.hash_count = int(k)                          # rbf:22
        self.bit_array = np.zeros(m, dtype=np.uint8)

    def _hash(self, item, seed: int) -> int:
        return xxh64(_item_bytes(item), seed) % self.size  # rbf:27

    def add(self, item) -> None:                          # rbf:29-33
        for i in range(self.hash_count):
            self.bit_array[self._hash(item, i)] = 1

    def contains(self, item) -> bool:                     # rbf:35-41
        return all(self.bit_array[self._hash(item, i)] for i in range(self.hash_count))


def get_optimal_size(n: int, p: float) -> int:            # rbf:184-197
    return int(math.ceil(-(n * math.log(p)) / (math.log(2) ** 2)))


def get_optimal_hash_count(m: int, n: int) -> float:      # rbf:199-214
    return max(0.1, (m / n) * math.log(2))


# ----------------------------------------------------------------------------
# Exact integer restatements used by the CUDA path
# ----------------------------------------------------------------------------
def activation_threshold(p_activation: float) -> int:
    """T such that  h < T  <=>  h / (2**64 - 1) < p_activation  (ivc:95-97).

    The quotient is Python's correctly-rounded int/int true division and is
    monotone in h, so a binary search over h in [0, 2**64] is exact.
    Returns a value in [0, 2**64]; 2**64 cannot occur for p_activation <= 1.
    """
    d = 2 ** 64 - 1
    lo, hi = 0, 2 ** 64          # invariant: all h < lo satisfy; all h >= hi fail
    while lo < hi:
        mid = (lo + hi) // 2
        if mid / d < p_activation:
            lo = mid + 1
        else:
            hi = mid
    return lo


def hash_index_modular(h1: int, h2: int, i: int, m: int) -> int:
    """((h1 % m) + i*(h2 % m)) % m  ==  (h1 + i*h2) % m   (ivc:81)."""
    return ((h1 % m) + i * (h2 % m)) % m


# ----------------------------------------------------------------------------
# Bloom + witness coder for a 0/1 vector -- ivc:140-307
# ----------------------------------------------------------------------------
P_STAR = 0.32453                                          # ivc:150


def calculate_optimal_params(n: int, p: float):
    """ivc:161-196, expression for expression (order of float ops preserved)."""
    if p <= 0.0001:
        return 0, 0
    if p >= P_STAR:
        return 0, 0
    q = 1 - p
    L = math.log(2)
    k = math.log2(q * (L ** 2) / p)                       # ivc:185
    if math.isnan(k) or k <= 0:
        return 0, 0
    gamma = 1 / L
    l = int(p * n * k * gamma)                            # ivc:193
    return max(0.1, k), max(1, l)


def compress(binary_input: np.ndarray, seeds=IVC_SEEDS, k_l_override=None):
    """BloomFilterCompressor.compress (ivc:198-266).

    Returns (bitmap uint8[l] or the raw input, witness list, p, n, ratio, k, l);
    k == 0 marks the raw-passthrough branches (ivc:215-218, ivc:223-225).
    `k_l_override=(k, l)` replaces _calculate_optimal_params (BASELINE config 5).
    """
    n = len(binary_input)
    ones_count = np.sum(binary_input)                     # ivc:211
    p = ones_count / n                                    # ivc:212
    if p >= P_STAR:
        return binary_input, [], p, n, 1.0, 0, 0
    k, l = calculate_optimal_params(n, p) if k_l_override is None else k_l_override
    if l == 0 or l >= n:
        return binary_input, [], p, n, 1.0, 0, 0
    bf = RationalBloomFilter(l, k, seeds)
    for i in range(n):                                    # ivc:235-237
        if binary_input[i] == 1:
            bf.add_index(i)
    witness: List[int] = []
    for i in range(n):                                    # ivc:245-253
        if bf.check_index(i):
            witness.append(binary_input[i])
    ratio = (l + len(witness)) / n                        # ivc:256-258
    return bf.bit_array, witness, p, n, ratio, k, l


def decompress(bloom_bitmap: np.ndarray, witness: Sequence[int], n: int, k: float,
               seeds=IVC_SEEDS) -> np.ndarray:
    """BloomFilterCompressor.decompress (ivc:268-307)."""
    if len(witness) == 0:                                 # ivc:282-284
        return bloom_bitmap
    bf = RationalBloomFilter(len(bloom_bitmap), k, seeds)
    bf.bit_array = bloom_bitmap                           # ivc:290
    out = np.zeros(n, dtype=np.uint8)
    j = 0
    for i in range(n):                                    # ivc:299-304
        if bf.check_index(i):
            out[i] = witness[j]
            j += 1
    return out


# ----------------------------------------------------------------------------
# Frame-difference mask -- VideoFrameCompressor._calculate_frame_diff, ivc:784-808
# ----------------------------------------------------------------------------
def bgr2gray(a: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(a, cv2.COLOR_BGR2GRAY) for uint8 / uint16 (ivc:794-795).  OpenCV is a third-party dependency of the
    reference (requirements.txt, opencv-python; 4.13.0 here): its 8/16-bit path is the 15-bit fixed point
    (3735*B + 19235*G + 9798*R + 16384) >> 15, pinned by tests/golden/gray_kat.json (generated with the real cv2)."""
    x = a.astype(np.uint32)
    return ((x[:, :, 0] * 3735 + x[:, :, 1] * 19235 + x[:, :, 2] * 9798 + 16384) >> 15).astype(a.dtype)


def frame_diff_mask(prev: np.ndarray, curr: np.ndarray, threshold: float, gray: bool = False) -> np.ndarray:
    """|int16(prev) - int16(curr)| > threshold  -> uint8 H x W  (ivc:788-808) on the Y plane (use_direct_yuv=True, channel 0
    is Y) or, gray=True, on the BGR->gray conversion of the non-YUV colour branch (ivc:792-795); int16 wrap-around
    for 16-bit samples is numpy's and is part of the reference behaviour (ivc:801).
    """
    if gray:
        pg, cg = bgr2gray(prev), bgr2gray(curr)           # ivc:794-795
    elif prev.ndim > 2 and prev.shape[2] > 1:
        pg = prev[:, :, 0].copy()                         # ivc:790
        cg = curr[:, :, 0].copy()                         # ivc:791
    else:
        pg, cg = prev.copy(), curr.copy()                 # ivc:797-798
    with np.errstate(over="ignore"):
        diff = np.abs(pg.astype(np.int16) - cg.astype(np.int16))   # ivc:801
    return (diff > threshold).astype(np.uint8)            # ivc:808


def changed_values_yuv(curr: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Interleaved Y,U,V of changed pixels truncated to uint8 (ivc:811-829)."""
    rows, cols = np.where(mask == 1)
    vals = curr[rows, cols, :]
    return vals.astype(np.uint8).reshape(-1)              # ivc:825 (uint8 truncation)


def apply_frame_diff(base: np.ndarray, mask: np.ndarray, changed: np.ndarray) -> np.ndarray:
    """_apply_frame_diff for colour frames (ivc:849-909)."""
    out = base.copy()
    rows, cols = np.where(mask == 1)
    ch = base.shape[2]
    if len(changed) == len(rows) * ch:                    # ivc:882
        out[rows, cols] = changed.reshape(-1, ch)
    return out
