"""
ORACLE / CPU BASELINE (test infrastructure only): loop-for-loop Python port of the reference's
CPU path for the hot path, using the reference's own hash dependency (python-xxhash) when it
is importable and the oracle's pure-Python XXH64 otherwise.  This is what `bench.py --impl
reference` and bench.py's `cpu_baseline` time on the GPU box's host cores (the reference itself
is Python source under /root/reference and does not travel to the GPU box).

    frame pair --_calculate_frame_diff mask (ivc:788-808)--> flat mask
               --BloomFilterCompressor.compress (ivc:198-266)--> bit array, witness

Structure kept exactly as the reference executes it: per item `str(item)`, two XXH64 per probe
recomputed inside the probe loop (ivc:77-78 inside ivc:107-108 / ivc:127-128), big-int modulo,
one byte per Bloom bit, Python list witness.
"""
import math

import numpy as np

try:                                        # the reference's dependency (requirements.txt:9)
    import xxhash as _xx

    def _h(s, seed):
        return _xx.xxh64_intdigest(s, seed)
    HASH_IMPL = "python-xxhash " + _xx.VERSION
except Exception:                           # pragma: no cover
    from oracle.rbf_oracle import xxh64 as _px

    def _h(s, seed):
        return _px(s.encode(), seed)
    HASH_IMPL = "pure-python"

P_STAR = 0.32453


class RationalBloomFilter:                                   # ivc:39-138
    def __init__(self, size, k_star):
        self.size = size
        self.k_star = k_star
        self.floor_k = math.floor(k_star)
        self.p_activation = k_star - self.floor_k
        self.bit_array = np.zeros(size, dtype=np.uint8)
        self.h1_seed = 0x12345678
        self.h2_seed = 0x87654321

    def _get_hash_indices(self, item, i):
        h1 = _h(str(item), self.h1_seed)
        h2 = _h(str(item), self.h2_seed)
        return (h1 + i * h2) % self.size

    def _determine_activation(self, item):
        return _h(str(item), 999) / (2 ** 64 - 1) < self.p_activation

    def add_index(self, index):
        for i in range(self.floor_k):
            self.bit_array[self._get_hash_indices(index, i)] = 1
        if self._determine_activation(index):
            self.bit_array[self._get_hash_indices(index, self.floor_k)] = 1

    def check_index(self, index):
        for i in range(self.floor_k):
            if self.bit_array[self._get_hash_indices(index, i)] == 0:
                return False
        if self._determine_activation(index):
            if self.bit_array[self._get_hash_indices(index, self.floor_k)] == 0:
                return False
        return True


def calculate_optimal_params(n, p):                           # ivc:161-196
    if p <= 0.0001 or p >= P_STAR:
        return 0, 0
    L = math.log(2)
    k = math.log2((1 - p) * (L ** 2) / p)
    if math.isnan(k) or k <= 0:
        return 0, 0
    return max(0.1, k), max(1, int(p * n * k * (1 / L)))


def compress(binary_input):                                   # ivc:198-266
    n = len(binary_input)
    p = np.sum(binary_input) / n
    if p >= P_STAR:
        return binary_input, [], p, n, 1.0
    k, l = calculate_optimal_params(n, p)
    if l == 0 or l >= n:
        return binary_input, [], p, n, 1.0
    bf = RationalBloomFilter(l, k)
    for i in range(n):
        if binary_input[i] == 1:
            bf.add_index(i)
    witness = []
    for i in range(n):
        if bf.check_index(i):
            witness.append(binary_input[i])
    return bf.bit_array, witness, p, n, (l + len(witness)) / n


def frame_diff_mask(prev, curr, threshold):                   # ivc:788-808 (direct-YUV branch)
    diff = np.abs(prev[:, :, 0].copy().astype(np.int16) - curr[:, :, 0].copy().astype(np.int16))
    return (diff > threshold).astype(np.uint8)


def encode_pair(prev, curr, threshold=3.0):
    """One unit of the metric: mask + Bloom insert + query of every position.  Returns pixels done."""
    mask = frame_diff_mask(prev, curr, threshold)
    compress(mask.flatten())
    return mask.size
