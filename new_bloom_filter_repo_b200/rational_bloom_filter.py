"""
Drop-in for the reference's rational_bloom_filter.py (StandardBloomFilter rbf:9-71,
RationalBloomFilter rbf:74-214): same constructor arguments, attributes and methods,
bit-identical bit arrays and answers -- computed by the sm_100a kernels behind the C ABI
(k_items_str in csrc/rbf_kernels.cu).  There is no host implementation here: without the
CUDA library and a B200 the constructors raise.

`add_many` / `contains_many` are batch forms of `add` / `contains` (one kernel launch for the
whole list); `add` / `contains` keep the reference's one-item signature.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List

import numpy as np

from . import _cabi


def _pack(items: Iterable) -> tuple:
    enc = [str(it).encode("utf-8") for it in items]            # str(item), rbf:27,115
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        offs[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    blob = np.frombuffer(b"".join(enc) + b"\0" * 16, dtype=np.uint8)
    return blob, offs, len(enc)


class _DeviceFilter:
    """Owns one rbf_filter handle (device-resident bit array)."""

    def __init__(self, size: int, k_star: float, seeds):
        if int(size) != size or size <= 0:
            raise ValueError("filter size must be a positive integer")
        self._h = C.c_void_p()
        sd = _cabi.seeds_struct(seeds)
        _cabi.check(_cabi.lib().rbf_filter_create(_cabi.ctx(), int(size), float(k_star), C.byref(sd), C.byref(self._h)),
                    _cabi.ctx())
        self._n = int(size)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _cabi.lib().rbf_filter_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def get_bits(self) -> np.ndarray:
        out = np.empty(self._n, dtype=np.uint8)
        _cabi.check(_cabi.lib().rbf_filter_get_bits(self._h, _cabi.ptr(out)), _cabi.ctx())
        return out

    def set_bits(self, bits) -> None:
        a = np.ascontiguousarray(np.asarray(bits, dtype=np.uint8))
        if a.shape != (self._n,):
            raise ValueError("bit_array must have %d entries" % self._n)
        _cabi.check(_cabi.lib().rbf_filter_set_bits(self._h, _cabi.ptr(a)), _cabi.ctx())

    def add_strings(self, items, standard_k=0) -> None:
        blob, offs, n = _pack(items)
        if n:
            _cabi.check(_cabi.lib().rbf_filter_add_strings(self._h, _cabi.ptr(blob), _cabi.ptr(offs), n, standard_k), _cabi.ctx())

    def check_strings(self, items, standard_k=0) -> np.ndarray:
        blob, offs, n = _pack(items)
        out = np.zeros(n, dtype=np.uint8)
        if n:
            _cabi.check(_cabi.lib().rbf_filter_check_strings(self._h, _cabi.ptr(blob), _cabi.ptr(offs), n, standard_k,
                                                             _cabi.ptr(out)), _cabi.ctx())
        return out

    def add_indices(self, idx) -> None:
        a = np.ascontiguousarray(np.asarray(idx, dtype=np.uint32))
        if a.size:
            _cabi.check(_cabi.lib().rbf_filter_add_indices(self._h, _cabi.ptr(a), a.size), _cabi.ctx())

    def check_indices(self, idx) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(idx, dtype=np.uint32))
        out = np.zeros(a.size, dtype=np.uint8)
        if a.size:
            _cabi.check(_cabi.lib().rbf_filter_check_indices(self._h, _cabi.ptr(a), a.size, _cabi.ptr(out)), _cabi.ctx())
        return out


class StandardBloomFilter:
    """rbf.StandardBloomFilter (rbf:9-71): k independent hashes xxh64(str(item), seed=i) % m."""

    def __init__(self, m: int, k: int):
        self.size = m                                   # rbf:21
        self.hash_count = int(k)                        # rbf:22
        self._dev = _DeviceFilter(m, float(max(self.hash_count, 0)), (0, 1, 0))

    @property
    def bit_array(self) -> List[int]:                   # rbf:23 keeps a Python list
        return self._dev.get_bits().tolist()

    @bit_array.setter
    def bit_array(self, bits) -> None:
        self._dev.set_bits(bits)

    def _hash(self, item: str, seed: int) -> int:      # rbf:25-27
        return _cabi.xxh64(str(item).encode("utf-8"), seed) % self.size

    def add(self, item: str) -> None:                   # rbf:29-33
        self.add_many([item])

    def contains(self, item: str) -> bool:              # rbf:35-41
        return bool(self.contains_many([item])[0])

    def add_many(self, items) -> None:
        if self.hash_count > 0:
            self._dev.add_strings(list(items), standard_k=self.hash_count)

    def contains_many(self, items) -> np.ndarray:
        items = list(items)
        if self.hash_count <= 0:
            return np.ones(len(items), dtype=bool)
        return self._dev.check_strings(items, standard_k=self.hash_count).astype(bool)

    @staticmethod
    def get_optimal_size(n: int, p: float) -> int:      # rbf:43-56
        m = -(n * math.log(p)) / (math.log(2) ** 2)
        return int(math.ceil(m))

    @staticmethod
    def get_optimal_hash_count(m: int, n: int) -> int:  # rbf:58-71
        k = (m / n) * math.log(2)
        return max(1, int(round(k)))


class RationalBloomFilter:
    """rbf.RationalBloomFilter (rbf:74-214): floor(k*) double-hashed probes plus one probe that is
    applied when xxh64(item, seed=ceil(k*)) / (2**64-1) < k* - floor(k*)."""

    def __init__(self, m: int, k_star: float):
        self.size = m                                   # rbf:92
        self.k_star = k_star                            # rbf:93
        self.floor_k = math.floor(k_star)               # rbf:94
        self.ceil_k = math.ceil(k_star)                 # rbf:95
        self.p_activation = k_star - self.floor_k       # rbf:96
        self.h1_seed = 0                                # rbf:100
        self.h2_seed = 1                                # rbf:101
        self._dev = _DeviceFilter(m, k_star, (self.h1_seed, self.h2_seed, self.ceil_k))

    @property
    def bit_array(self) -> List[int]:
        return self._dev.get_bits().tolist()

    @bit_array.setter
    def bit_array(self, bits) -> None:
        self._dev.set_bits(bits)

    def _get_hash_indices(self, item: str, i: int) -> int:       # rbf:103-119 (scalar helper, parameters only)
        b = str(item).encode("utf-8")
        return _cabi.lib().rbf_probe_index(_cabi.xxh64(b, self.h1_seed), _cabi.xxh64(b, self.h2_seed), int(i), int(self.size))

    def _determine_activation(self, item: str) -> bool:          # rbf:121-137
        return _cabi.xxh64(str(item).encode("utf-8"), self.ceil_k) < _cabi.activation_threshold(self.p_activation)

    def add(self, item: str) -> None:                             # rbf:139-156
        self._dev.add_strings([item])

    def contains(self, item: str) -> bool:                        # rbf:158-182
        return bool(self._dev.check_strings([item])[0])

    def add_many(self, items) -> None:
        self._dev.add_strings(list(items))

    def contains_many(self, items) -> np.ndarray:
        return self._dev.check_strings(list(items)).astype(bool)

    @staticmethod
    def get_optimal_size(n: int, p: float) -> int:                # rbf:184-197
        m = -(n * math.log(p)) / (math.log(2) ** 2)
        return int(math.ceil(m))

    @staticmethod
    def get_optimal_hash_count(m: int, n: int) -> float:          # rbf:199-214
        k_star = (m / n) * math.log(2)
        return max(0.1, k_star)
