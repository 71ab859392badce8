"""
ctypes binding of oracle/_build/librbf_oracle.so (the C oracle; test infrastructure).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librbf_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rbf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
        L.orc_hash_index_item.restype = C.c_uint64
        L.orc_hash_index_item.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_probe_index.restype = C.c_uint64
        L.orc_probe_index.argtypes = [C.c_uint64] * 4
        L.orc_unit_div.restype = C.c_double
        L.orc_unit_div.argtypes = [C.c_uint64]
        L.orc_activation_threshold.restype = C.c_uint64
        L.orc_activation_threshold.argtypes = [C.c_double]
        L.orc_optimal_params.restype = C.c_uint64
        L.orc_optimal_params.argtypes = [C.c_uint64, C.c_double, C.POINTER(C.c_double)]
        L.orc_compress_kl.restype = C.c_uint64
        L.orc_compress_kl.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64,
                                      C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_decompress_kl.restype = C.c_uint64
        L.orc_decompress_kl.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                        C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_filter_add_strings.restype = None
        L.orc_filter_add_strings.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64,
                                             C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_filter_check_strings.restype = None
        L.orc_filter_check_strings.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64,
                                               C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_frame_diff_mask.restype = C.c_uint64
        L.orc_frame_diff_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                          C.c_double, C.c_void_p]
        _lib = L
    return _lib


IVC_SEEDS = (0x12345678, 0x87654321, 999)
P_STAR = 0.32453


def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().orc_xxh64(data, len(data), seed & (2 ** 64 - 1))


def activation_threshold(p_act: float) -> int:
    return lib().orc_activation_threshold(float(p_act))


def optimal_params(n: int, p: float):
    k = C.c_double(0.0)
    l = lib().orc_optimal_params(int(n), float(p), C.byref(k))
    return k.value, int(l)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def compress_kl(mask: np.ndarray, k: float, l: int, seeds=IVC_SEEDS):
    """Insert + witness loops (ivc:232-253) for given (k, l) -> (bits uint8[l], witness uint8[w])."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    bits = np.zeros(l, dtype=np.uint8)
    wit = np.zeros(len(mask), dtype=np.uint8)
    w = lib().orc_compress_kl(_ptr(mask), len(mask), float(k), int(l), *[int(s) for s in seeds],
                              _ptr(bits), _ptr(wit))
    return bits, wit[:w].copy()


def compress(mask: np.ndarray, seeds=IVC_SEEDS, k_l_override=None):
    """BloomFilterCompressor.compress (ivc:198-266) -> (bitmap, witness, p, n, ratio, k, l); k=0 => raw."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    n = len(mask)
    ones = int(mask.sum(dtype=np.uint64))
    p = np.float64(ones) / np.float64(n)
    if p >= P_STAR:
        return mask, np.zeros(0, np.uint8), float(p), n, 1.0, 0.0, 0
    k, l = optimal_params(n, p) if k_l_override is None else k_l_override
    if l == 0 or l >= n:
        return mask, np.zeros(0, np.uint8), float(p), n, 1.0, 0.0, 0
    bits, wit = compress_kl(mask, k, l, seeds)
    return bits, wit, float(p), n, (l + len(wit)) / n, k, l


def decompress(bits: np.ndarray, witness: np.ndarray, n: int, k: float, seeds=IVC_SEEDS):
    """BloomFilterCompressor.decompress (ivc:268-307)."""
    if len(witness) == 0:
        return bits
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    witness = np.ascontiguousarray(witness, dtype=np.uint8)
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_decompress_kl(_ptr(bits), len(bits), _ptr(witness), len(witness), n, float(k),
                            *[int(s) for s in seeds], _ptr(out))
    return out


def _pack_strings(items):
    enc = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in items]
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(e) for e in enc])
    blob = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
    return blob, offs


def filter_add_strings(bits: np.ndarray, k: float, seeds, items):
    blob, offs = _pack_strings(items)
    lib().orc_filter_add_strings(_ptr(bits), len(bits), float(k), *[int(s) for s in seeds],
                                 _ptr(blob), _ptr(offs), len(items))


def filter_check_strings(bits: np.ndarray, k: float, seeds, items) -> np.ndarray:
    blob, offs = _pack_strings(items)
    out = np.zeros(len(items), dtype=np.uint8)
    lib().orc_filter_check_strings(_ptr(bits), len(bits), float(k), *[int(s) for s in seeds],
                                   _ptr(blob), _ptr(offs), len(items), _ptr(out))
    return out


def frame_diff_mask(prev: np.ndarray, curr: np.ndarray, threshold: float):
    """(mask uint8 H x W, ones) for interleaved H x W x C frames, Y = channel 0 (ivc:788-808)."""
    prev = np.ascontiguousarray(prev)
    curr = np.ascontiguousarray(curr)
    h, w = prev.shape[:2]
    ch = prev.shape[2] if prev.ndim == 3 else 1
    mask = np.zeros(h * w, dtype=np.uint8)
    ones = lib().orc_frame_diff_mask(_ptr(prev), _ptr(curr), h * w, ch, prev.dtype.itemsize,
                                     float(threshold), _ptr(mask))
    return mask.reshape(h, w), int(ones)
