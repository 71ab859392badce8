"""
Keyframe codec + YUV frame wrapper: host glue kept byte-compatible with the reference's
fixed_video_compressor.py (FixedVideoCompressor fvc:15-334) so that keyframe payloads written by
either implementation decode with the other.  zlib entropy coding is CPU work outside the
accelerated path; nothing here touches the GPU.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List

import numpy as np


class YUVFrame:
    """ndarray wrapper carrying the separate Y/U/V planes (fvc:289-334 `add_yuv_info_to_frame`)."""

    def __init__(self, data: np.ndarray, yuv_info: Dict = None):
        self.data = data
        if yuv_info is None:
            yuv_info = {"format": "YUV444", "y_plane": data[:, :, 0].copy(), "u_plane": data[:, :, 1].copy(),
                        "v_plane": data[:, :, 2].copy()}
        self.yuv_info = yuv_info

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    nbytes = property(lambda self: self.data.nbytes)
    size = property(lambda self: self.data.size)
    T = property(lambda self: self.data.T)

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def copy(self):
        return YUVFrame(self.data.copy(), {k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.yuv_info.items()})

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    def tobytes(self):
        return self.data.tobytes()

    def astype(self, dtype):
        return self.data.astype(dtype)

    def flatten(self):
        return self.data.flatten()

    def reshape(self, *a, **k):
        return self.data.reshape(*a, **k)


_DTYPES = {1: np.uint8, 2: np.uint16}


def _put_blob(parts: List[bytes], raw: bytes) -> None:
    z = zlib.compress(raw, 9)
    parts.append(struct.pack("<I", len(z)))
    parts.append(z)


class FixedVideoCompressor:
    """Per-frame zlib level-9 codec (fvc:15-285): `<III` h, w, itemsize | `<I` len | zlib(frame) |
    `<B` has_yuv [| `<H` len fmt | fmt | 3 x (`<I` len | zlib(plane) | `<II` shape)]."""

    def __init__(self, verbose=True):
        self.verbose = verbose

    def compress_frame(self, frame) -> bytes:                                    # fvc:27-74
        parts: List[bytes] = [struct.pack("<III", frame.shape[0], frame.shape[1], frame.dtype.itemsize)]
        _put_blob(parts, frame.tobytes())
        info = getattr(frame, "yuv_info", None)
        parts.append(struct.pack("<B", 1 if info is not None else 0))
        if info is not None:
            fmt = info.get("format", "YUV444").encode("utf-8")
            parts.append(struct.pack("<H", len(fmt)))
            parts.append(fmt)
            for name in ("y_plane", "u_plane", "v_plane"):
                plane = info[name]
                _put_blob(parts, plane.tobytes())
                parts.append(struct.pack("<II", *plane.shape))
        return b"".join(parts)

    def compress_frame_async(self, frame, pool):
        """compress_frame with the (up to four) zlib level-9 blobs as separate jobs of `pool`; returns a callable that assembles
        exactly the bytes compress_frame(frame) returns.  The keyframe of a GOP is the longest serial piece of compress_video."""
        info = getattr(frame, "yuv_info", None)
        raws = [frame.tobytes()] + ([info[name].tobytes() for name in ("y_plane", "u_plane", "v_plane")] if info is not None else [])
        futs = [pool.submit(zlib.compress, r, 9) for r in raws]

        def assemble() -> bytes:
            z = [f.result() for f in futs]
            parts: List[bytes] = [struct.pack("<III", frame.shape[0], frame.shape[1], frame.dtype.itemsize),
                                  struct.pack("<I", len(z[0])), z[0], struct.pack("<B", 1 if info is not None else 0)]
            if info is not None:
                fmt = info.get("format", "YUV444").encode("utf-8")
                parts += [struct.pack("<H", len(fmt)), fmt]
                for zi, name in zip(z[1:], ("y_plane", "u_plane", "v_plane")):
                    parts += [struct.pack("<I", len(zi)), zi, struct.pack("<II", *info[name].shape)]
            return b"".join(parts)
        return assemble

    def decompress_frame(self, data: bytes):                                     # fvc:76-181
        h, w, isz = struct.unpack_from("<III", data, 0)
        (zlen,) = struct.unpack_from("<I", data, 12)
        pos = 16
        raw = zlib.decompress(data[pos:pos + zlen])
        pos += zlen
        dtype = _DTYPES.get(isz, np.float32)
        gray = h * w * isz
        if len(raw) > gray and len(raw) % gray == 0:
            frame = np.frombuffer(raw, dtype=dtype).reshape((h, w, len(raw) // gray))
        else:
            frame = np.frombuffer(raw, dtype=dtype).reshape((h, w))
        has_yuv = pos < len(data) and data[pos] == 1
        pos += 1
        if not has_yuv:
            return frame
        (flen,) = struct.unpack_from("<H", data, pos)
        pos += 2
        info = {"format": data[pos:pos + flen].decode("utf-8")}
        pos += flen
        for name in ("y_plane", "u_plane", "v_plane"):
            (zl,) = struct.unpack_from("<I", data, pos)
            pos += 4
            plane_raw = zlib.decompress(data[pos:pos + zl])
            pos += zl
            ph, pw = struct.unpack_from("<II", data, pos)
            pos += 8
            info[name] = np.frombuffer(plane_raw, dtype=np.uint8).reshape((ph, pw))
        return YUVFrame(frame, info)

    def compress_video(self, frames) -> List[bytes]:                             # fvc:183-198
        return [self.compress_frame(f) for f in frames]

    def decompress_video(self, compressed_frames) -> list:                       # fvc:200-215
        return [self.decompress_frame(c) for c in compressed_frames]

    def verify_lossless(self, original_frames, decompressed_frames) -> Dict:     # fvc:217-285
        if len(original_frames) != len(decompressed_frames):
            return {"lossless": False,
                    "reason": f"Frame count mismatch: {len(original_frames)} vs {len(decompressed_frames)}",
                    "avg_difference": float("inf")}
        exact, diff_frames, max_diff, max_diff_frame = 0, [], 0, -1
        for i, (o, d) in enumerate(zip(original_frames, decompressed_frames)):
            od = o.data if hasattr(o, "yuv_info") else np.asarray(o)
            dd = d.data if hasattr(d, "yuv_info") else np.asarray(d)
            if np.array_equal(od, dd):
                exact += 1
                continue
            fd = np.mean(np.abs(od.astype(np.float32) - dd.astype(np.float32)))
            diff_frames.append(i)
            if fd > max_diff:
                max_diff, max_diff_frame = fd, i
        ok = exact == len(original_frames)
        res = {"lossless": ok, "exact_lossless": ok, "avg_difference": 0.0 if not diff_frames else max_diff,
               "max_difference": max_diff, "max_diff_frame": max_diff_frame, "exact_frame_matches": exact,
               "total_frames": len(original_frames), "diff_frames": diff_frames}
        if self.verbose:
            print(f"Lossless verification: {'SUCCESS' if ok else 'FAILED'}")
            print(f"Exact frame matches: {exact}/{len(original_frames)}")
        return res

    def add_yuv_info_to_frame(self, yuv_frame):                                  # fvc:287-334
        return YUVFrame(yuv_frame)
