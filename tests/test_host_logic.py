"""CPU tests of host-side glue: keyframe codec wire format (byte-identical to the reference's
FixedVideoCompressor), YUVFrame wrapper, shard partitioning."""
import hashlib

import numpy as np

from tests.util import golden_json, golden_npz


def test_keyframe_payload_bytes_match_reference():
    from new_bloom_filter_repo_b200.fixed_video_compressor import FixedVideoCompressor
    g = golden_json("keyframe_kat.json")
    arr = golden_npz("keyframe_arrays.npz")
    comp = FixedVideoCompressor(verbose=False)
    for rec in g["cases"]:
        f = arr[rec["name"] + "/frame"]
        fr = comp.add_yuv_info_to_frame(f) if rec["yuv"] else f
        payload = comp.compress_frame(fr)
        assert hashlib.sha256(payload).hexdigest() == rec["payload_sha256"]
        assert payload == arr[rec["name"] + "/payload"].tobytes()
        back = comp.decompress_frame(payload)
        data = back.data if hasattr(back, "yuv_info") else back
        assert np.array_equal(data, f) and data.dtype == f.dtype
        assert hasattr(back, "yuv_info") == rec["yuv"]
        res = comp.verify_lossless([fr], [back])
        assert res["lossless"] and res["exact_frame_matches"] == 1


def test_verify_lossless_detects_difference():
    from new_bloom_filter_repo_b200.fixed_video_compressor import FixedVideoCompressor
    comp = FixedVideoCompressor(verbose=False)
    a = np.zeros((4, 4, 3), np.uint8)
    b = a.copy(); b[1, 1, 1] = 9
    r = comp.verify_lossless([a, a], [a, b])
    assert not r["lossless"] and r["diff_frames"] == [1] and r["max_diff_frame"] == 1
    assert comp.verify_lossless([a], [a, a])["lossless"] is False


def test_shard_partition():
    from new_bloom_filter_repo_b200.distributed import shard_pairs
    for pairs in (1, 7, 8, 29, 290, 299):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_pairs(pairs, r, world)
                assert 0 <= lo <= hi <= pairs
                seen += list(range(lo, hi))
            assert seen == list(range(pairs))
            sizes = [shard_pairs(pairs, r, world)[1] - shard_pairs(pairs, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_keyframe_async_assembly_equals_serial():
    """The GOP loop compresses a keyframe's zlib blobs as separate pool jobs: the assembled payload is byte-identical."""
    from concurrent.futures import ThreadPoolExecutor
    from new_bloom_filter_repo_b200.fixed_video_compressor import FixedVideoCompressor
    c = FixedVideoCompressor(verbose=False)
    rng = np.random.default_rng(0)
    with ThreadPoolExecutor(3) as pool:
        for shape, dt in (((48, 64, 3), np.uint8), ((32, 40), np.uint8), ((20, 24, 3), np.uint16)):
            f = rng.integers(0, 200, shape).astype(dt)
            assert c.compress_frame_async(f, pool)() == c.compress_frame(f)
            if f.ndim == 3:
                y = c.add_yuv_info_to_frame(f)
                assert c.compress_frame_async(y, pool)() == c.compress_frame(y)
