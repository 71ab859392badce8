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


def test_bench_counters_go_stale_with_the_sources(tmp_path, monkeypatch):
    """bench.py reports roofline.traffic / issue_frac only while the kernel sources still hash to what was profiled."""
    import json
    import bench
    c = bench.kernel_counters("k_query4")
    assert c is not None and c["stale"] is False and c["dram_bytes_per_pair"] > 1e5 and 50 < c["thread_inst_per_px"] < 200
    fake = tmp_path / "profiles"
    fake.mkdir()
    c2 = dict(c, sources_sha16="0" * 16)
    c2.pop("stale")
    (fake / "r02_k_query4_counters.json").write_text(json.dumps(c2))
    import os
    for name in c["sources"]:                                  # same sources, different recorded hash -> stale
        os.makedirs(os.path.dirname(tmp_path / name), exist_ok=True)
        (tmp_path / name).write_bytes(open(os.path.join(bench.ROOT, name), "rb").read())
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.kernel_counters("k_query4")["stale"] is True
    assert bench.kernel_counters("k_no_such_kernel") is None


def test_bench_stream_frames_equals_fill_stream_slices():
    """Strong scaling: a rank builds only frames lo..hi of the shared stream; they must equal the full stream's slice."""
    import bench
    for dt in (np.uint8, np.uint16):
        full = np.empty((9, 24, 40, 3), dt)
        bench.fill_stream(full, seed=3)
        for lo, hi in ((0, 3), (3, 8), (7, 8)):
            out = np.empty((hi - lo + 1, 24, 40, 3), dt)
            bench.stream_frames(9, 24, 40, 3, lo, hi, out)
            assert np.array_equal(out, full[lo:hi + 1])


def test_bench_host_cores_is_positive_and_bounded():
    import os
    import bench
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
