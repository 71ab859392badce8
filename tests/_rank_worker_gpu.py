"""Worker for tests/test_gpu_parity.py::test_two_ranks_p2p_equals_nccl (needs >= 2 GPUs): both exchanges of the packed bit arrays
deliver, on every rank, exactly what the owners hold."""
import hashlib
import json
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import new_bloom_filter_repo_b200 as pkg
    from new_bloom_filter_repo_b200 import distributed as rdist
    from tests.util import synth_stream
    total_frames = 23
    frames = synth_stream(270, 480, total_frames, 7, [0.05, 0.1, 0.02, 0.0, 0.2])
    out = {"rank": rank}
    ref = None
    for gather in ("nccl", "p2p"):
        enc = rdist.ShardedStreamEncoder(dist, 270, 480, 3, np.uint8, total_frames, gather=gather)
        lay = enc.layout
        enc.upload(frames[lay["first_frame"]: lay["first_frame"] + lay["frames"]])
        for _ in range(3):                                 # several exchanges: both halves of the peer buffers are used
            enc.encode(3.0)
        got = enc.gathered()
        hdr = enc.headers()
        out[gather] = {"used": enc.gather, "sha": hashlib.sha256(got.tobytes()).hexdigest(), "l": [h[0] for h in hdr]}
        if rank == 0 and ref is None:                      # single-GPU encode of the whole stream
            st = pkg.FrameStream(270, 480, 3, np.uint8, max_frames=total_frames)
            st.upload(frames)
            res = st.encode_consecutive(total_frames, 3.0)
            ref = np.zeros_like(got)
            for t, r in enumerate(res):
                bm = st.fetch(t, want_mask=False)[0]
                ref[t, :bm.size] = bm
            out["single_gpu_sha"] = hashlib.sha256(ref.tobytes()).hexdigest()
            out["single_gpu_l"] = [r.l for r in res]
            st.close()
        enc.close()
    with open(os.path.join(sys.argv[1], "rank%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
