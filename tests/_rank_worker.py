"""Worker for tests/test_multirank_cpu.py: world_size-2 gloo run of the host-side multi-GPU logic."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from new_bloom_filter_repo_b200 import distributed as rdist  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pairs = int(sys.argv[1])
    lo, hi = rdist.shard_pairs(pairs, rank, world)
    ident = rdist.broadcast_unique_id(dist, lambda: (np.arange(128) * 7 + 3).astype(np.uint8))
    slot = rdist.agree_slot_bytes(dist, 1000 + 777 * rank)
    # strong-scaling layout: per-rank header lists gathered as objects, rows of the gathered slot buffer in stream order
    lay = rdist.shard_layout(pairs, rank, world)
    table = [None] * world
    dist.all_gather_object(table, [(lay["lo"] + t, rank) for t in range(lay["pairs"])])
    headers = [h for part in table for h in part]
    rows = rdist.gathered_pair_rows(pairs, world)
    out = {"rank": rank, "world": world, "lo": lo, "hi": hi, "id_sum": int(ident.sum()), "id0": int(ident[1]), "slot": slot,
           "layout": lay, "headers": headers, "rows": rows}
    with open(os.path.join(sys.argv[2], "rank%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
