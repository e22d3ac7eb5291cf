"""Worker of tests/test_multi_rank.py: one rank of a world_size-N gloo job.  Each rank steps its shard of the global
batch (on the emulated-warp debug build: no GPU in the CPU suite) and the counters are all-reduced."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    sys.path.insert(0, p)

import numpy as np
import torch.distributed as dist


def main():
    out_path, n_total, cycles = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_backend
    emu_backend.use_emulator(rebuild=False)
    import golden_io
    import microrts_b200 as M
    import parity as P
    from microrts_b200 import sharding
    maps = golden_io.load_maps()
    key = "8x8/basesWorkers8x8"
    utt = M.UnitTypeTable(1, 1)
    first, count = sharding.shard(n_total, rank, world)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), count)
    b.reset(sharding.global_seeds(1234, first, count))
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.step(cycles, cycles)
    total = sharding.reduce_stats_host(b.stats())
    res = b.results()
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, res.tolist()))
    if rank == 0:
        rows = []
        for f, r in sorted(gathered):
            rows.extend(r)
        json.dump(dict(stats=total, results=rows), open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
