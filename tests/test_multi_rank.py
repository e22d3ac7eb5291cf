"""N>1 path on CPU: world_size-2 (and 3) gloo jobs shard a global batch by rank; the reduced counters and the per-game
results must equal the single-rank run (results are independent of the number of ranks)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition():
    from microrts_b200 import sharding
    for n in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _c1) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
    s = sharding.global_seeds(10, 4, 3)
    assert s.tolist() == [14, 15, 16]


def run_world(world, n_total, cycles, tmp_path):
    subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build.sh")])
    out = tmp_path / ("w%d.json" % world)
    port = 29500 + os.getpid() % 2000 + world
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), str(out), str(n_total), str(cycles)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    return json.load(open(out))


@pytest.mark.timeout(900)
def test_gloo_world_sizes_agree(tmp_path):
    one = run_world(1, 6, 400, tmp_path)
    two = run_world(2, 6, 400, tmp_path)
    three = run_world(3, 6, 400, tmp_path)
    assert one["stats"]["cycles"] > 0
    assert one == two == three
