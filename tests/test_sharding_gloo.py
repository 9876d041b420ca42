"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): byte-range sharding with the counted record
phase, rank-ordered concatenation and counter reduction.  The per-shard processor is the CPU oracle
(test infrastructure) so the test needs no GPU; test_sharding_gpu runs the same plan on the CUDA path."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _worker(rank, world, port, fixture, qualtype, mode_name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as orc
    from sickle_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = open(os.path.join(GOLDEN, fixture), "rb").read()
    mode = {"se": orc.MODE_SE, "pei": orc.MODE_PE_INTER}[mode_name]
    lpu = 8 if mode_name == "pei" else 4
    counts = sharding.gather_newline_counts(data, world, rank, dist)
    b = sharding.shard_bounds(data, world, lpu, counts)
    shard = data[b[rank]:b[rank + 1]]
    r = orc.run(mode, orc.make_params(qualtype), shard, batch_len=1 << 40)
    assert r["rc"] == 0
    outs = [None] * world
    dist.all_gather_object(outs, r["out"])
    total = sharding.reduce_counters({k: v for k, v in r["counters"].items() if isinstance(v, int)}, dist, world)
    worst = sharding.reduce_max(float(rank + 1), dist, world)
    if rank == 0:
        q.put((b, [b"".join(o[s] for o in outs) for s in range(3)], total, worst))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("fixture,qualtype,mode_name", [("se_r150.fastq", "sanger", "se"),
                                                         ("varlen_illumina.fastq", "illumina", "se"),
                                                         ("il15_inter.fastq", "illumina", "pei")])
def test_sharded_equals_whole(world, fixture, qualtype, mode_name):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as orc

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + world * 7 + len(fixture)) % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, fixture, qualtype, mode_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    bounds, outs, total, worst = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = open(os.path.join(GOLDEN, fixture), "rb").read()
    mode = {"se": orc.MODE_SE, "pei": orc.MODE_PE_INTER}[mode_name]
    whole = orc.run(mode, orc.make_params(qualtype), data, batch_len=1 << 40)
    assert outs == whole["out"]
    for k in ("kept", "discard", "kept_p", "discard_p", "kept_s1", "kept_s2"):
        assert total[k] == whole["counters"][k]
    assert worst == float(world)
    assert bounds[0] == 0 and bounds[-1] == len(data) and bounds == sorted(bounds)
    lpu = 8 if mode_name == "pei" else 4
    for x in bounds[1:-1]:
        assert data[x - 1:x] == b"\n" and data.count(b"\n", 0, x) % lpu == 0


def test_snap_forward_cases():
    from sickle_b200 import sharding

    rec = b"@r\nACGT\n+\n@@@@\n"          # quality line starts with '@': content-based guessing would fail
    data = rec * 6
    for pos in range(len(data)):
        got = sharding.snap_forward(data, pos, data.count(b"\n", 0, pos), 4)
        want = ((pos + len(rec) - 1) // len(rec)) * len(rec)
        assert got == want, (pos, got, want)
    assert len(rec) == 15 and sharding.shard_bounds(data, 4, 4) == [0, 30, 45, 75, 90]
