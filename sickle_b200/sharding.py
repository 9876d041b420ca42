"""Host-side multi-GPU plumbing: cut one FASTQ byte stream into per-rank shards, combine results.

Reads (pairs) are independent, so the path shards with no data-path collective:
contiguous byte ranges, one per rank.  Which of the four record lines a range boundary falls on
cannot be guessed from content ('@' and '+' are valid quality characters), so it is *counted*:
every rank counts the newlines of its raw range, an exclusive prefix sum of those world_size
integers gives the global line number at each boundary, and the boundary is moved forward to the
next line that starts a record (pair).  The only exchanges are that integer prefix and the final
counters / timings (torch.distributed all_gather / all_reduce on a few scalars; gloo on CPU, NCCL
on GPUs).  Outputs are concatenated in rank order.
"""
from __future__ import annotations


def raw_range(nbytes: int, world: int, rank: int):
    """Contiguous split of [0, nbytes) before snapping."""
    return nbytes * rank // world, nbytes * (rank + 1) // world


def snap_forward(data: bytes, pos: int, lines_before: int, lines_per_unit: int) -> int:
    """First byte >= pos that starts a unit, given that `lines_before` newlines precede `pos`.

    A unit (record: 4 lines, pair: 8) starts right after newline number k*lines_per_unit (1-based)
    or at byte 0."""
    if pos == 0:
        return 0
    # is `pos` itself a line start?  only if the previous byte is a newline
    line_no = lines_before          # index of the line containing byte `pos`
    if data[pos - 1:pos] == b"\n" and line_no % lines_per_unit == 0:
        return pos
    # otherwise skip to the end of the current line, then whole lines until a unit boundary
    p = pos
    while True:
        nl = data.find(b"\n", p)
        if nl < 0:
            return len(data)
        line_no += 1
        p = nl + 1
        if line_no % lines_per_unit == 0:
            return p


def shard_bounds(data: bytes, world: int, lines_per_unit: int, newline_counts=None):
    """Unit-aligned shard boundaries [b_0=0, b_1, ..., b_world=len(data)] (trailing partial unit stays
    in the last shard and is dropped there, as the reference does at end of file).

    `newline_counts[r]` = newlines in rank r's raw range; computed here when not given (in a real
    run every rank counts its own range and the list comes from one all_gather)."""
    n = len(data)
    if newline_counts is None:
        newline_counts = [data.count(b"\n", *raw_range(n, world, r)) for r in range(world)]
    bounds = [0]
    before = 0
    for r in range(1, world):
        before += newline_counts[r - 1]
        lo, _ = raw_range(n, world, r)
        bounds.append(max(bounds[-1], snap_forward(data, lo, before, lines_per_unit)))
    bounds.append(n)
    return bounds


def gather_newline_counts(data: bytes, world: int, rank: int, dist=None):
    """Each rank counts its raw range; one all_gather of an integer."""
    lo, hi = raw_range(len(data), world, rank)
    mine = data.count(b"\n", lo, hi)
    if dist is None or world == 1:
        return [mine]
    out = [None] * world
    dist.all_gather_object(out, mine)
    return out


def reduce_counters(counters: dict, dist=None, world: int = 1):
    """Sum per-rank counters (all_reduce of a handful of int64)."""
    if dist is None or world == 1:
        return dict(counters)
    import torch

    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64)
    dist.all_reduce(t)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def reduce_max(value: float, dist=None, world: int = 1, device=None) -> float:
    """Max over ranks of a device-measured time."""
    if dist is None or world == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
