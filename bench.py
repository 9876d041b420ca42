#!/usr/bin/env python
"""bench.py -- throughput of the trimming hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--config c2|c3|c3m|c4]   # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W                    # the reference's CPU path (oracle/_ref)

Workloads (BASELINE.json `configs`):
  c2  (default) configs[1] / configs[4]: `sickle se -t sanger -q 20 -l 20` over synthetic 150 bp Sanger reads with
      3'-decaying quality (R150 = 325 bytes per record).  ONE logical input of N*K*B reads is cut into N
      contiguous byte ranges by sickle_b200/sharding.py (newline counts -> exclusive prefix -> record phase,
      SURVEY.md 8-e); rank r trims its shard batch by batch.  At N = 1 this is configs[1]; at N > 1 it is
      configs[4] at weak scaling (K*B reads per GPU).  No data-path collective: the only exchanges are a few
      integers over gloo.
  c3  configs[2]: `sickle pe -f -r -o -p -s`, two files of R150 mates (names end in /1 and /2).
  c3m configs[2], second half: the same mates interleaved, `pe -c -M`.
  c4  configs[3]: `sickle se -t illumina -x -n`, reads of 1-20 kb (log-uniform), every 5th record with `+name`.
One *step* = one batch (default 1,000,000 reads; c4: ~255 MB of long reads); every step reads a different batch
and a batch is larger than the 126 MB L2.

value   : reads/s with inputs and outputs resident in HBM (CUDA events on the launching stream, max over
          ranks).  A pass is exactly --steps steps; passes are repeated until >= 1 s has been timed and the
          median pass is reported (min and max beside it).
e2e     : the same metric through the host-facing C ABI (sk_submit / sk_wait): pinned host input, H2D, kernels,
          D2H of the trimmed bytes, every step -- with the host-link bound measured in the same run
          (pinned cudaMemcpyAsync, one direction, both directions, all ranks at once).
roofline: algorithmic bytes (FASTQ bytes in + trimmed bytes out) per step / the step's device
          time, against the measured HBM copy peak in MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RECORD_BYTES = 325
PIECE = 250_000          # records per generator call: the logical input is the concatenation of these pieces
METRICS = {
    "c2": "trimmed reads/s (sickle se, 150 bp Sanger reads, -q 20 -l 20)",
    "c3": "trimmed reads/s (sickle pe -f -r -o -p -s, 2 x 150 bp Sanger mates, -q 20 -l 20)",
    "c3m": "trimmed reads/s (sickle pe -c -M, interleaved 150 bp Sanger mates, -q 20 -l 20)",
    "c4": "trimmed reads/s (sickle se -t illumina -x -n, 1-20 kb reads, -q 20 -l 20)",
}
WORKLOADS = {
    "c2": "sickle se -t sanger -q 20 -l 20, synthetic R150 (325 B/record), configs[1]; one input sharded by byte range over the ranks (configs[4]) when n_gpus > 1",
    "c3": "sickle pe -f -r -t sanger -o -p -s, synthetic 2 x R150 (327 B/record), configs[2]",
    "c3m": "sickle pe -c -M -t sanger, synthetic interleaved R150 (327 B/record), configs[2] -M half",
    "c4": "sickle se -t illumina -x -n, synthetic reads of 1-20 kb (log-uniform), every 5th with +name, configs[3]",
}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def lib_sha256():
    from sickle_b200 import capi

    h = hashlib.sha256()
    with open(capi.LIB_PATH, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def src_sha256():
    """sha256 over the kernel sources (path + bytes, sorted): nvcc's output is not byte-reproducible, so a library
    rebuilt from the very sources of the capture has another lib_sha256."""
    import glob

    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "sickle_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "sickle_b200", "csrc", "*.cuh")) +
                   [os.path.join(ROOT, "include", "sickle_b200.h"), os.path.join(ROOT, "Makefile")])
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode() + b"\0")
        h.update(open(f, "rb").read())
    return h.hexdigest()


def lib_is_built_from_tree():
    """The loaded library is the in-tree one and not older than any kernel source."""
    import glob

    from sickle_b200 import capi

    default = os.path.join(ROOT, "sickle_b200", "libsickle_b200.so")
    if os.path.abspath(capi.LIB_PATH) != default:
        return False
    srcs = glob.glob(os.path.join(ROOT, "sickle_b200", "csrc", "*.cu*")) + [os.path.join(ROOT, "include", "sickle_b200.h")]
    return os.path.getmtime(default) >= max(os.path.getmtime(f) for f in srcs)


def ncu_traffic(config):
    """DRAM bytes (read + write) of the dominant kernel's launch over one batch of this workload, from the
    committed `ncu --set full` capture -- only if that capture was taken from the very library that is
    loaded now (profiles/ncu_traffic.json records the sha256 of the .so next to the numbers), or from a build of the
    very kernel sources the loaded in-tree library was built from (src_sha256); else None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = d.get(config)
        if e and (e.get("lib_sha256") == lib_sha256() or
                  (e.get("src_sha256") and e.get("src_sha256") == src_sha256() and lib_is_built_from_tree())):
            return float(e["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# reference / CPU arm
# ---------------------------------------------------------------------------------------------
def ref_binary():
    """(path, kind, label).  `se -a N` of the unpatched reference races and crashes,
    so the timed binary is oracle/_ref/sickle_sync: the reference's own sources with Trim_Single's
    output made synchronous (4 lines, see oracle/Makefile)."""
    sync = os.path.join(ROOT, "oracle", "_ref", "sickle_sync")
    if os.path.exists(sync):
        return sync, "reference", "oracle/_ref/sickle_sync (reference sources, synchronous se output)"
    port = os.path.join(ROOT, "oracle", "_build", "sickle_oracle")
    if not os.path.exists(port):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return port, "port", "oracle/_build/sickle_oracle (C restatement, 1 thread)"


def shm_dir():
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    return tempfile.mkdtemp(prefix="sickle_bench_", dir=d)


def write_r150_file(path, n_reads, seed=2):
    """The c2 workload as a file.  Generated on the GPU when there is one (the numpy generator makes
    ~80 k reads/s: 10 M reads would take two minutes), piece by piece like the resident input."""
    from sickle_b200 import synth

    dev = None
    try:
        import torch

        if torch.cuda.is_available():
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:  # noqa: BLE001
        dev = None
    with open(path, "wb") as f:
        done = 0
        while done < n_reads:
            m = min(PIECE, n_reads - done)
            if dev is not None:
                f.write(synth.r150_records_torch(m, done, dev, seed=seed).cpu().numpy().tobytes())
            else:
                f.write(synth.fixed_length_records(m, 150, "sanger", seed=seed, start=done).tobytes())
            done += m
    return "torch" if dev is not None else "numpy"


def time_reference(n_reads, repeats, warmup, seed=2):
    """Run the reference CLI `se` on an n_reads sample; returns (reads/s, cores, kind, label, per-run s)."""
    binary, kind, label = ref_binary()
    cores = os.cpu_count() or 1
    threads = cores if kind == "reference" else 1
    d = shm_dir()
    try:
        inp = os.path.join(d, "in.fastq")
        write_r150_file(inp, n_reads, seed)
        times = []
        for it in range(warmup + repeats):
            out = os.path.join(d, "out.fastq")
            t0 = time.perf_counter()
            rc = subprocess.run([binary, "se", "-f", inp, "-t", "sanger", "-o", out, "-q", "20", "-l", "20",
                                 "-a", str(threads), "-b", "512"], stdout=subprocess.DEVNULL,
                                stderr=subprocess.DEVNULL, timeout=3600).returncode
            dt = time.perf_counter() - t0
            if rc != 0:
                raise RuntimeError("reference exited %d" % rc)
            if it >= warmup:
                times.append(dt)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    total = sum(times)
    return n_reads * len(times) / total, threads, kind, label, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_reads
    value, cores, kind, label, times = time_reference(n, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": METRICS["c2"], "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"], "reads_per_step": n, "binary": label, "threads": cores, "files": "/dev/shm",
                   "note": "each step is one whole run of the reference program (process start, file read, trim, file write) "
                           "on a %d-read sample of the workload" % n},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": "%d steps x %d reads, file to file on /dev/shm, -a %d" % (args.steps, n, cores)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fastq_gb_s": value * RECORD_BYTES / 1e9,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def pin_to_gpu(local):
    """Keep this rank's threads -- and with them its pinned allocations -- on the CPUs next to its GPU
    (NVML's ideal affinity).  BENCH_NO_AFFINITY=1 switches it off for A/B runs."""
    if os.environ.get("BENCH_NO_AFFINITY"):
        return {"set": False, "why": "BENCH_NO_AFFINITY"}
    before = len(os.sched_getaffinity(0))
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return {"set": True, "how": "nvmlDeviceSetCpuAffinity", "cpus_before": before, "cpus": len(after),
                "first": after[0] if after else None, "last": after[-1] if after else None}
    except Exception as e:  # noqa: BLE001
        return {"set": False, "why": repr(e)[:120], "cpus": before}


class Group:
    """The few integers the ranks exchange (gloo over CPU tensors; no collective touches the data path)."""

    def __init__(self, world):
        self.world = world
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("gloo")
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def max_list(self, xs):
        if not self.dist:
            return list(xs)
        import torch

        t = torch.tensor(list(xs), dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_list(self, xs):
        if not self.dist:
            return list(xs)
        import torch

        t = torch.tensor(list(xs), dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def gather(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def r150_range(torch, synth, dev, first, count, seed=2, suffix=b""):
    """Records [first, first + count) of the logical R150 input, as a flat uint8 device tensor."""
    w = RECORD_BYTES + len(suffix)
    out = torch.empty(count * w, dtype=torch.uint8, device=dev)
    done = 0
    while done < count:
        p = (first + done) // PIECE
        lo = first + done - p * PIECE
        m = min(PIECE - lo, count - done)
        rec = synth.r150_records_torch(PIECE, p * PIECE, dev, seed=seed, suffix=suffix)
        out[done * w:(done + m) * w] = rec[lo:lo + m].reshape(-1)
        del rec
        done += m
    return out


def plan_shard(torch, synth, sharding, grp, dev, rank, world, total_records, lines_per_unit=4):
    """Cut the logical input (total_records R150 records) into `world` contiguous byte ranges and return this
    rank's unit-aligned range [b_lo, b_hi) as (first_record, n_records, split_info).  The cut is the product's:
    raw byte ranges, newline counts of the raw ranges (counted on the device), an exclusive prefix over the
    ranks, and a forward snap to the next line that starts a record (sickle_b200/sharding.py)."""
    t0 = time.perf_counter()
    nbytes = total_records * RECORD_BYTES
    lo, hi = sharding.raw_range(nbytes, world, rank)
    # newlines of the raw range, counted on the device piece by piece (a shard is tens of GB: no whole-range temporaries)
    mine, window = 0, b""
    r_first, r_end = lo // RECORD_BYTES, min(total_records, -(-hi // RECORD_BYTES))
    r = r_first
    while r < r_end:
        n = min(PIECE - r % PIECE, r_end - r)
        buf = r150_range(torch, synth, dev, r, n)
        off = r * RECORD_BYTES
        a, b = max(lo, off) - off, min(hi, off + n * RECORD_BYTES) - off
        mine += int(torch.count_nonzero(buf[a:b] == 10).item())
        if r == r_first and rank:
            # the byte in front of the boundary and a few records after it are enough to move it to the next record start
            window = bytes(buf[lo - off - 1:min(lo - off + 4096, buf.numel())].cpu().numpy().tobytes()) if lo > off else b""
        del buf
        r += n
    if rank and not window:          # the raw boundary is the first byte of a piece: fetch the byte before it as well
        buf = r150_range(torch, synth, dev, r_first - 1, 14)
        window = bytes(buf[RECORD_BYTES - 1:RECORD_BYTES - 1 + 4097].cpu().numpy().tobytes())
        del buf
    counts = grp.gather(mine)
    before = sum(counts[:rank])
    start = lo - 1 + sharding.snap_forward(window, 1, before, lines_per_unit) if rank else 0
    starts = grp.gather(start) + [nbytes]
    b_lo, b_hi = starts[rank], starts[rank + 1]
    assert b_lo % RECORD_BYTES == 0 and b_hi % RECORD_BYTES == 0, (b_lo, b_hi)       # (fixed-size records: a check of the snap)
    return b_lo // RECORD_BYTES, (b_hi - b_lo) // RECORD_BYTES, {
        "raw_range": [lo, hi], "newlines_in_raw_range": mine, "lines_before": before, "snapped_range": [b_lo, b_hi],
        "plan_s": round(time.perf_counter() - t0, 3)}


def make_batches(args, torch, synth, dev, first_record, nb):
    """Device-resident input batches of this rank: list of (in0, n0, in1, n1, records)."""
    B = args.batch_reads
    out = []
    if args.config == "c2":
        data = r150_range(torch, synth, dev, first_record, nb * B)
        for b in range(nb):
            out.append((data[b * B * RECORD_BYTES:(b + 1) * B * RECORD_BYTES], B * RECORD_BYTES, None, 0, B))
        return out, data
    if args.config in ("c3", "c3m"):
        w = RECORD_BYTES + 2
        half = B // 2
        f = r150_range(torch, synth, dev, first_record, nb * half, seed=3, suffix=b"/1")
        r = r150_range(torch, synth, dev, first_record, nb * half, seed=3 + 7919, suffix=b"/2")
        if args.config == "c3":
            for b in range(nb):
                out.append((f[b * half * w:(b + 1) * half * w], half * w, r[b * half * w:(b + 1) * half * w], half * w, 2 * half))
            return out, (f, r)
        inter = torch.empty(2 * f.numel(), dtype=torch.uint8, device=dev).view(nb * half, 2, w)
        inter[:, 0] = f.view(nb * half, w)
        inter[:, 1] = r.view(nb * half, w)
        inter = inter.view(-1)
        del f, r
        for b in range(nb):
            out.append((inter[b * 2 * half * w:(b + 1) * 2 * half * w], 2 * half * w, None, 0, 2 * half))
        return out, inter
    # c4: a 4,000-read file of 1-20 kb reads per distinct batch, repeated to ~255 MB
    import numpy as np

    keep = []
    for b in range(nb):
        v = synth.variable_length_records(4000, 1000, 20000, "illumina", 70 + b % 4)
        arr = np.frombuffer(v, dtype=np.uint8)
        rep = max(1, args.c4_batch_bytes // arr.size)
        buf = torch.zeros(((arr.size * rep + 64 + 15) & ~15), dtype=torch.uint8, device=dev)
        buf[:arr.size * rep] = torch.from_numpy(arr.copy()).to(dev).repeat(rep)
        keep.append(buf)
        out.append((buf, arr.size * rep, None, 0, 4000 * rep))
    return out, keep


def make_params(args, capi):
    if args.config == "c2":
        return capi.make_params("sanger", 20, 20)
    if args.config == "c3":
        return capi.make_params("sanger", 20, 20, mode=capi.MODE_PE_2FILE, has_singles=True)
    if args.config == "c3m":
        return capi.make_params("sanger", 20, 20, mode=capi.MODE_PE_INTER_M, has_singles=False)
    return capi.make_params("illumina", 20, 20, x=True, n=True)


def link_probe(torch, grp, rank, world, nbytes=256 << 20, reps=3):
    """Pinned cudaMemcpyAsync bandwidth of this rank's host link, GB/s per direction: one direction at a time and
    both at once, this rank alone and all ranks together (SURVEY.md 8-d asks for the bound in the same run)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(2)]
        torch.cuda.synchronize()
        if h2d:
            with torch.cuda.stream(s1):
                ev[0][0].record()
                for _ in range(reps):
                    d_a.copy_(h_in, non_blocking=True)
                ev[0][1].record()
        if d2h:
            with torch.cuda.stream(s2):
                ev[1][0].record()
                for _ in range(reps):
                    h_out.copy_(d_b, non_blocking=True)
                ev[1][1].record()
        torch.cuda.synchronize()
        gb = reps * nbytes / 1e6
        return (round(gb / ev[0][0].elapsed_time(ev[0][1]), 1) if h2d else None,
                round(gb / ev[1][0].elapsed_time(ev[1][1]), 1) if d2h else None)

    run(True, True)
    res = {}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        for who in (list(range(world)) if world > 1 else []) + [-1]:       # one rank at a time, then all ranks together
            grp.barrier()
            if who in (-1, rank):
                a, b = run(h2d, d2h)
                key = "%s_%s" % (name, "all" if who < 0 else "solo")
                res[key] = [a, b] if name == "both" else (a if h2d else b)
            grp.barrier()
    if world == 1:
        for name in ("h2d", "d2h", "both"):
            res[name + "_solo"] = res[name + "_all"]
    del h_in, h_out, d_a, d_b
    return res


def run_cuda_arm(args):
    import torch

    from sickle_b200 import capi, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    affinity = pin_to_gpu(local)                      # before anything allocates pinned memory
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    grp = Group(world)
    capi.load()

    B = args.batch_reads
    K = args.steps
    W = max(args.warmup, 3)
    params = make_params(args, capi)
    # distinct batches resident in HBM: one per step unless that exceeds --max-resident-gb
    if args.config == "c4":
        step_in_bytes = args.c4_batch_bytes
        nb = max(1, min(K, 8))
    else:
        step_in_bytes = B * (RECORD_BYTES if args.config == "c2" else RECORD_BYTES + 2)
        nb = max(1, min(K, int(args.max_resident_gb * 1e9 / 2.2 / step_in_bytes)))
    split = None
    first = 0
    if args.config == "c2":
        # one logical input, cut by byte range over the ranks (3 extra records so that no raw boundary is a record boundary)
        first, nrec, split = plan_shard(torch, synth, sharding, grp, dev, rank, world, world * nb * B + 3)
        assert nrec >= nb * B, (nrec, nb, B)
    else:
        first = rank * nb * B
    batches, keepalive = make_batches(args, torch, synth, dev, first, nb)
    cap = ((max(b[1] + b[3] for b in batches) + 64 + 15) & ~15)
    outs = [torch.empty((nb, cap), dtype=torch.uint8, device=dev) for _ in range(3 if args.config == "c3" else 1)]
    torch.cuda.synchronize()

    slot = max(max(b[1], b[3]) for b in batches) + 16
    ctx = capi.Context(params, slot, 0, device=local)
    # an explicit non-default stream: its handle is passed to the library, and the timing events
    # are recorded on that same stream (handle 0 would mean "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    assert sp != 0
    clocks = ClockSampler(local)

    def step(b):
        i0, n0, i1, n1, _ = batches[b]
        if args.config == "c3":
            ptrs, caps = [outs[0][b].data_ptr(), outs[1][b].data_ptr(), outs[2][b].data_ptr()], [cap, cap, cap]
        else:
            ptrs, caps = [outs[0][b].data_ptr(), 0, 0], [cap, 0, 0]
        ctx.trim_device(i0.data_ptr(), n0, i1.data_ptr() if i1 is not None else 0, n1, ptrs, caps, sp)

    # --- pass 0 (untimed): every batch once, with its summary -> bytes out, per-stage times, launches
    out_bytes, kept, stage, launches, fused, recs = [], 0, [0.0] * 4, 0, 0, 0
    for b in range(nb):
        step(b)
        r = ctx.result_device(sp)
        if r.error.kind:
            raise RuntimeError("data error kind %d in synthetic batch %d" % (r.error.kind, b))
        assert r.records[0] + r.records[1] == batches[b][4], (r.records[0], r.records[1], batches[b][4])
        out_bytes.append(sum(r.out_bytes[k] for k in range(3)))
        kept += r.kept + r.kept_p + r.kept_s1 + r.kept_s2
        launches = r.kernel_launches
        fused += r.fused
        recs += batches[b][4]
        for k in range(4):
            stage[k] += r.stage_ms[k] / nb
    reads_per_step = recs / nb

    def one_pass():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K):
            step(k % nb)
        e1.record(stream)
        return e0, e1

    for w in range(W):
        step(w % nb)
    torch.cuda.synchronize()
    # calibration: how many K-step passes make >= --min-timed-s of timed region (same count on every rank)
    e0, e1 = one_pass()
    torch.cuda.synchronize()
    cal_ms = grp.max_list([e0.elapsed_time(e1)])[0]
    npass = int(min(args.max_passes, max(1, math.ceil(args.min_timed_s * 1e3 / max(cal_ms, 1e-3)))))
    if rank == 0:
        clocks.start()
    grp.barrier()
    torch.cuda.synchronize()
    evs = [one_pass() for _ in range(npass)]
    torch.cuda.synchronize()
    grp.barrier()
    pass_ms = grp.max_list([a.elapsed_time(b) for a, b in evs])          # per pass, max over ranks
    clk = clocks.stop() if rank == 0 else None
    # per-kernel times of a step (library events), once more now that everything is warm
    stage = [0.0] * 4
    for b in range(nb):
        step(b)
        r = ctx.result_device(sp)
        for k in range(4):
            stage[k] += r.stage_ms[k] / nb
    ctx.close()

    ms = statistics.median(pass_ms)
    value = world * reads_per_step * K / (ms / 1e3)
    ms_per_step = ms / K
    in_bytes = sum(b[1] + b[3] for b in batches) / nb
    alg_bytes = in_bytes + sum(out_bytes[k % nb] for k in range(K)) / K
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
    all_fused = fused == nb

    # --- c2 at N > 1: the sharded run writes the same bytes as one GPU (checked on a small input, every run)
    shard_check = sharded_equals_whole(args, torch, synth, sharding, capi, grp, dev, rank, world, local) if (args.config == "c2" and world > 1) else None

    # --- e2e: host-facing C ABI, pinned host buffers, H2D + kernels + D2H every step
    e2e = None
    if not args.kernel_only:
        e2e = run_e2e(args, torch, capi, grp, batches, params, slot, local, world, reads_per_step)
        e2e["affinity"] = affinity

    if rank != 0:
        grp.close()
        return
    if all_fused:
        kernel = "kf_fused (parse+trim+route+emit, single pass) + summary, per step"
        stage_ms = {"kf_fused": stage[0], "summary": stage[3]}
    else:
        kernel = "K1 line index + K2 trim/route + K3 emit + summary, per step"
        stage_ms = {"k1_index": stage[0], "k2_trim_route": stage[1], "k3_emit": stage[2], "summary": stage[3]}
    line = {
        "metric": METRICS[args.config], "value": value, "unit": "reads/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "config_id": args.config,
                   "reads_per_step": reads_per_step, "bytes_in_per_step": in_bytes, "distinct_batches_resident": nb,
                   "l2_policy": "each step reads a different batch of %.0f MB (> 126 MB L2)" % (in_bytes / 1e6),
                   "reads_per_gpu_per_pass": reads_per_step * K, "kept_fraction": kept / max(recs, 1),
                   "split": split, "sharded_equals_whole": shard_check},
        "timing": {"passes": npass, "steps_per_pass": K, "timed_region_s": sum(pass_ms) / 1e3,
                   "ms_per_step_min": min(pass_ms) / K, "ms_per_step_median": ms_per_step, "ms_per_step_max": max(pass_ms) / K,
                   "how": "CUDA events on the launching stream around each pass of exactly `steps` steps, max over ranks per pass; value = median pass"},
        "fastq_gb_s": value * in_bytes / max(reads_per_step, 1) / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.config) if all_fused or args.config in ("c3", "c4") else None,
                     "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_step": alg_bytes, "stage_ms": stage_ms},
        "e2e": e2e,
        "gpu_launches": launches * K * npass,
        "clocks": clk,
        "collective": None,
    }
    if world == 1 and args.config == "c2" and not args.no_cpu_baseline and not args.kernel_only:
        try:
            v, cores, kind, label, times = time_reference(args.cpu_reads, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                                    "sample": "%d reads of the same workload, file to file on /dev/shm, %s, %.1f s"
                                              % (args.cpu_reads, label, times[0])}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference",
                                    "sample": "failed: %r" % (e,)}
    print(json.dumps(line), flush=True)
    grp.close()


def sharded_equals_whole(args, torch, synth, sharding, capi, grp, dev, rank, world, local, total=1_000_003):
    """configs[4]'s correctness property, checked in every multi-GPU run: a small logical input cut over the
    ranks exactly like the timed one, every shard trimmed on its GPU, the outputs concatenated in rank order
    (through files on /dev/shm) and compared with one GPU trimming the whole input."""
    first, nrec, _ = plan_shard(torch, synth, sharding, grp, dev, rank, world, total)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    tag = os.environ.get("MASTER_PORT", "0")

    def trim(first_, n_):
        data = r150_range(torch, synth, dev, first_, n_)
        out = torch.empty(data.numel() + 64, dtype=torch.uint8, device=dev)
        with capi.Context(capi.make_params("sanger", 20, 20), data.numel() + 16, 0, device=local) as c:
            c.trim_device(data.data_ptr(), data.numel(), 0, 0, [out.data_ptr(), 0, 0], [out.numel(), 0, 0], None)
            r = c.result_device(None)
        assert r.error.kind == 0 and r.records[0] == n_
        return out[:r.out_bytes[0]].cpu().numpy().tobytes(), r.kept

    mine, kept = trim(first, nrec)
    path = os.path.join(d, "sickle_bench_%s_shard%d.bin" % (tag, rank))
    with open(path, "wb") as f:
        f.write(mine)
    grp.barrier()
    res = None
    if rank == 0:
        h = hashlib.md5()
        n = 0
        for r_ in range(world):
            b = open(os.path.join(d, "sickle_bench_%s_shard%d.bin" % (tag, r_)), "rb").read()
            h.update(b)
            n += len(b)
        whole, kept_all = trim(0, total)
        res = {"reads": total, "ranks": world, "bytes": n, "md5_sharded": h.hexdigest(), "md5_one_gpu": hashlib.md5(whole).hexdigest()}
        res["equal"] = res["md5_sharded"] == res["md5_one_gpu"] and n == len(whole)
    grp.barrier()
    os.unlink(path)
    return res


def run_e2e(args, torch, capi, grp, batches, params, slot, local, world, reads_per_step):
    import ctypes as C

    probe = link_probe(torch, grp, int(os.environ.get("RANK", "0")), world)
    nslots = args.e2e_slots
    ctx = capi.Context(params, slot, nslots, device=local)
    n_in = 2 if batches[0][2] is not None else 1
    for s in range(nslots):
        i0, n0, i1, n1, _ = batches[s % len(batches)]
        for which, (t, n) in enumerate(((i0, n0), (i1, n1))[:n_in]):
            addr = ctx.in_buffer_address(s, which)
            host = torch.frombuffer((C.c_char * n).from_address(addr), dtype=torch.uint8)
            host.copy_(t[:n])
    torch.cuda.synchronize()
    sizes = [(batches[s % len(batches)][1], batches[s % len(batches)][3]) for s in range(nslots)]

    def run(steps):
        d2h = 0
        pending = []
        for k in range(steps):
            s = k % nslots
            if len(pending) == nslots:
                r = ctx.wait(pending.pop(0))
                d2h += sum(r.out_bytes[j] for j in range(3))
            ctx.submit(s, 0, sizes[s][0], 0, sizes[s][1])
            pending.append(s)
        while pending:
            r = ctx.wait(pending.pop(0))
            d2h += sum(r.out_bytes[j] for j in range(3))
        return d2h

    K = args.steps
    run(max(args.warmup, 3))
    t0 = time.perf_counter()
    run(K)
    cal = grp.max_list([time.perf_counter() - t0])[0]
    npass = int(min(50, max(1, math.ceil(args.min_timed_s / max(cal, 1e-6)))))
    dts, d2h = [], 0
    for _ in range(npass):
        grp.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = run(K)
        torch.cuda.synchronize()
        dts.append(time.perf_counter() - t0)
    dts = grp.max_list(dts)
    ctx.close()
    dt = statistics.median(dts)
    h2d = sum(sizes[k % nslots][0] + sizes[k % nslots][1] for k in range(K)) / K
    value = world * reads_per_step * K / dt
    # host-link bound at this N: every rank copies in both directions at once
    allp = grp.gather(probe)
    both = [p["both_all"] for p in allp]
    h2d_all = statistics.mean(b[0] for b in both)
    d2h_all = statistics.mean(b[1] for b in both)
    in_per_read, out_per_read = h2d / reads_per_step, (d2h / K) / reads_per_step
    bound = world * min(h2d_all * 1e9 / in_per_read, d2h_all * 1e9 / max(out_per_read, 1e-9))
    return {"value": value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h / K,
            "ms_per_step": 1e3 * dt / K, "passes": npass, "ms_per_step_min": 1e3 * min(dts) / K, "ms_per_step_max": 1e3 * max(dts) / K,
            "slots": nslots, "h2d_gb_s_per_gpu": h2d * K / dt / 1e9, "d2h_gb_s_per_gpu": d2h / dt / 1e9,
            "bound": {"unit": "GB/s per direction per GPU (pinned cudaMemcpyAsync, 256 MiB, same run)",
                      "h2d_solo": probe["h2d_solo"], "d2h_solo": probe["d2h_solo"], "both_solo": probe["both_solo"],
                      "both_all_mean": [round(h2d_all, 1), round(d2h_all, 1)], "both_all_per_rank": both,
                      "reads_per_s": bound,
                      "how": "n_gpus x min(H2D rate / input bytes per read, D2H rate / output bytes per read), rates with every rank copying both ways at once"},
            "frac_of_bound": value / bound,
            "write_combined_input": bool(os.environ.get("SICKLE_B200_WC_INPUT")),
            "note": "sk_submit/sk_wait over pinned host buffers, %d slots in flight; host wall clock, max over ranks, median pass" % nslots}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c3m", "c4", "c5"],
                    help="BASELINE.json workload (c5 = c2: at n_gpus > 1 the c2 input is one input sharded over the ranks)")
    ap.add_argument("--batch-reads", type=int, default=1_000_000)
    ap.add_argument("--c4-batch-bytes", type=int, default=255_000_000)
    ap.add_argument("--max-resident-gb", type=float, default=100.0)
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="passes of `steps` steps are repeated until this much is timed")
    ap.add_argument("--max-passes", type=int, default=400)
    ap.add_argument("--e2e-slots", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=10_000_000, help="sample size of the cpu_baseline leg")
    ap.add_argument("--ref-reads", type=int, default=2_000_000, help="reads per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and cpu_baseline legs (for ncu runs)")
    args = ap.parse_args()
    if args.config == "c5":
        args.config = "c2"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
