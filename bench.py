#!/usr/bin/env python
"""bench.py -- throughput of the trimming hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1]): `sickle se -t sanger -q 20 -l 20` over synthetic 150 bp Sanger
reads with 3'-decaying quality (record shape R150 = 325 bytes).  One *step* = one batch of
--batch-reads reads (default 1,000,000 = 325 MB, larger than the 126 MB L2, and every step reads a
different batch).  With the default --steps 100 the timed region is the whole 100 M-read job.

value  : reads/s over the timed region with inputs and outputs resident in HBM (kernel path only,
         CUDA events on the launching stream, max over ranks).
e2e    : the same metric through the host-facing C ABI (sk_submit / sk_wait): pinned host input,
         H2D, kernels, D2H of the trimmed bytes, every step.
roofline: algorithmic bytes (FASTQ bytes in + trimmed bytes out) per step divided by
         the step's device time, against the measured HBM copy peak in MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
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
METRIC = "trimmed reads/s (sickle se, 150 bp Sanger reads, -q 20 -l 20)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes (read + write) of one kf_fused launch over the same 1 M-read batch, from the committed
    `ncu --set full` capture (profiles/r1_fused_v7_ncu_full_summary.csv); None if the file is missing."""
    import csv

    p = os.path.join(ROOT, "profiles", "r1_fused_v7_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(p)))
        h, units, r = rows[0], rows[1], rows[2]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(k)
            tot += float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        return tot
    except Exception:  # noqa: BLE001
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


def time_reference(n_reads, repeats, warmup, seed=2):
    """Run the reference CLI `se` on an n_reads sample; returns (reads/s, cores, kind, label, per-run s)."""
    from sickle_b200 import synth

    binary, kind, label = ref_binary()
    cores = os.cpu_count() or 1
    threads = cores if kind == "reference" else 1
    d = shm_dir()
    try:
        inp = os.path.join(d, "in.fastq")
        with open(inp, "wb") as f:
            done = 0
            while done < n_reads:
                m = min(250_000, n_reads - done)
                f.write(synth.fixed_length_records(m, 150, "sanger", seed=seed, start=done).tobytes())
                done += m
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
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "sickle se -t sanger -q 20 -l 20, synthetic R150 (325 B/record)",
                   "reads_per_step": n, "binary": label, "threads": cores, "files": "/dev/shm"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": "%d steps x %d reads, file to file on /dev/shm, -a %d" % (args.steps, n, cores)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fastq_gb_s": value * RECORD_BYTES / 1e9,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist

    from sickle_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    capi.load()

    B = args.batch_reads
    nbytes = B * RECORD_BYTES
    stride = (nbytes + 64 + 15) & ~15
    # distinct batches resident in HBM: one per step unless that exceeds --max-resident-gb
    nb = max(1, min(args.steps, int(args.max_resident_gb * 1e9 / 2 / stride)))
    inp = torch.zeros((nb, stride), dtype=torch.uint8, device=dev)
    out = torch.empty((nb, stride), dtype=torch.uint8, device=dev)
    for b in range(nb):
        start = (rank * nb + b) * B
        done = 0
        while done < B:  # generate in 250k-read pieces to bound temporaries
            m = min(250_000, B - done)
            rec = synth.r150_records_torch(m, start + done, dev, seed=2)
            inp[b, done * RECORD_BYTES:(done + m) * RECORD_BYTES] = rec.reshape(-1)
            done += m
    torch.cuda.synchronize()

    params = capi.make_params("sanger", 20, 20)
    ctx = capi.Context(params, nbytes + 16, 0, device=local)
    # an explicit non-default stream: its handle is passed to the library, and the timing events
    # are recorded on that same stream (handle 0 would mean "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    assert sp != 0
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()   # sampled from here to the end of the e2e leg (the timed regions are < 1 s)

    def step(b):
        ctx.trim_device(inp[b].data_ptr(), nbytes, 0, 0, [out[b].data_ptr(), 0, 0], [stride, 0, 0], sp)

    # --- pass 0 (untimed): every batch once, with its summary -> bytes out, per-stage times, launches
    out_bytes, kept, stage, launches, fused = [], 0, [0.0] * 4, 0, 0
    for b in range(nb):
        step(b)
        r = ctx.result_device(sp)
        if r.error.kind:
            raise RuntimeError("data error kind %d in synthetic batch %d" % (r.error.kind, b))
        assert r.records[0] == B, (r.records[0], B)
        out_bytes.append(r.out_bytes[0])
        kept += r.kept
        launches = r.kernel_launches
        fused += r.fused
        for k in range(4):
            stage[k] += r.stage_ms[k] / nb

    for w in range(max(args.warmup, 3)):
        step(w % nb)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for k in range(args.steps):
        step(k % nb)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    ctx.close()

    value = world * B * args.steps / (ms / 1e3)
    ms_per_step = ms / args.steps
    alg_bytes = nbytes + sum(out_bytes[k % nb] for k in range(args.steps)) / args.steps
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9

    # --- e2e: host-facing C ABI, pinned host buffers, H2D + kernels + D2H every step
    e2e = None if args.kernel_only else run_e2e(args, torch, dist, capi, inp, nb, nbytes, local, world, dev)
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "sickle se -t sanger -q 20 -l 20, synthetic R150 (325 B/record), configs[1]",
                   "reads_per_step": B, "bytes_in_per_step": nbytes, "distinct_batches_resident": nb,
                   "l2_policy": "each step reads a different 325 MB batch (> 126 MB L2)",
                   "reads_per_gpu_timed": B * args.steps, "kept_fraction": kept / (nb * B)},
        "fastq_gb_s": value * RECORD_BYTES / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic() if fused == nb else None, "peak_source": peak_src,
                     "kernel": ("kf_fused (parse+trim+route+emit, single pass) + summary, per step" if fused == nb else
                                "K1 line index + K2 trim/route + K3 emit + summary, per step"),
                     "algorithmic_bytes_per_step": alg_bytes,
                     "stage_ms": ({"kf_fused": stage[0], "summary": stage[3]} if fused == nb else
                                  {"k1_index": stage[0], "k2_trim_route": stage[1], "k3_emit": stage[2],
                                   "summary": stage[3]})},
        "e2e": e2e,
        "gpu_launches": launches * args.steps,
        "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline and not args.kernel_only:
        try:
            v, cores, kind, label, times = time_reference(args.cpu_reads, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                                    "sample": "%d reads of the same workload, file to file on /dev/shm, %s, %.1f s"
                                              % (args.cpu_reads, label, times[0])}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference",
                                    "sample": "failed: %r" % (e,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, capi, inp, nb, nbytes, local, world, dev):
    import ctypes as C

    nslots = 3
    ctx = capi.Context(capi.make_params("sanger", 20, 20), nbytes + 16, nslots, device=local)
    for s in range(nslots):
        addr = ctx.in_buffer_address(s, 0)
        host = torch.frombuffer((C.c_char * nbytes).from_address(addr), dtype=torch.uint8)
        host.copy_(inp[s % nb, :nbytes])
    torch.cuda.synchronize()

    def run(steps):
        d2h = 0
        pending = []
        for k in range(steps):
            s = k % nslots
            if len(pending) == nslots:
                r = ctx.wait(pending.pop(0))
                d2h += r.out_bytes[0]
            ctx.submit(s, 0, nbytes)
            pending.append(s)
        while pending:
            r = ctx.wait(pending.pop(0))
            d2h += r.out_bytes[0]
        return d2h

    run(max(args.warmup, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d2h = run(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ctx.close()
    return {"value": world * args.batch_reads * args.steps / dt, "unit": "reads/s",
            "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": 1e3 * dt / args.steps,
            "slots": nslots, "h2d_gb_s": nbytes * args.steps / dt / 1e9,
            "note": "sk_submit/sk_wait over pinned host buffers, 3 slots in flight; host wall clock, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-reads", type=int, default=1_000_000)
    ap.add_argument("--max-resident-gb", type=float, default=100.0)
    ap.add_argument("--cpu-reads", type=int, default=4_000_000, help="sample size of the cpu_baseline leg")
    ap.add_argument("--ref-reads", type=int, default=250_000, help="reads per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and cpu_baseline legs (for ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
