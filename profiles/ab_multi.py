"""Same-GPU A/B of several builds of the library in ONE process (kernel path only, device-resident input).

    python profiles/ab_multi.py [--rounds R] [--steps K] [--workload se|pe|pem|a8] libA.so libB.so[@ENV=VALUE,...] ...

The synthetic batches (bench.py's configs[1] workload unless --workload says otherwise) are generated once;
every library gets its own context (its own dlopen handle, so each build keeps its own kernels) and the
libraries are timed in alternation, R rounds of K steps each, CUDA events on the launching stream.
Prints one JSON line per library: median / min ms per 1 M-read step and the fraction of the measured
HBM peak on algorithmic bytes.  A knock-out build (SK_KO_*) writes wrong bytes by construction: its
line says what the phase costs, nothing else.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sickle_b200 import capi, synth  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


class Lib:
    """capi.Context against an explicit .so (capi.load() caches one library per process)."""

    def __init__(self, spec, params, slot_bytes):
        # "lib.so@NAME=VALUE[,NAME=VALUE]": environment the library reads when the context is created
        path, _, envs = spec.partition("@")
        self.path = spec
        saved = capi._lib, capi.LIB_PATH
        capi._lib, capi.LIB_PATH = None, os.path.abspath(path)
        old = {}
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            old[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            self.ctx = capi.Context(params, slot_bytes, 0, device=0)
        finally:
            capi._lib, capi.LIB_PATH = saved
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--workload", default="se", choices=["se", "pe", "pem", "a8"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, RB = args.reads, 325
    nbytes = B * RB
    stride = (nbytes + 64 + 15) & ~15
    inp = torch.zeros((args.batches, stride), dtype=torch.uint8, device=dev)
    out = torch.empty((args.batches, stride), dtype=torch.uint8, device=dev)
    out2 = torch.empty(stride, dtype=torch.uint8, device=dev)
    for b in range(args.batches):
        done = 0
        while done < B:
            m = min(250_000, B - done)
            inp[b, done * RB:(done + m) * RB] = synth.r150_records_torch(m, b * B + done, dev, seed=2).reshape(-1)
            done += m
    torch.cuda.synchronize()
    mode = {"se": capi.MODE_SE, "a8": capi.MODE_SE, "pe": capi.MODE_PE_INTER, "pem": capi.MODE_PE_INTER_M}[args.workload]
    params = capi.make_params("sanger", 20, 20, mode=mode, emulate_threads=8 if args.workload == "a8" else 1)
    libs = [Lib(p, params, nbytes + 16) for p in args.libs]
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    res = {l.path: [] for l in libs}
    info = {}
    for l in libs:   # pass 0: bytes out, fused or not
        ob = []
        for b in range(args.batches):
            l.ctx.trim_device(inp[b].data_ptr(), nbytes, 0, 0, [out[b].data_ptr(), 0, out2.data_ptr()], [stride, 0, stride], sp)
            try:
                r = l.ctx.result_device(sp)
                ob.append(r.out_bytes[0] + r.out_bytes[2])
            except capi.SickleError as e:   # a knock-out build may trip a capacity check: it is timed all the same
                r = None
                err = str(e)
        info[l.path] = ({"out_bytes": statistics.mean(ob), "fused": r.fused, "launches": r.kernel_launches} if r is not None else
                        {"out_bytes": 0, "fused": -1, "launches": -1, "error": err})
    for rd in range(args.rounds):
        for l in libs:
            for w in range(3):
                l.ctx.trim_device(inp[w].data_ptr(), nbytes, 0, 0, [out[w].data_ptr(), 0, out2.data_ptr()], [stride, 0, stride], sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for k in range(args.steps):
                b = k % args.batches
                l.ctx.trim_device(inp[b].data_ptr(), nbytes, 0, 0, [out[b].data_ptr(), 0, out2.data_ptr()], [stride, 0, stride], sp)
            e1.record(stream)
            torch.cuda.synchronize()
            res[l.path].append(e0.elapsed_time(e1) / args.steps)
            try:
                l.ctx.result_device(sp)
            except capi.SickleError:
                pass
    pk = peak()
    base_alg = nbytes + info[libs[0].path]["out_bytes"]
    for l in libs:
        ms = res[l.path]
        med = statistics.median(ms)
        print(json.dumps({"lib": os.path.basename(l.path), "workload": args.workload, "ms_median": round(med, 4), "ms_min": round(min(ms), 4),
                          "frac_of_peak_on_shipped_bytes": round(base_alg / med / 1e6 / pk, 4),
                          "fused": info[l.path]["fused"], "launches": info[l.path]["launches"],
                          "out_bytes": int(info[l.path]["out_bytes"]), "error": info[l.path].get("error"), "env_ch": os.environ.get("SICKLE_B200_FUSED_CH")}), flush=True)
    for l in libs:
        l.ctx.close()


if __name__ == "__main__":
    main()
