"""Kernel-path throughput of the other rows of the hot path (device-resident input, library events):
read lengths 50-250, interleaved and two-file paired end, the general (K1/K2/K3) path and `-a N`
order emulation.  bench.py measures the headline configuration only; this is the table behind
DESIGN.md section 6.   python profiles/workloads.py > profiles/r1_workloads_v7.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sickle_b200 import capi, synth  # noqa: E402

def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


def device_bytes(mat, target_bytes, dev):
    t = torch.from_numpy(np.ascontiguousarray(mat)).to(dev).reshape(-1)
    rep = max(1, target_bytes // t.numel())
    buf = torch.zeros(t.numel() * rep + 64, dtype=torch.uint8, device=dev)
    buf[:t.numel() * rep] = t.repeat(rep)
    return buf, t.numel() * rep, mat.shape[0] * rep


def measure(name, mode, inputs, n_records, env=None, emulate_threads=1, steps=10, qualtype="sanger", x=False, n=False):
    dev = torch.device("cuda:0")
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        p = capi.make_params(qualtype, 20, 20, x, n, mode=mode, emulate_threads=emulate_threads, has_singles=True)
        slot = max(n for _, n in inputs) + 16
        ctx = capi.Context(p, slot, 0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    st = torch.cuda.Stream()
    outs = [torch.empty(slot + 64, dtype=torch.uint8, device=dev) for _ in range(3)]
    caps = [slot + 64, slot + 64, 2 * slot + 64]
    outs[2] = torch.empty(caps[2], dtype=torch.uint8, device=dev)
    in1 = inputs[1] if len(inputs) > 1 else (None, 0)
    ms, fused, res = [], 0, None
    for it in range(3 + steps):
        ctx.trim_device(inputs[0][0].data_ptr(), inputs[0][1], in1[0].data_ptr() if in1[0] is not None else 0, in1[1],
                        [o.data_ptr() for o in outs], caps, st.cuda_stream)
        res = ctx.result_device(st.cuda_stream)
        assert res.error.kind == 0
        if it >= 3:
            ms.append(res.kernel_ms)
            fused += res.fused
    ctx.close()
    t = float(np.median(ms))
    bytes_in = sum(n for _, n in inputs)
    bytes_out = sum(res.out_bytes[k] for k in range(3))
    gbs = (bytes_in + bytes_out) / t / 1e6
    print(json.dumps({"workload": name, "records": n_records, "ms": round(t, 4), "reads_per_s": round(n_records / t * 1e3),
                      "fused_batches": "%d/%d" % (fused, steps), "launches": res.kernel_launches,
                      "stage_ms": [round(x, 4) for x in res.stage_ms],
                      "bytes_in": bytes_in, "bytes_out": int(bytes_out), "algorithmic_GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / hbm_peak(), 4)}), flush=True)


def main():
    dev = torch.device("cuda:0")
    target = 300_000_000
    if len(sys.argv) > 1 and sys.argv[1] == "--general-only":   # short run for profiling K1/K2/K3 under ncu
        m = synth.fixed_length_records(200_000, 150, "sanger", seed=190)
        buf, n, recs = device_bytes(m, target, dev)
        measure("se R150, general path (K1/K2/K3)", capi.MODE_SE, [(buf, n)], recs, env={"SICKLE_B200_PATH": "general"}, steps=3)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--a8-only":        # short run for profiling the -a N passes under ncu
        m = synth.fixed_length_records(200_000, 150, "sanger", seed=190)
        buf, n, recs = device_bytes(m, target, dev)
        measure("se R150, -a 8 reference order (index pass + ordered emit)", capi.MODE_SE, [(buf, n)], recs, emulate_threads=8, steps=3)
        measure("pe interleaved R150, -a 8 (index pass + routing + K3)", capi.MODE_PE_INTER, [(buf, n)], recs, emulate_threads=8, steps=3)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--c4-only":        # short run for profiling the long-read kernels under ncu
        v = synth.variable_length_records(4000, 1000, 20000, "illumina", 70)
        arr = np.frombuffer(v, dtype=np.uint8)
        rep = max(1, target // arr.size)
        buf = torch.zeros(arr.size * rep + 64, dtype=torch.uint8, device=dev)
        buf[:arr.size * rep] = torch.from_numpy(arr.copy()).to(dev).repeat(rep)
        flags = dict(x="-x" in sys.argv, n="-n" in sys.argv)
        measure("se long reads 1-20 kb, illumina, %s (configs[3])" % flags, capi.MODE_SE, [(buf, arr.size * rep)], 4000 * rep,
                qualtype="illumina", steps=3, **flags)
        return
    for L in (50, 75, 100, 150, 250):
        m = synth.fixed_length_records(200_000, L, "sanger", seed=40 + L)
        buf, n, recs = device_bytes(m, target, dev)
        measure("se R%d" % L, capi.MODE_SE, [(buf, n)], recs)
        if L == 150:
            measure("se R150, general path (K1/K2/K3)", capi.MODE_SE, [(buf, n)], recs, env={"SICKLE_B200_PATH": "general"})
            measure("se R150, -a 8 reference order (index pass + ordered emit)", capi.MODE_SE, [(buf, n)], recs, emulate_threads=8)
            measure("se R150, -a 64 reference order (index pass + routing + K3)", capi.MODE_SE, [(buf, n)], recs, emulate_threads=64)
            measure("se R150, -a 8, K1/K2/K3 only", capi.MODE_SE, [(buf, n)], recs, emulate_threads=8, env={"SICKLE_B200_PATH": "general"})
        del buf
    f, r, inter = synth.paired_records(100_000, 150, "sanger", seed=50)
    buf, n, recs = device_bytes(inter, target, dev)
    measure("pe interleaved R150 (-c, -m, -s)", capi.MODE_PE_INTER, [(buf, n)], recs)
    measure("pe interleaved R150, -M", capi.MODE_PE_INTER_M, [(buf, n)], recs)
    del buf
    b0, n0, r0 = device_bytes(f, target // 2, dev)
    b1, n1, r1 = device_bytes(r, target // 2, dev)
    measure("pe two files R150 (-f -r, -o -p -s; two passes of the single-pass kernel)", capi.MODE_PE_2FILE, [(b0, n0), (b1, n1)], r0 + r1)
    # variable-length reads (adapter-trimmed input), 36..151 bases
    v = synth.variable_length_records(150_000, 36, 151, "sanger", 60)
    vb = v if isinstance(v, (bytes, bytearray)) else v.tobytes()
    arr = np.frombuffer(vb, dtype=np.uint8)
    rep = max(1, target // arr.size)
    buf = torch.zeros(arr.size * rep + 64, dtype=torch.uint8, device=dev)
    buf[:arr.size * rep] = torch.from_numpy(arr.copy()).to(dev).repeat(rep)
    measure("se variable length 36-151", capi.MODE_SE, [(buf, arr.size * rep)], 150_000 * rep)
    del buf
    # BASELINE.json configs[3]: reads of 1-20 kb, -x -n, Illumina / Solexa encodings (general path: a warp per read)
    for qualtype, seed in (("illumina", 70), ("solexa", 71)):
        v = synth.variable_length_records(4000, 1000, 20000, qualtype, seed)
        arr = np.frombuffer(v, dtype=np.uint8)
        rep = max(1, target // arr.size)
        buf = torch.zeros(arr.size * rep + 64, dtype=torch.uint8, device=dev)
        buf[:arr.size * rep] = torch.from_numpy(arr.copy()).to(dev).repeat(rep)
        measure("se long reads 1-20 kb, %s, -x -n (configs[3])" % qualtype, capi.MODE_SE, [(buf, arr.size * rep)], 4000 * rep,
                qualtype=qualtype, x=True, n=True)
        del buf


if __name__ == "__main__":
    main()
