#!/usr/bin/env python
"""Debug helper: per-phase cycle split of kf_fused (needs the -DSK_PHASE_TIMING build:
nvcc ... -DSK_PHASE_TIMING -shared sickle_b200/csrc/capi.cu -o sickle_b200/libsickle_b200_timing.so)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sickle_b200 import capi, synth  # noqa: E402

capi.LIB_PATH = os.path.join(ROOT, "sickle_b200", "libsickle_b200_timing.so")
lib = capi.load()
dev = torch.device("cuda:0")
B = 1_000_000
rec = synth.r150_records_torch(B, 0, dev, seed=2)
n = rec.numel()
inp = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
inp[:n] = rec.reshape(-1)
out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
ctx = capi.Context(capi.make_params("sanger"), n + 16, 0)
st = torch.cuda.Stream()
for it in range(3):
    ctx.trim_device(inp.data_ptr(), n, 0, 0, [out.data_ptr(), 0, 0], [n + 64, 0, 0], st.cuda_stream)
    r = ctx.result_device(st.cuda_stream)
buf = (C.c_ulonglong * 12)()
lib.sk_debug_phase_cycles(buf, 1)
ctx.trim_device(inp.data_ptr(), n, 0, 0, [out.data_ptr(), 0, 0], [n + 64, 0, 0], st.cuda_stream)
r = ctx.result_device(st.cuda_stream)
lib.sk_debug_phase_cycles(buf, 0)
names = ["(warp 0's own S5-S7 work, part of the S5-S7 line)", "ticket + S1 load", "S2 masks+scan", "S3 pos + S4 look-back#1",
         "S5+S6 trim + S7 scan, to barrier", "(flush group: look-back#2)", "S8a staging copy (warp 0)",
         "(flush group: look-back#2 + flush, overlaps S5-S7)", "S7b totals + publish + descriptors", "S8a: wait for the slowest warp",
         "-", "-"]
tot = buf[1] + buf[2] + buf[3] + buf[4] + buf[6] + buf[8] + buf[9]
print("kernel %.3f ms, fused=%d; cycles summed over CTAs' thread 0:" % (r.kernel_ms, r.fused))
for k, v in zip(names, buf):
    print("  %-28s %6.2f%%" % (k, 100.0 * v / tot))
