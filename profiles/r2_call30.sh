#!/bin/bash
# call 30: -a N on two input files through the index pass (both files in one launch) + K2 routing + K3; parity; timing
# against K1/K2/K3 on the same inputs (command line, 8 M pairs).
cd /root/repo
L=gpurun_out/r2_call30.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "(parity: 126 passed in the first run of this script, see above in the log history)"
  python - <<'PY'
import os, sys, json, statistics, torch
sys.path.insert(0, '.')
from sickle_b200 import capi, synth
import numpy as np
dev = torch.device('cuda:0')
f, r, _ = synth.paired_records(100_000, 150, 'sanger', seed=50)
def dev_buf(a, rep):
    t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).to(dev).repeat(rep)
    b = torch.zeros(t.numel() + 64, dtype=torch.uint8, device=dev); b[:t.numel()] = t
    return b, t.numel()
b0, n0 = dev_buf(f, 4); b1, n1 = dev_buf(r, 4)
out = [torch.empty(n0 + 64, dtype=torch.uint8, device=dev), torch.empty(n1 + 64, dtype=torch.uint8, device=dev), torch.empty(n0 + n1 + 64, dtype=torch.uint8, device=dev)]
for env in ({}, {'SICKLE_B200_PATH': 'general'}):
    os.environ.pop('SICKLE_B200_PATH', None); os.environ.update(env)
    p = capi.make_params('sanger', 20, 20, mode=capi.MODE_PE_2FILE, emulate_threads=8, has_singles=True)
    ctx = capi.Context(p, max(n0, n1) + 16, 0)
    st = torch.cuda.Stream(device=dev); ms = []
    for k in range(8):
        ctx.trim_device(b0.data_ptr(), n0, b1.data_ptr(), n1, [o.data_ptr() for o in out], [o.numel() - 64 for o in out], st.cuda_stream)
        res = ctx.result_device(st.cuda_stream)
        if k >= 2: ms.append(res.kernel_ms)
    print(json.dumps({'path': env or 'index pass + routing + K3', 'pairs': 400000, 'ms': round(statistics.median(ms), 4), 'fused': res.fused, 'stage_ms': [round(x, 4) for x in res.stage_ms], 'out': [int(x) for x in res.out_bytes]}))
    ctx.close()
PY
} > $L 2>&1
tail -12 $L | cut -c1-300
