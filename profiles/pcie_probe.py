"""Host<->device copy bandwidth per GPU, one GPU at a time and all together (evidence for why the
end-to-end number does not scale with GPUs on a given host while the kernel number does).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("gloo", init_method="env://", rank=rank, world_size=world)
N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return reps * N / (time.perf_counter() - t0) / 1e9


run(True, True, 1)
res = {}
for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    for who in list(range(world)) + [-1]:          # one rank at a time, then all ranks together
        dist.barrier()
        if who in (-1, rank):
            res["%s_%s" % (name, "all" if who < 0 else "solo")] = round(run(h2d, d2h), 1)
        dist.barrier()
out = [None] * world
dist.all_gather_object(out, res)
if rank == 0:
    print(json.dumps({"unit": "GB/s per direction per GPU, 1 GiB pinned copies", "ranks": out}))
dist.destroy_process_group()
