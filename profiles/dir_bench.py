"""Workflow-level scaling: trim a directory of FASTQ files with trim_all.py on 1..N GPUs
(one bin/sickle process per GPU at a time; files on tmpfs).

    python profiles/dir_bench.py --files 8 --reads 4000000 --gpus 1,2
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sickle_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=8)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--procs-per-gpu", default="1,2")
    a = ap.parse_args()
    ind, outd = os.path.join(a.dir, "dir_bench_in"), os.path.join(a.dir, "dir_bench_out")
    shutil.rmtree(ind, ignore_errors=True)
    os.makedirs(ind)
    for k in range(a.files):
        with open(os.path.join(ind, "lane%02d.fastq" % k), "wb") as f:
            for s in range(0, a.reads, 1_000_000):
                synth.fixed_length_records(min(1_000_000, a.reads - s), 150, "sanger", seed=2, start=k * a.reads + s).tofile(f)
    total = a.files * a.reads
    for g in [int(x) for x in a.gpus.split(",")]:
        for ppg in [int(x) for x in a.procs_per_gpu.split(",")]:
            shutil.rmtree(outd, ignore_errors=True)
            t0 = time.perf_counter()
            rc = subprocess.call([sys.executable, os.path.join(ROOT, "trim_all.py"), "se", "sanger", ind, outd, "--gpus", str(g),
                                  "--procs-per-gpu", str(ppg)], stdout=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            print(json.dumps({"gpus": g, "procs_per_gpu": ppg, "files": a.files, "reads_per_file": a.reads, "rc": rc,
                              "wall_s": round(dt, 2), "reads_per_s": round(total / dt)}), flush=True)
    shutil.rmtree(ind, ignore_errors=True)
    shutil.rmtree(outd, ignore_errors=True)


if __name__ == "__main__":
    main()
