"""The command line on 1, 2, 4 (, 8) physical GPUs: one FASTQ file on tmpfs through `bin/sickle se` with
SICKLE_B200_GPUS=G (host/trimmer.cpp run_devices: the input is cut on the host into whole-record batches, batch k
runs on device k mod G, outputs are appended in batch order), wall clock, the program's own stage line, and the md5
of the output -- which must not depend on G.  Then the reference (`sickle_sync se -a <cores>`) on a prefix of the
same input.  This is BASELINE.json configs[4] through the product, on a stated subsample.

    python profiles/multi_gpu_cli.py [--reads 24000000] [--gpus 1,2,4] [--ref-reads 4000000]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=24_000_000)
    ap.add_argument("--ref-reads", type=int, default=4_000_000)
    ap.add_argument("--gpus", default="1,2,4")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    src, out = os.path.join(a.dir, "mg_in.fastq"), os.path.join(a.dir, "mg_out.fastq")
    t0 = time.perf_counter()
    how = bench.write_r150_file(src, a.reads)
    print(json.dumps({"generated": a.reads, "bytes": os.path.getsize(src), "with": how, "s": round(time.perf_counter() - t0, 1)}), flush=True)
    exe = os.path.join(ROOT, "bin", "sickle")
    md5s = {}
    for extra, tag in (([], "input order"), (["-a", "16"], "-a 16 (reference output order)")):
        for g in [int(x) for x in a.gpus.split(",")]:
            best, line = None, None
            for _ in range(a.repeat):
                if os.path.exists(out):
                    os.unlink(out)
                env = dict(os.environ, SICKLE_B200_GPUS=str(g))
                t0 = time.perf_counter()
                p = subprocess.run([exe, "se", "-f", src, "-t", "sanger", "-o", out, "-d"] + extra, capture_output=True, env=env)
                dt = time.perf_counter() - t0
                if p.returncode != 0:
                    print(json.dumps({"gpus": g, "mode": tag, "rc": p.returncode, "stderr": p.stderr.decode()[-400:]}), flush=True)
                    break
                if best is None or dt < best:
                    best = dt
                    line = [l for l in p.stderr.decode().splitlines() if l.startswith("[sickle_b200]")]
            if best is None:
                continue
            m = md5_file(out)
            md5s.setdefault(tag, set()).add(m)
            print(json.dumps({"gpus": g, "mode": tag, "reads": a.reads, "wall_s": round(best, 3), "reads_per_s": round(a.reads / best),
                              "in_GBps": round(os.path.getsize(src) / best / 1e9, 2), "out_bytes": os.path.getsize(out), "md5": m,
                              "stages": line}), flush=True)
    print(json.dumps({"md5_independent_of_gpu_count": {k: len(v) == 1 for k, v in md5s.items()}}), flush=True)
    # the reference on a prefix (whole records: R150 is 325 bytes per record)
    ref = os.path.join(ROOT, "oracle", "_ref", "sickle_sync")
    if os.path.exists(ref) and a.ref_reads:
        pre = os.path.join(a.dir, "mg_prefix.fastq")
        with open(src, "rb") as f, open(pre, "wb") as g_:
            left = a.ref_reads * 325
            while left:
                b = f.read(min(left, 1 << 24))
                g_.write(b)
                left -= len(b)
        t0 = time.perf_counter()
        p = subprocess.run([ref, "se", "-f", pre, "-t", "sanger", "-o", out + ".ref", "-a", str(os.cpu_count()), "-b", "512"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        print(json.dumps({"tool": "reference sickle_sync -a %d" % os.cpu_count(), "rc": p.returncode, "reads": a.ref_reads,
                          "wall_s": round(dt, 2), "reads_per_s": round(a.ref_reads / dt)}), flush=True)
        # same bytes as this program on the same prefix (-a 1 order: compare sorted record multisets by size only)
        if os.path.exists(out):
            os.unlink(out)
        subprocess.run([exe, "se", "-f", pre, "-t", "sanger", "-o", out], capture_output=True)
        print(json.dumps({"prefix_same_size_as_reference": os.path.getsize(out) == os.path.getsize(out + ".ref")}), flush=True)
    for f_ in os.listdir(a.dir):
        if f_.startswith("mg_"):
            os.unlink(os.path.join(a.dir, f_))


if __name__ == "__main__":
    main()
