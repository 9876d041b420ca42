"""Wall-clock of the command-line tools on the same FASTQ file (page cache / tmpfs, no disk):
bin/sickle (this repo, GPU) against oracle/_ref/sickle (the unmodified reference, all host cores).

    python profiles/cli_bench.py [--reads 8000000] [--dir /dev/shm] [--skip-ref]

Prints one JSON object per tool.  This is the "whole program" view of SURVEY.md 8-f1 (host I/O
path); bench.py's e2e figure is the library call with host buffers.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sickle_b200 import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def timed(cmd, env=None):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return time.perf_counter() - t0, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE for the tools' environment")
    ap.add_argument("--input-format", default="plain", choices=["plain", "bgzf", "gzip"])
    ap.add_argument("--gzip-out", action="store_true", help="-g (b200 only: the reference's -g is broken)")
    ap.add_argument("--paired", action="store_true", help="interleaved pairs (pe -c) instead of se")
    ap.add_argument("--threads", type=int, default=0, help="pass -a N to bin/sickle too (reference output order)")
    ap.add_argument("--two-files", action="store_true", help="pe -f -r -o -p -s (forward / reverse files)")
    a = ap.parse_args()
    src = os.path.join(a.dir, "cli_bench_in.fastq")
    src2 = os.path.join(a.dir, "cli_bench_in2.fastq")
    chunk = 1_000_000
    if a.two_files:
        a.paired = True
        with open(src, "wb") as f, open(src2, "wb") as f2:
            for s in range(0, a.reads, chunk):
                n = min(chunk, a.reads - s)
                fw, rv, _ = synth.paired_records(n // 2, 150, "sanger", seed=3, start=s // 2)
                fw.tofile(f)
                rv.tofile(f2)
    fast = False
    if not a.paired and not a.two_files:
        try:   # the GPU generator (same model): 24 M reads in seconds instead of minutes
            import torch

            if torch.cuda.is_available():
                import bench

                bench.write_r150_file(src, a.reads)
                fast = True
        except Exception:  # noqa: BLE001
            fast = False
    with open(src, "ab" if (a.two_files or fast) else "wb") as f:
        for s in range(0, 0 if (a.two_files or fast) else a.reads, chunk):
            n = min(chunk, a.reads - s)
            if a.paired:
                synth.paired_records(n // 2, 150, "sanger", seed=3, start=s // 2)[2].tofile(f)
            else:
                synth.fixed_length_records(n, 150, "sanger", seed=2, start=s).tofile(f)
    if a.input_format != "plain":
        gz = src + ".gz"
        t0 = time.perf_counter()
        if a.input_format == "bgzf":
            subprocess.run([os.path.join(ROOT, "bin", "io_tool"), src, gz, str(1 << 26), "1"], check=True, stdout=subprocess.PIPE)
        else:
            subprocess.run("gzip -1 -c %s > %s" % (src, gz), shell=True, check=True)
        print(json.dumps({"compressed_with": a.input_format, "s": round(time.perf_counter() - t0, 2),
                          "ratio": round(os.path.getsize(src) / os.path.getsize(gz), 2)}))
        os.unlink(src)
        src = gz
    size = os.path.getsize(src) + (os.path.getsize(src2) if a.two_files else 0)
    outs = {}
    tools = [("b200", os.path.join(ROOT, "bin", "sickle"), ["-d"])]
    if not a.skip_ref:
        # se: the reference sources with synchronous output (oracle/Makefile) -- the unpatched binary's
        # detached writer thread races and crashes on large inputs
        ref = "sickle" if a.paired else "sickle_sync"
        tools.append(("reference", os.path.join(ROOT, "oracle", "_ref", ref), ["-a", str(os.cpu_count())]))
    env = dict(os.environ)
    for kv in a.env:
        k, v = kv.split("=", 1)
        env[k] = v
    for name, exe, extra in tools:
        out = os.path.join(a.dir, "cli_bench_out_%s.fastq" % name)
        sng = os.path.join(a.dir, "cli_bench_sng_%s.fastq" % name)
        if a.two_files:
            cmd = [exe, "pe", "-f", src, "-r", src2, "-t", "sanger", "-o", out, "-p", out + ".2", "-s", sng] + extra
        elif a.paired:
            cmd = [exe, "pe", "-c", src, "-t", "sanger", "-m", out, "-s", sng] + extra
        else:
            cmd = [exe, "se", "-f", src, "-t", "sanger", "-o", out] + extra
        if a.gzip_out and name == "b200":
            cmd.append("-g")
        if a.threads and name == "b200":
            cmd += ["-a", str(a.threads)]
        best = None
        for _ in range(a.repeat if name == "b200" else 1):
            for p_ in (out, sng):
                if os.path.exists(p_):
                    os.unlink(p_)
            dt, p = timed(cmd, env)
            if p.returncode != 0:
                print(json.dumps({"tool": name, "rc": p.returncode, "stderr": p.stderr.decode()[-400:]}))
                best = None
                break
            if best is None or dt < best:
                best, best_p = dt, p
        if best is None:
            continue
        outs[name] = out
        print(json.dumps({"tool": name, "cmd": " ".join(cmd[1:]), "reads": a.reads, "in_bytes": size,
                          "out_bytes": os.path.getsize(out), "wall_s": round(best, 3),
                          "reads_per_s": round(a.reads / best), "in_GBps": round(size / best / 1e9, 2),
                          "summary": best_p.stdout.decode().strip().splitlines()[-3:],
                          "stages": [l for l in best_p.stderr.decode().splitlines() if l.startswith("[sickle_b200]")]}))
    # same bytes?  (the reference at -a N > 1 permutes records: compare sizes there, md5 against a
    # synchronous reference run on a prefix is what tests/ do)
    if "b200" in outs and "reference" in outs:
        print(json.dumps({"same_size": os.path.getsize(outs["b200"]) == os.path.getsize(outs["reference"])}))
    for f_ in os.listdir(a.dir):
        if f_.startswith("cli_bench_"):
            os.unlink(os.path.join(a.dir, f_))


if __name__ == "__main__":
    main()
