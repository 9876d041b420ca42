#!/usr/bin/env python
"""Trim every FASTQ file of a directory -- the workflow driver of the reference (trim_all.py:1-108),
rebuilt for a multi-GPU host: files are independent, so they are dealt to the visible GPUs and one
`bin/sickle` process runs per GPU at a time (no collective, no shared state).

    python trim_all.py [se|pe] [solexa|illumina|sanger] input_dir/ output_dir/ [threads] [max_batch]
                       [--gpus N] [--procs-per-gpu P] [--dry-run]

Same positional arguments, file discovery and output names as the reference driver:
  se : every *.fq / *.fastq            -> <name>.trim.fastq
  pe : every *<sep>1.fq / *<sep>1.fastq (sep "." or, if fewer than two such files, "_") with its
       *<sep>2.* mate                   -> *1.trim.fastq, *2.trim.fastq, *s.trim.fastq (singles)
Existing outputs are skipped.  `threads` is passed on as `-a` only when given (`-a N` asks this
implementation to reproduce the reference's N-thread output ORDER, which costs speed; the default is
input order, i.e. the reference's `-a 1` bytes).  `max_batch` is passed on as `-b`.
Differences: gzip inputs (*.fq.gz, *.fastq.gz) are found too; the mate name is derived by replacing
the suffix (the reference's `rstrip` eats trailing characters of names like `s1.1.fq`); a failing
pair does not stop the other GPUs' files, the exit code is non-zero if any file failed.
"""
import argparse
import os
import queue
import subprocess
import sys
import threading

EXTS = (".fq", ".fastq", ".fq.gz", ".fastq.gz")


def split_ext(name):
    for e in sorted(EXTS, key=len, reverse=True):
        if name.endswith(e):
            return name[:-len(e)], e
    return name, ""


def plan(mode, qual_type, input_dir, output_dir, threads, max_batch, sickle):
    """List of (label, argv, outputs) -- pure function of the directory listing (unit-tested on CPU)."""
    names = sorted(n for n in os.listdir(input_dir) if split_ext(n)[1])
    extra = (["-a", str(threads)] if threads else []) + (["-b", str(max_batch)] if max_batch else [])
    jobs = []
    if mode == "se":
        for n in names:
            stem, _ = split_ext(n)
            out = os.path.join(output_dir, stem + ".trim.fastq")
            jobs.append((n, [sickle, "se", "-t", qual_type, "-f", os.path.join(input_dir, n), "-o", out] + extra, [out]))
    elif mode == "pe":
        sep = "."
        firsts = [n for n in names if split_ext(n)[0].endswith(sep + "1")]
        if len(firsts) < 2:
            sep = "_"
            firsts = [n for n in names if split_ext(n)[0].endswith(sep + "1")]
        for n in firsts:
            stem, ext = split_ext(n)
            base = stem[:-1]                                   # "...<sep>"
            mate = base + "2" + ext
            if mate not in names:
                raise FileNotFoundError("Input %s don't exist" % os.path.join(input_dir, mate))
            o1, o2, os_ = (os.path.join(output_dir, base + k + ".trim.fastq") for k in ("1", "2", "s"))
            jobs.append((n, [sickle, "pe", "-t", qual_type, "-f", os.path.join(input_dir, n), "-r",
                             os.path.join(input_dir, mate), "-o", o1, "-p", o2, "-s", os_] + extra, [o1, o2, os_]))
    else:
        raise ValueError("There is no '%s' mode available" % mode)
    return jobs


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=30).stdout
        return max(1, sum(1 for l in out.decode().splitlines() if l.startswith("GPU ")))
    except Exception:  # noqa: BLE001
        return 1


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("mode", choices=["se", "pe"])
    ap.add_argument("qual_type", choices=["solexa", "illumina", "sanger"])
    ap.add_argument("input_dir")
    ap.add_argument("output_dir")
    ap.add_argument("threads", nargs="?", type=int, default=0)
    ap.add_argument("max_batch", nargs="?", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=0, help="GPUs to use (default: all visible)")
    ap.add_argument("--procs-per-gpu", type=int, default=2,
                    help="files in flight per GPU (a single file is bound by its host-side file writes, not by the GPU)")
    ap.add_argument("--dry-run", action="store_true", help="print the commands, run nothing")
    a = ap.parse_args(argv)
    here = os.path.dirname(os.path.abspath(__file__))
    sickle = os.environ.get("SICKLE_B200_BIN") or os.path.join(here, "bin", "sickle")   # (override: an installed binary)
    jobs = plan(a.mode, a.qual_type, a.input_dir, a.output_dir, a.threads, a.max_batch, sickle)
    todo = []
    for label, cmd, outs in jobs:
        if any(os.path.exists(o) for o in outs):
            print("%s already exists, skiping it." % outs[0])
        else:
            todo.append((label, cmd))
    if a.dry_run:
        for _, cmd in todo:
            print("\t> " + " ".join(cmd))
        return 0
    if not os.path.exists(sickle):
        print("%s not found: run `make lib cli` first" % sickle, file=sys.stderr)
        return 1
    os.makedirs(a.output_dir, exist_ok=True)
    n_gpus = a.gpus or gpu_count()
    q = queue.Queue()
    for j in todo:
        q.put(j)
    failed = []

    # One long-lived `sickle batch` process per worker: CUDA start-up and the pinned buffers are paid
    # once per process, not once per file (measured on a B200 host: 1-2 s per file otherwise).  Each
    # process sees only its GPU (CUDA start-up touches every visible device).
    visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v]

    def gpu_env(gpu):
        return dict(os.environ, SICKLE_B200_DEVICE="0", CUDA_VISIBLE_DEVICES=visible[gpu] if gpu < len(visible) else str(gpu))

    def quote(arg):
        return '"%s"' % arg if (" " in arg or "\t" in arg) else arg

    def spawn(gpu):
        return subprocess.Popen([sickle, "batch"], env=gpu_env(gpu), stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                universal_newlines=True, bufsize=1)

    def worker(gpu):
        # A batch process that dies (crash, usage error) fails the file it was working on only: the worker
        # starts another one and goes on with the queue, as the reference's driver does with one process per
        # file (trim_all.py:62-108).  Three deaths in a row and the worker gives up -- what is then still
        # queued is taken by the other workers or reported as failed below.
        proc, deaths = spawn(gpu), 0
        try:
            while deaths < 3:
                try:
                    label, cmd = q.get_nowait()
                except queue.Empty:
                    return
                print("\t[gpu %d] > %s" % (gpu, " ".join(cmd)), flush=True)
                rc = None
                try:
                    proc.stdin.write(" ".join(quote(c) for c in cmd[1:]) + "\n")
                    proc.stdin.flush()
                    for line in proc.stdout:
                        if line.startswith("##rc "):
                            rc = int(line.split()[1])
                            break
                        sys.stdout.write(line)
                except OSError:
                    pass
                if rc is None:                      # the batch process died
                    failed.append((label, proc.wait()))
                    deaths += 1
                    proc = spawn(gpu)
                    continue
                deaths = 0
                if rc != 0:
                    failed.append((label, rc))
        finally:
            try:
                proc.stdin.close()
            except OSError:
                pass
            proc.wait()

    used = min(n_gpus, max(1, len(todo)))
    threads = [threading.Thread(target=worker, args=(k % used,))
               for k in range(min(used * max(1, a.procs_per_gpu), max(1, len(todo))))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    while True:                                     # every worker gave up: nothing queued goes unreported
        try:
            label, _ = q.get_nowait()
        except queue.Empty:
            break
        failed.append((label, -1))
    for label, rc in failed:
        print("FAILED (%d): %s" % (rc, label), file=sys.stderr)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
