#!/usr/bin/env python
"""Generate tests/golden/*.fastq fixtures and golden.json from the UNMODIFIED reference binary.

Run in the build container (where /root/reference exists) after `make -C oracle ref`:

    python tests/golden/make_golden.py

It (1) writes small deterministic synthetic FASTQ inputs next to this script, (2) runs the
reference (`oracle/_ref/sickle`; for `se` the `sickle_sync` build, because the unpatched `se`
mode races and crashes, SURVEY.md 9-D5 -- the unpatched binary is run as well and must agree
whenever it finishes) over a matrix of flags, and (3) records md5 / size / summary counters of
every output file, plus exit code and stderr for the error fixtures, in golden.json.

The committed fixtures + golden.json are what the tests read; /root/reference is never needed at
test time.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sickle_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "sickle")
REF_SYNC = os.path.join(ROOT, "oracle", "_ref", "sickle_sync")


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def write(name: str, data: bytes) -> str:
    path = os.path.join(HERE, name)
    with open(path, "wb") as f:
        f.write(data)
    return name


# ---------------------------------------------------------------------------------------------
# fixtures
# ---------------------------------------------------------------------------------------------
def illumina15_pairs(n_pairs: int, seed: int):
    """Illumina-1.5 style pairs: offset 64, 'B' (Q2) tails, many Ns, names .../1 .../2."""
    rng = np.random.default_rng(seed)
    L = 150
    recs = {1: [], 2: []}
    for i in range(n_pairs):
        x, y = 1500 + 7 * i, 1900 + 13 * (i % 97)
        for mate in (1, 2):
            name = ("@2242:2:1101:%d:%d/%d" % (x, y, mate)).encode()
            seq = synth._BASES[rng.integers(0, 4, L)].copy()
            seq[rng.random(L) < 0.01] = ord("N")
            if rng.random() < 0.3:
                seq[0] = ord("N")
            q = synth._quals(rng, 1, L, 0, 40, 0.15)[0]
            if rng.random() < 0.5:  # the 'B' tail of CASAVA 1.5
                t = int(rng.integers(1, 80))
                q[L - t:] = 2
            qual = (q + 64).astype(np.uint8)
            recs[mate].append((name, seq.tobytes(), qual.tobytes()))
    inter, fwd, rev = [], [], []
    for i in range(n_pairs):
        for mate, dst in ((1, fwd), (2, rev)):
            name, seq, qual = recs[mate][i]
            inter.append(name + b"\n" + seq + b"\n+\n" + qual + b"\n")
            dst.append(name + b"\n" + seq + b"\n+" + name[1:] + b"\n" + qual + b"\n")
    return b"".join(inter), b"".join(fwd), b"".join(rev)


def tiny_reads(n: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 31))
        seq = synth._BASES[rng.integers(0, 4, L)].copy()
        if rng.random() < 0.1:
            seq[int(rng.integers(0, L))] = ord("N")
        if rng.random() < 0.05:
            seq[int(rng.integers(0, L))] = ord("n")
        q = rng.integers(2, 42, L)
        if rng.random() < 0.5:
            q = np.sort(q)[::-1]
        out.append(b"@t%05d\n" % i + seq.tobytes() + b"\n+\n" + (q + 33).astype(np.uint8).tobytes() + b"\n")
    return b"".join(out)


def make_fixtures():
    fx = {}
    fx["se_r150"] = write("se_r150.fastq", synth.fixed_length_records(400, 150, "sanger", seed=2).tobytes())
    f, r, inter = synth.paired_records(300, 150, "sanger", seed=3)
    fx["pe_r150_f"] = write("pe_r150_f.fastq", f.tobytes())
    fx["pe_r150_r"] = write("pe_r150_r.fastq", r.tobytes())
    fx["pe_r150_inter"] = write("pe_r150_inter.fastq", inter.tobytes())
    inter, fwd, rev = illumina15_pairs(250, seed=15)
    fx["il15_inter"] = write("il15_inter.fastq", inter)
    fx["il15_f"] = write("il15_f.fastq", fwd)
    fx["il15_r"] = write("il15_r.fastq", rev)
    small = synth.variable_length_records(160, 1, 400, "illumina", seed=41)
    big = synth.variable_length_records(16, 1000, 12000, "illumina", seed=42, name_prefix="@long")
    fx["varlen_illumina"] = write("varlen_illumina.fastq", small + big)
    small = synth.variable_length_records(160, 1, 400, "solexa", seed=43)
    big = synth.variable_length_records(16, 1000, 12000, "solexa", seed=44, name_prefix="@long")
    fx["varlen_solexa"] = write("varlen_solexa.fastq", small + big)
    fx["tiny_reads"] = write("tiny_reads.fastq", tiny_reads(300, seed=5))

    # --- edge / error fixtures (built from se_r150) ---
    base = synth.fixed_length_records(40, 100, "sanger", seed=9, name_fmt="@E%04d", bad_frac=0.0)
    W = base.shape[1]
    nm = 7  # "@E0000" + "\n"
    q0 = nm + 100 + 1 + 2  # column where the quality string starts

    def good_quals(b):
        # all-Q40 reads with Q2 from position 40 on: the 3' break happens at window 36
        # (ws = 10), so positions [0, 46) are range-checked and 46.. are not.
        b[:, q0:q0 + 100] = 33 + 40
        b[:, q0 + 40:q0 + 100] = 33 + 2
        return b

    b = good_quals(base.copy())
    b[17, q0 + 45] = ord(" ")                       # inside the visited prefix -> error, position 46
    fx["err_qual_range"] = write("err_qual_range.fastq", b.tobytes())
    b = good_quals(base.copy())
    b[17, q0 + 46] = ord(" ")                       # just past the visited prefix -> silently kept
    fx["ok_qual_unvisited"] = write("ok_qual_unvisited.fastq", b.tobytes())
    b = good_quals(base.copy())
    b[5, q0 + 3] = 200                              # >= 0x80 is negative as signed char
    fx["err_qual_highbit"] = write("err_qual_highbit.fastq", b.tobytes())
    data = good_quals(base.copy()).tobytes()
    fx["err_no_final_newline"] = write("err_no_final_newline.fastq", data[:-1])
    fx["ok_trailing_partial"] = write("ok_trailing_partial.fastq", data + b"@partial\nACGT\n")
    recs = [data[i * W:(i + 1) * W] for i in range(40)]
    bad = recs[:]
    bad[11] = bad[11].replace(b"@E0011", b"xE0011")
    fx["err_id_char"] = write("err_id_char.fastq", b"".join(bad))
    bad = recs[:]
    bad[12] = b"@\n" + bad[12].split(b"\n", 1)[1]
    fx["err_id_short"] = write("err_id_short.fastq", b"".join(bad))
    bad = recs[:]
    parts = bad[13].split(b"\n")
    parts[3] = parts[3][:-7]
    bad[13] = b"\n".join(parts)
    fx["err_len_mismatch"] = write("err_len_mismatch.fastq", b"".join(bad))
    fx["ok_crlf"] = write("ok_crlf.fastq", data.replace(b"\n", b"\r\n"))
    return fx


# ---------------------------------------------------------------------------------------------
# running the reference
# ---------------------------------------------------------------------------------------------
def run(cmd, timeout=60):
    try:
        p = subprocess.run(cmd, capture_output=True, timeout=timeout)
        return p.returncode, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")
    except subprocess.TimeoutExpired:
        return 124, "", "timeout"


def parse_counts(stdout: str):
    c = {}
    pats = {
        "total": r"Total FastQ records: (\d+)", "kept": r"FastQ records kept: (\d+)",
        "discard": r"FastQ records discarded: (\d+)",
        "kept_p": r"FastQ paired records kept: (\d+)", "discard_p": r"FastQ paired records discarded: (\d+)",
        "kept_s": r"FastQ single records kept: (\d+)", "discard_s": r"FastQ single records discarded: (\d+)",
        "kept_s1": r"FastQ single records kept: \d+ \(from PE1: (\d+)", "kept_s2": r"single records kept: \d+ \(from PE1: \d+, from PE2: (\d+)",
        "discard_s1": r"single records discarded: \d+ \(from PE1: (\d+)", "discard_s2": r"single records discarded: \d+ \(from PE1: \d+, from PE2: (\d+)",
    }
    for k, pat in pats.items():
        m = re.search(pat, stdout)
        if m:
            c[k] = int(m.group(1))
    return c


def strip_debug(stderr: str) -> str:
    return stderr


def ref_case(mode: str, inputs: dict, flags: list, outputs: list, threads: int = 1):
    """Run the reference; returns dict(outputs={name:{md5,bytes}}, counts, rc, stderr)."""
    import collections
    with tempfile.TemporaryDirectory() as td:
        def once(binary, attempt):
            cmd = [binary, mode]
            for k, v in inputs.items():
                cmd += [k, os.path.join(HERE, v)]
            outs = {}
            for k in outputs:
                outs[k] = os.path.join(td, "out_%s_%d" % (k.strip("-"), attempt))
                cmd += [k, outs[k]]
            cmd += flags + ["-a", str(threads)]
            rc, so, se = run(cmd)
            res = {"rc": rc, "counts": parse_counts(so), "stderr": se, "outputs": {}}
            for k, pth in outs.items():
                data = open(pth, "rb").read() if os.path.exists(pth) else b""
                res["outputs"][k] = {"md5": md5(data), "bytes": len(data)}
                if os.path.exists(pth):
                    os.unlink(pth)
            return res

        if mode == "se":
            first = once(REF_SYNC, 0)
            res = once(REF, 1)
            # unpatched run: must agree with sync whenever it exits cleanly with the same rc
            agree = (res["rc"] == first["rc"] and res["outputs"] == first["outputs"])
            first["unpatched_agrees"] = bool(agree) if res["rc"] in (0, 1) else "crash rc=%d" % res["rc"]
            return first
        # pe: concurrent output_paired threads race for batch_lock, so the batch order of a run
        # is occasionally permuted.  Take the clear majority of up to 15 runs.
        seen = collections.Counter()
        keep = {}
        for attempt in range(15):
            res = once(REF, attempt)
            key = json.dumps(res["outputs"], sort_keys=True) + str(res["rc"])
            seen[key] += 1
            keep[key] = res
            top = seen.most_common(2)
            if top[0][1] >= 3 and (len(top) == 1 or top[0][1] >= 2 * top[1][1]):
                keep[top[0][0]]["runs"] = dict(agree=top[0][1], total=attempt + 1)
                return keep[top[0][0]]
        raise RuntimeError("reference never produced a majority output for %r %r: %r" % (mode, flags, seen.most_common(4)))


def main():
    if not (os.path.exists(REF) and os.path.exists(REF_SYNC)):
        sys.exit("build the reference first: make -C oracle ref")
    fx = make_fixtures()
    cases = []

    def add(cid, mode, inputs, flags, outputs, threads=1, note=""):
        r = ref_case(mode, inputs, flags, outputs, threads)
        if r["rc"] not in (0, 1):
            raise RuntimeError("reference crashed on %s: rc=%d" % (cid, r["rc"]))
        r.update({"id": cid, "mode": mode, "inputs": inputs, "flags": flags, "threads": threads, "note": note})
        if r["rc"] == 0:
            r["stderr"] = ""
        cases.append(r)
        print("%-44s rc=%d %s" % (cid, r["rc"], {k: v["bytes"] for k, v in r["outputs"].items()}), flush=True)

    flagsets = {
        "default": [], "q30l50": ["-q", "30", "-l", "50"], "q2l0": ["-q", "2", "-l", "0"],
        "q35l0": ["-q", "35", "-l", "0"], "x": ["-x"], "n": ["-n"], "xn": ["-x", "-n"],
        "q41l1": ["-q", "41", "-l", "1"], "q0": ["-q", "0"], "xq30l0": ["-x", "-q", "30", "-l", "0"],
    }
    se_fixt = {
        "se_r150": ["sanger"], "pe_r150_inter": ["sanger"], "il15_inter": ["illumina", "sanger", "solexa"],
        "il15_f": ["illumina"], "varlen_illumina": ["illumina", "sanger", "solexa"],
        "varlen_solexa": ["solexa", "sanger"], "tiny_reads": ["sanger"],
    }
    for fixture, types in se_fixt.items():
        for t in types:
            for fname, fl in flagsets.items():
                add("se.%s.%s.%s" % (fixture, t, fname), "se", {"-f": fx[fixture]}, ["-t", t] + fl, ["-o"])
    for n in (2, 3, 4, 8):
        add("se.se_r150.sanger.default.a%d" % n, "se", {"-f": fx["se_r150"]}, ["-t", "sanger"], ["-o"], threads=n)
        # (varlen_* cannot be used with -a N: a batch with fewer records than threads makes the
        #  reference's output_single walk a stale queue and segfault, trim_single.cpp:384-385)
        add("se.il15_inter.illumina.xn.a%d" % n, "se", {"-f": fx["il15_inter"]},
            ["-t", "illumina", "-x", "-n"], ["-o"], threads=n)

    pe_sets = [("pe_r150", "sanger"), ("il15", "illumina"), ("il15", "sanger")]
    for base, t in pe_sets:
        for fname, fl in flagsets.items():
            add("pe2.%s.%s.%s" % (base, t, fname), "pe", {"-f": fx[base + "_f"], "-r": fx[base + "_r"]},
                ["-t", t] + fl, ["-o", "-p", "-s"])
            add("pei.%s.%s.%s" % (base, t, fname), "pe", {"-c": fx[base + "_inter"]}, ["-t", t] + fl, ["-m", "-s"])
        add("pei_nosingles.%s.%s" % (base, t), "pe", {"-c": fx[base + "_inter"]}, ["-t", t], ["-m"])
        for n in (2, 4):
            add("pe2.%s.%s.default.a%d" % (base, t, n), "pe", {"-f": fx[base + "_f"], "-r": fx[base + "_r"]},
                ["-t", t], ["-o", "-p", "-s"], threads=n)
            add("pei.%s.%s.default.a%d" % (base, t, n), "pe", {"-c": fx[base + "_inter"]}, ["-t", t], ["-m", "-s"],
                threads=n)

    for name in ("err_qual_range", "ok_qual_unvisited", "err_qual_highbit", "err_no_final_newline",
                 "ok_trailing_partial", "err_id_char", "err_id_short", "err_len_mismatch", "ok_crlf"):
        add("se.%s" % name, "se", {"-f": fx[name]}, ["-t", "sanger"], ["-o"])
    add("se.err_qual_range.short", "se", {"-f": fx["err_qual_range"]}, ["-t", "sanger", "-l", "101"], ["-o"],
        note="reads shorter than -l are discarded before any quality byte is range-checked")

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "pentalpha/sickle @ /root/reference",
                   "cases": cases}, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
