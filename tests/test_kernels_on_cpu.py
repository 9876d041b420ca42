"""The CUDA kernels themselves -- sickle_b200/csrc/kf_fused.cuh (single pass, four tile sizes) and the
general path k1_index / k2_trim / k3_emit -- compiled for the HOST and run on CPU against the oracle.

tests/host_stub/simt/simt_host.h supplies just enough of the CUDA execution model: a CTA is an OS
thread, its threads are fibers, barriers and warp collectives are rendezvous, global atomics are the
host's, and the CTAs of a grid run concurrently (the decoupled look-backs spin on their predecessors).
The few PTX statements in the kernels have a plain C++ spelling beside them (`#if defined(__CUDACC__)`);
the SASS nvcc produces is unchanged by that.  tests/host_stub/kernels_harness.cpp launches the kernels
the way capi.cu does and compares output streams, counters, consumed bytes, record counts and the first
data error with so_run on the same bytes.

What this pins without a GPU: the parsing, the record / tile ownership rules, look-backs, the window
arithmetic, routing, scans and every byte the kernels write -- i.e. the logic.  What it cannot show:
timing, memory-model races that need real parallel warps, the TMA / mbarrier path (a memcpy here).  The
`-m gpu` tests run the same comparisons on the device.  Test infrastructure only.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
QT = {"sanger": 1, "solexa": 2, "illumina": 3}
MODES = {"se": 0, "pe2": 1, "pei": 2, "peM": 3}


def build_harness(name, defines=(), extra=()):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, name)
    stub = os.path.join(ROOT, "tests", "host_stub")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-fno-extern-tls-init"] + list(extra) + ["-D" + d for d in defines] +
                          ["-I" + os.path.join(stub, "simt"), "-I" + os.path.join(stub, "simt", "include"),
                           "-I" + os.path.join(ROOT, "sickle_b200", "csrc"), "-I" + os.path.join(ROOT, "oracle"),
                           "-x", "c++", os.path.join(stub, "kernels_harness.cpp"), "-x", "c", os.path.join(ROOT, "oracle", "sickle_oracle.c"),
                           "-o", out, "-lpthread"])
    return out


@pytest.fixture(scope="module")
def harness():
    return build_harness("kernels_harness")


def run(exe, path, mode="se", qualtype="sanger", q=20, l=20, x=False, n=False, singles=True, kernel="fused9", ctas=3, first=0, path2=None,
        threads=1, env=None):
    cmd = [exe, path, str(MODES[mode]), str(QT[qualtype]), str(q), str(l), str(int(x)), str(int(n)), str(int(singles)), kernel, str(ctas), str(first)]
    if path2:
        cmd.append(path2)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, KH_THREADS=str(threads), **(env or {})))
    return p.returncode, p.stdout.strip(), p.stderr


def check(exe, path, tag, kernels=("fused3", "fused5", "fused7", "fused9", "fused11", "general"), **kw):
    """Every kernel path must agree with the oracle; a fused kernel may instead hand the batch over
    (FASTFAIL), which is what capi.cu's rerun_if_needed then sends through the general path."""
    seen = []
    for k in kernels:
        rc, out, err = run(exe, path, kernel=k, **kw)
        assert rc == 0 and (out.startswith("OK") or (k != "general" and out.startswith("FASTFAIL"))), (tag, k, out, err[-500:])
        seen.append(out.split()[0])
    return seen


def test_bench_workload_and_flag_combinations(harness, tmp_path):
    from sickle_b200 import synth

    se = str(tmp_path / "se.fq")
    open(se, "wb").write(synth.fixed_length_records(3000, 150, "sanger", seed=5).tobytes())
    f, r, inter = synth.paired_records(1200, 150, "sanger", seed=6)
    pf, pr, il = (str(tmp_path / n) for n in ("f.fq", "r.fq", "il.fq"))
    open(pf, "wb").write(f.tobytes()); open(pr, "wb").write(r.tobytes()); open(il, "wb").write(inter.tobytes())
    for ctas in (1, 2, 5):
        assert check(harness, se, ("se", ctas), ctas=ctas, first=ctas * 3) == ["OK"] * 6      # 150-base reads never leave the fused kernel
    for q, l, x, n in ((30, 5, True, False), (10, 0, False, True), (25, 1, True, True), (0, 0, False, False), (41, 151, False, False)):
        check(harness, se, ("se flags", q, l, x, n), q=q, l=l, x=x, n=n, first=7)
    for singles in (True, False):
        check(harness, il, ("pei", singles), mode="pei", singles=singles, first=11)
    check(harness, il, "peM", mode="peM", singles=False)
    K2F = ("fused5", "fused7", "fused9", "general")     # two files: both passes of the single-pass kernel, and K1/K2/K3
    assert check(harness, pf, "pe2", kernels=K2F, mode="pe2", path2=pr, first=4) == ["OK"] * 4
    assert check(harness, pf, "pe2 no singles", kernels=K2F, mode="pe2", path2=pr, singles=False) == ["OK"] * 4
    check(harness, pf, "pe2 -x -n", kernels=K2F, mode="pe2", path2=pr, q=30, l=5, x=True, n=True, first=9, ctas=5)


def test_many_tiles_and_ctas(harness, tmp_path):
    """Hundreds of tiles over 8 concurrent CTAs (tickets, both look-backs, deferred flush across many
    predecessors): 60,000 fixed-length reads, 30,000 interleaved pairs of mixed lengths."""
    from sickle_b200 import synth
    from test_oracle_fuzz_vs_ref import _records

    se, il = str(tmp_path / "se.fq"), str(tmp_path / "il.fq")
    open(se, "wb").write(synth.fixed_length_records(60000, 150, "sanger", seed=35).tobytes())
    rng = np.random.default_rng(36)
    open(il, "wb").write(_records(rng, 60000, 220, "sanger"))
    assert check(harness, se, "se 60k", kernels=("fused7", "fused9", "general"), ctas=8, first=13) == ["OK"] * 3
    assert check(harness, il, "pei 60k", kernels=("fused5", "general"), mode="pei", ctas=8, first=2, n=True) == ["OK"] * 2


def test_two_files_on_the_single_pass_kernel(harness, tmp_path):
    """`pe -f -r` through both passes of kf_fused: mates whose records differ in size (the tickets are dealt in
    proportion to the files' tile counts), files with unequal record counts or a cut-off tail (the extra records
    are not part of the batch and stay unconsumed), many tiles over concurrent CTAs, random thread order, a
    damaged record (the kernel hands the batch over), flags and encodings."""
    from sickle_b200 import synth

    K = ("fused5", "fused7", "fused9")

    def files(a, b):
        pa, pb = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
        open(pa, "wb").write(a); open(pb, "wb").write(b)
        return pa, pb

    # mates of different, varying lengths: 9,000 pairs, file 2 about half as large as file 1
    pa, pb = files(synth.variable_length_records(9000, 120, 320, "illumina", 81, log_uniform=False, plus_name_every=3),
                   synth.variable_length_records(9000, 80, 170, "illumina", 82, log_uniform=False))
    seen = check(harness, pa, "varlen", kernels=K + ("general",), mode="pe2", path2=pb, qualtype="illumina", ctas=6, first=5)
    assert seen[-1] == "OK" and "OK" in seen[:3]
    check(harness, pa, "varlen no singles -n", kernels=K, mode="pe2", path2=pb, qualtype="illumina", n=True, singles=False, ctas=3)
    # unequal record counts / a file that ends inside a record
    f, r, _ = synth.paired_records(2600, 150, "sanger", seed=78)
    fb, rb = f.tobytes(), r.tobytes()
    for a, b in ((fb, rb[:len(rb) * 3 // 4]), (fb[:len(fb) // 3 + 17], rb), (fb[:-60], rb), (fb, rb[:327 * 5]), (fb[:200], rb)):
        pa, pb = files(a, b)
        assert check(harness, pa, ("unequal", len(a), len(b)), kernels=K + ("general",), mode="pe2", path2=pb, ctas=4, first=11) == ["OK"] * 4
    # many tiles, 8 CTAs, fibres in random order
    f, r, _ = synth.paired_records(40000, 150, "sanger", seed=79)
    pa, pb = files(f.tobytes(), r.tobytes())
    assert check(harness, pa, "40k pairs", kernels=("fused7", "fused9"), mode="pe2", path2=pb, ctas=8, first=2) == ["OK"] * 2
    for seed in (1, 2):
        rc, out, err = run(harness, pa, mode="pe2", path2=pb, kernel="fused9", ctas=5, env={"SIMT_SHUFFLE": str(seed)})
        assert rc == 0 and out.startswith("OK"), (seed, out, err[-300:])
    # a damaged record in either file: the batch is handed to the general path (which reports the error)
    lines = rb.split(b"\n")
    lines[4 * 1700 + 3] = lines[4 * 1700 + 3][:90]
    pa, pb = files(fb, b"\n".join(lines))
    assert check(harness, pa, "damaged mate 2", kernels=K, mode="pe2", path2=pb) == ["FASTFAIL"] * 3
    rc, out, err = run(harness, pa, mode="pe2", path2=pb, kernel="general")
    assert rc == 0 and out.startswith("OK error kind=5 record=1700"), out


def test_reference_thread_order(harness, tmp_path):
    """-a N: the general path deals the records of a batch to N queues and emits queue after queue, as the
    reference does (src/trim_single.cpp:263,273-274; src/trim_paired.cpp:349,388,403)."""
    from sickle_b200 import synth

    se, pf, pr, il = (str(tmp_path / n) for n in ("se.fq", "f.fq", "r.fq", "il.fq"))
    open(se, "wb").write(synth.fixed_length_records(2500, 150, "sanger", seed=25).tobytes())
    f, r, inter = synth.paired_records(1000, 150, "sanger", seed=26)
    open(pf, "wb").write(f.tobytes()); open(pr, "wb").write(r.tobytes()); open(il, "wb").write(inter.tobytes())
    for threads in (2, 3, 4, 8, 16):
        check(harness, se, ("se", threads), kernels=("general",), threads=threads, first=threads)
        check(harness, il, ("pei", threads), kernels=("general",), mode="pei", threads=threads)
        check(harness, pf, ("pe2", threads), kernels=("general",), mode="pe2", path2=pr, threads=threads)


def test_golden_inputs(harness, golden):
    """The committed fixtures (reference-pinned through the oracle), all three encodings."""
    done = 0
    for case in golden["cases"]:
        if case["threads"] > 1 or case["rc"] != 0 and "err_" not in case["id"]:
            continue
        fl = case["flags"]
        if done % 8 and "err_" not in case["id"] and "ok_" not in case["id"]:
            done += 1
            continue
        done += 1
        kw = dict(qualtype=fl[fl.index("-t") + 1], q=int(fl[fl.index("-q") + 1]) if "-q" in fl else 20,
                  l=int(fl[fl.index("-l") + 1]) if "-l" in fl else 20, x="-x" in fl, n="-n" in fl)
        ins = case["inputs"]
        if case["mode"] == "se":
            check(harness, os.path.join(golden["dir"], ins["-f"]), case["id"], **kw)
        elif "-c" in ins:
            if "-M" in case["outputs"]:
                continue
            check(harness, os.path.join(golden["dir"], ins["-c"]), case["id"], mode="pei", singles="-s" in case["outputs"], **kw)
        else:
            check(harness, os.path.join(golden["dir"], ins["-f"]), case["id"], kernels=("fused5", "fused9", "general"), mode="pe2",
                  path2=os.path.join(golden["dir"], ins["-r"]), singles="-s" in case["outputs"], **kw)
    assert done > 100


def test_read_lengths_from_1_to_12000(harness, tmp_path):
    """Variable-length reads: short ones overflow the records-per-tile limit of the large tiles (FASTFAIL ->
    smaller tile or general path), long ones exceed the halo (general path, whole warp per read)."""
    from test_oracle_fuzz_vs_ref import _records

    rng = np.random.default_rng(8)
    for lmax, nrec in ((12, 2000), (40, 1500), (100, 1200), (250, 1000), (2000, 150), (12000, 24)):
        for qualtype in ("sanger", "illumina", "solexa"):
            p = str(tmp_path / ("l%d_%s.fq" % (lmax, qualtype)))
            open(p, "wb").write(_records(rng, nrec, lmax, qualtype))
            seen = check(harness, p, (lmax, qualtype), qualtype=qualtype, n=lmax == 100, x=lmax == 40, first=int(rng.integers(0, 16)))
            if lmax == 250:
                assert seen[0] == "OK"            # the smallest tile holds them; larger tiles may exceed 128 records
            if lmax == 12000:
                assert seen[:5] == ["FASTFAIL"] * 5


def test_long_reads_config4(harness, tmp_path):
    """BASELINE.json configs[3]: reads of 1-20 kb with -x and -n, Illumina / Solexa encodings, '+name' lines."""
    from sickle_b200 import synth

    for qualtype, seed in (("illumina", 4), ("solexa", 5)):
        p = str(tmp_path / (qualtype + ".fq"))
        open(p, "wb").write(synth.variable_length_records(16, 1000, 20000, qualtype, seed))
        assert check(harness, p, qualtype, kernels=("fused9", "general"), qualtype=qualtype, x=True, n=True, ctas=4, first=3) == ["FASTFAIL", "OK"]
        check(harness, p, (qualtype, "q30"), kernels=("general",), qualtype=qualtype, q=30, l=100, ctas=3)


def test_long_reads_with_out_of_range_quality_bytes(harness, tmp_path):
    """Long reads take the warp-wide path, whose coarse pass screens more bytes than the reference's loop
    visits: a quality byte outside the encoding's range -- at the front, in the middle, near the end, inside
    or beyond the visited prefix -- must still give the reference's verdict (same error record / position /
    byte, or no error at all when the loop never gets there)."""
    from sickle_b200 import synth

    rng = np.random.default_rng(77)
    verdicts = set()
    for case in range(24):
        qualtype = ("sanger", "illumina", "solexa")[case % 3]
        data = bytearray(synth.variable_length_records(6, 1100, 9000, qualtype, 300 + case, p_N=0.0, p_n=0.0))
        lines = bytes(data).split(b"\n")
        rec = int(rng.integers(0, 6))
        qline_start = sum(len(x) + 1 for x in lines[:4 * rec + 3])
        L = len(lines[4 * rec + 3])
        where = (0, 1, L // 20, L // 3, L // 2, L - L // 12, L - 2, L - 1)[case % 8]
        data[qline_start + where] = (0x1f, 0x7f, 0x80, 0xff, 0x20, 0x0b)[case % 6]
        p = str(tmp_path / ("bad%d.fq" % case))
        open(p, "wb").write(bytes(data))
        for x in (False, True):
            rc, out, err = run(harness, p, kernel="general", qualtype=qualtype, x=x, ctas=3, first=case % 16)
            assert rc == 0 and out.startswith("OK"), (case, x, out, err[-400:])
            verdicts.add("error" in out.lower())
    assert verdicts == {True, False}      # some bytes are met by the loop, some lie beyond its break


def test_general_path_in_two_kernels(harness, tmp_path):
    """Batches of long records run K2 as two kernels (k2_trim_only: trimming, every warp on its own; k2_trim_route<true>:
    routing + scan from the stored verdicts) and K3 with one record per warp: long reads, and -- the form must be right
    for any input -- short reads, -a N order, two files, interleaved pairs, -M, a data error."""
    from sickle_b200 import synth

    env = {"KH_K2_SPLIT": "2"}   # units a warp of k2_trim_only draws at a time: 2 (long records) or 32

    def ok(path, **kw):
        rc, out, err = run(harness, path, kernel="general", env=kw.pop("env", env), **kw)
        assert rc == 0 and out.startswith("OK"), (path, kw, out, err[-300:])
        return out

    for qualtype, seed in (("illumina", 14), ("solexa", 15)):
        p = str(tmp_path / (qualtype + ".fq"))
        open(p, "wb").write(synth.variable_length_records(20, 1000, 20000, qualtype, seed))
        ok(p, qualtype=qualtype, x=True, n=True, ctas=4, first=3)
        ok(p, qualtype=qualtype, q=30, l=100, ctas=3)
    se, pf, pr, il = (str(tmp_path / n) for n in ("se.fq", "f.fq", "r.fq", "il.fq"))
    data = synth.fixed_length_records(3000, 150, "sanger", seed=5).tobytes()
    open(se, "wb").write(data)
    f, r, inter = synth.paired_records(1200, 150, "sanger", seed=6)
    open(pf, "wb").write(f.tobytes()); open(pr, "wb").write(r.tobytes()); open(il, "wb").write(inter.tobytes())
    for e in (env, {"KH_K2_SPLIT": "32"}):
        ok(se, ctas=3, env=e)
        ok(se, ctas=3, threads=4, env=e)
        ok(pf, mode="pe2", path2=pr, ctas=3, env=e)
        ok(il, mode="pei", ctas=3, env=e)
        ok(il, mode="peM", singles=False, ctas=2, env=e)
    lines = data.split(b"\n")
    lines[4 * 2000 + 3] = b"\x7f" + lines[4 * 2000 + 3][1:]
    bad = str(tmp_path / "bad.fq")
    open(bad, "wb").write(b"\n".join(lines))
    assert ok(bad, ctas=3).startswith("OK error kind=6 record=2000")


def test_reference_order_on_the_index_pass(harness, tmp_path):
    """-a N on one input (capi.cu's launch_hybrid): the single-pass kernel's S1-S6 as an index + verdict pass
    (kf_fused<CH, 3>: the line index K1 writes and the verdicts k2_trim_only writes, from one read of the input), then
    k2_trim_route<true>, K3 and the summary.  Same bytes as the oracle in the reference's -a N order, single end and
    interleaved pairs (+ -M), every tile size; a data error or a record longer than the halo hands the batch over."""
    from sickle_b200 import synth
    from test_oracle_fuzz_vs_ref import _records

    se, il, var, lng, bad, pf, pr = (str(tmp_path / n) for n in ("se.fq", "il.fq", "var.fq", "long.fq", "bad.fq", "f.fq", "r.fq"))
    data = synth.fixed_length_records(3000, 150, "sanger", seed=5).tobytes()
    open(se, "wb").write(data)
    f2, r2, il2 = synth.paired_records(1200, 150, "sanger", seed=6)
    open(il, "wb").write(il2.tobytes())
    open(pf, "wb").write(f2.tobytes())
    open(pr, "wb").write(r2.tobytes()[:len(r2.tobytes()) * 3 // 4])   # (the second file shorter, ending inside a record)
    open(var, "wb").write(_records(np.random.default_rng(91), 1500, 250, "sanger"))
    open(lng, "wb").write(synth.variable_length_records(12, 6000, 9000, "illumina", 3))
    for k in ("index3", "index5", "index7", "index9"):
        for threads in (1, 2, 3, 8, 32):
            for path, kw in ((se, dict(first=5)), (se, dict(x=True, n=True, q=30, l=5, ctas=2)), (il, dict(mode="pei", first=3)),
                             (il, dict(mode="pei", singles=False)), (il, dict(mode="peM", singles=False, ctas=2)), (var, dict(n=True, first=9)),
                             (pf, dict(mode="pe2", path2=pr, first=7)), (pf, dict(mode="pe2", path2=pr, singles=False, ctas=2))):
                rc, out, err = run(harness, path, kernel=k, threads=threads, **kw)
                assert rc == 0 and (out.startswith("OK") or out.startswith("FASTFAIL")), (k, threads, kw, out, err[-400:])
                if k == "index9" and path != var:
                    assert out.startswith("OK"), (k, threads, kw, out)
    rc, out, err = run(harness, lng, kernel="index9", qualtype="illumina", threads=4)
    assert rc == 0 and out.startswith("FASTFAIL"), out              # records longer than the halo
    lines = data.split(b"\n")
    lines[4 * 2000 + 3] = b"\x7f" + lines[4 * 2000 + 3][1:]
    open(bad, "wb").write(b"\n".join(lines))
    rc, out, err = run(harness, bad, kernel="index9", threads=4)
    assert rc == 0 and out.startswith("FASTFAIL"), out              # a data error: the general path reports it


def test_reference_order_on_the_single_pass_kernel(harness, tmp_path):
    """-a N (N <= 32), single end (capi.cu's launch_ordered): index + verdict pass, kfo_offsets, then the ordered
    emit pass -- a tile staged queue by queue and flushed as up to N segments, every segment's place in the output known
    beforehand.  The oracle's bytes in the reference's -a N order, for every tile size, N from 2 to 32, several CTA
    counts and input phases; hand-over on a data error and on records longer than the halo."""
    from sickle_b200 import synth
    from test_oracle_fuzz_vs_ref import _records

    se, var, lng, bad, big = (str(tmp_path / n) for n in ("se.fq", "var.fq", "long.fq", "bad.fq", "big.fq"))
    data = synth.fixed_length_records(3000, 150, "sanger", seed=5).tobytes()
    open(se, "wb").write(data)
    open(var, "wb").write(_records(np.random.default_rng(92), 1500, 250, "sanger"))
    open(lng, "wb").write(synth.variable_length_records(12, 6000, 9000, "illumina", 3))
    open(big, "wb").write(synth.fixed_length_records(40000, 150, "sanger", seed=37).tobytes())
    for k in ("order3", "order5", "order7", "order9"):
        for threads in (2, 3, 5, 8, 16, 31, 32):
            for path, kw in ((se, dict(first=5)), (se, dict(x=True, n=True, q=30, l=5, ctas=2)), (se, dict(q=41, l=30, ctas=1, first=15)),
                             (var, dict(n=True, first=9, ctas=4))):
                rc, out, err = run(harness, path, kernel=k, threads=threads, **kw)
                assert rc == 0 and (out.startswith("OK") or out.startswith("FASTFAIL")), (k, threads, kw, out, err[-400:])
                if k == "order9" and path != var:
                    assert out.startswith("OK"), (k, threads, kw, out)
    for threads in (2, 8, 32):   # hundreds of tiles over 8 concurrent CTAs
        rc, out, err = run(harness, big, kernel="order9", threads=threads, ctas=8, first=13)
        assert rc == 0 and out.startswith("OK"), (threads, out, err[-400:])
    rc, out, err = run(harness, lng, kernel="order9", qualtype="illumina", threads=4)
    assert rc == 0 and out.startswith("FASTFAIL"), out
    lines = data.split(b"\n")
    lines[4 * 2000 + 3] = b"\x7f" + lines[4 * 2000 + 3][1:]
    open(bad, "wb").write(b"\n".join(lines))
    rc, out, err = run(harness, bad, kernel="order9", threads=4)
    assert rc == 0 and out.startswith("FASTFAIL"), out


def test_damaged_inputs(harness, tmp_path):
    """Seeded random files, three quarters of them damaged (missing / blank / doubled line, flipped /
    deleted / inserted byte, cut-off tail): same first data error (kind, record, position, byte) or same
    bytes as the oracle, on the general path; the fused kernels either agree or hand the batch over."""
    from test_oracle_fuzz_vs_ref import FLAGSETS, _damage, _records

    rng = np.random.default_rng(2718)
    n_err = n_ok = 0
    p = str(tmp_path / "d.fq")
    for case in range(160):
        qualtype = ["sanger", "illumina", "solexa"][case % 3]
        data = _records(rng, int(rng.integers(60, 300)), int(rng.choice([12, 40, 90, 160])), qualtype)
        if case % 4:
            data = _damage(rng, data)
        if data and not data.endswith(b"\n"):
            data = data[:-1] + b"\n"          # (the host patches an unterminated last line before the device sees it)
        fl = FLAGSETS[case % len(FLAGSETS)]
        open(p, "wb").write(data)
        mode = "pei" if case % 5 == 0 else "se"
        kw = dict(mode=mode, qualtype=qualtype, q=fl["q"], l=fl["l"], x=fl["x"], n=fl["n"], first=case % 16, ctas=1 + case % 4)
        rc, out, err = run(harness, p, kernel="general", **kw)
        assert rc == 0 and out.startswith("OK"), (case, out, err[-300:])
        if out.startswith("OK error"):
            n_err += 1
        else:
            n_ok += 1
        for k in ("fused5", "fused9"):
            rc, out2, err = run(harness, p, kernel=k, **kw)
            assert rc == 0, (case, k, out2, err[-300:])
            if out.startswith("OK error"):
                assert out2.startswith("FASTFAIL"), (case, k, out2)       # data errors always go to the general path
    assert n_err > 50 and n_ok > 40, (n_err, n_ok)


def test_edge_inputs(harness, tmp_path):
    rec = b"@r1\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n"
    cases = {"empty": b"", "one": rec, "newlines": b"\n\n\n\n", "tail": rec * 3 + b"@r2\nACGT\n+\n", "unterminated": rec * 2 + b"@r3\nAC",
             "crlf": rec.replace(b"\n", b"\r\n") * 4, "plus_name": rec.replace(b"+\n", b"+r1 something\n") * 50,
             "at_in_quals": b"".join(b"@r%d\nACGTACGTACGTACGTACGTACGT\n+\n@@@@@@@@@@@@++++++++++++\n" % i for i in range(300))}
    for name, data in cases.items():
        p = str(tmp_path / (name + ".fq"))
        open(p, "wb").write(data)
        for mode in ("se", "pei"):
            check(harness, p, (name, mode), mode=mode, first=3)


def test_random_thread_order(harness, tmp_path):
    """SIMT_SHUFFLE: the fibers of every scheduling round run in a random order instead of by thread index.
    Code that only works because a lower-numbered thread happened to run first (a missing barrier, an
    unordered hand-over through shared memory) then produces wrong bytes; the kernels do not care.  The
    second half shows the check has teeth: with one __syncthreads removed the kernel no longer passes."""
    import shutil

    from sickle_b200 import synth
    from test_oracle_fuzz_vs_ref import _records

    se, il, var = (str(tmp_path / n) for n in ("se.fq", "il.fq", "var.fq"))
    open(se, "wb").write(synth.fixed_length_records(2000, 150, "sanger", seed=65).tobytes())
    open(il, "wb").write(synth.paired_records(1000, 150, "sanger", seed=66)[2].tobytes())
    open(var, "wb").write(_records(np.random.default_rng(67), 1500, 250, "sanger"))
    cases = ((se, dict(kernel="fused9", first=5)), (se, dict(kernel="general", first=5)), (il, dict(kernel="fused7", mode="pei", first=3)),
             (il, dict(kernel="fused5", mode="peM", singles=False, ctas=2)), (var, dict(kernel="general", x=True, n=True, ctas=4, first=2)),
             (var, dict(kernel="fused5", n=True, first=5)), (se, dict(kernel="order9", threads=8, first=4)),
             (se, dict(kernel="index7", threads=3, first=2)))
    for seed in ("1", "2", "3"):
        for path, kw in cases:
            rc, out, err = run(harness, path, env={"SIMT_SHUFFLE": seed}, **kw)
            assert rc == 0 and out.startswith("OK"), (seed, kw, out, err[-500:])
    # mutation: drop the barrier that makes the newline positions visible before S5 reads them
    mut = tmp_path / "csrc"
    shutil.copytree(os.path.join(ROOT, "sickle_b200", "csrc"), mut)
    src = (mut / "kf_fused.cuh").read_text()
    barrier = "        if (!kSaved) __syncthreads();   // newline positions visible to every thread"
    assert src.count(barrier) == 1
    (mut / "kf_fused.cuh").write_text(src.replace(barrier, "        // barrier removed:"))
    stub = os.path.join(ROOT, "tests", "host_stub")
    exe = str(tmp_path / "mutated_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-fno-extern-tls-init", "-I" + os.path.join(stub, "simt"),
                           "-I" + os.path.join(stub, "simt", "include"), "-I" + str(mut), "-I" + os.path.join(ROOT, "oracle"),
                           "-x", "c++", os.path.join(stub, "kernels_harness.cpp"), "-x", "c", os.path.join(ROOT, "oracle", "sickle_oracle.c"),
                           "-o", exe, "-lpthread"])
    verdicts = []
    for seed in ("1", "2", "3"):
        rc, out, err = run(exe, se, env={"SIMT_SHUFFLE": seed}, kernel="fused9", first=5)
        verdicts.append(rc == 0 and out.startswith("OK"))
    assert not any(verdicts), verdicts


def test_no_out_of_bounds_access_under_asan(tmp_path):
    """The kernels under AddressSanitizer + UBSan with input and output buffers sized exactly as capi.cu sizes
    them (64 bytes of padding, nothing more): no global read or write outside them on either path, and no
    undefined behaviour (over-wide shifts, signed overflow, misaligned access) that a CPU and a GPU would
    resolve differently."""
    from sickle_b200 import synth
    from test_oracle_fuzz_vs_ref import _records

    exe = build_harness("kernels_harness_asan", extra=("-g", "-fsanitize=address,undefined"))
    se, il, var = (str(tmp_path / n) for n in ("se.fq", "il.fq", "var.fq"))
    open(se, "wb").write(synth.fixed_length_records(2000, 150, "sanger", seed=55).tobytes())
    open(il, "wb").write(synth.paired_records(1000, 150, "sanger", seed=56)[2].tobytes())
    open(var, "wb").write(_records(np.random.default_rng(57), 1500, 250, "sanger"))
    env = {"KH_TIGHT": "1", "ASAN_OPTIONS": "detect_leaks=0"}
    for path, kw in ((se, dict(kernel="fused9", first=0)), (se, dict(kernel="fused7", first=11, ctas=2)), (se, dict(kernel="general", first=7)),
                     (il, dict(kernel="fused9", mode="pei", first=3)), (il, dict(kernel="fused5", mode="peM", singles=False)),
                     (il, dict(kernel="general", mode="pei", first=9)), (var, dict(kernel="fused5", n=True, first=5)),
                     (var, dict(kernel="general", x=True, n=True, ctas=4, first=2)),
                     (se, dict(kernel="order9", threads=8, first=3)), (se, dict(kernel="order5", threads=32, first=14, ctas=2)),
                     (se, dict(kernel="index9", threads=3, first=6)), (il, dict(kernel="index7", mode="pei", threads=5, first=1))):
        rc, out, err = run(exe, path, env=env, **kw)
        assert rc == 0 and out.startswith("OK") and "AddressSanitizer:" not in err and "runtime error" not in err, (kw, out, err[-1500:])


def test_capacity_overflows_are_flagged(harness, tmp_path):
    """A line index or an output buffer that is too small is reported (capi.cu turns the flags into
    SK_E_CAPACITY), never written past: the harness allocates the full size and tells the kernels less."""
    from sickle_b200 import synth

    se = str(tmp_path / "se.fq")
    open(se, "wb").write(synth.fixed_length_records(2000, 150, "sanger", seed=45).tobytes())
    rc, out, err = run(harness, se, kernel="general", env={"KH_LINE_CAP": "1000"})
    assert rc == 0 and out == "OVERFLOW line_index=1 output=0", (out, err[-300:])
    for k, threads in (("general", 1), ("fused9", 1), ("fused5", 1), ("order9", 4), ("index9", 4)):
        rc, out, err = run(harness, se, kernel=k, threads=threads, env={"KH_OUT_CAP": "100000"})
        assert rc == 0 and out == "OVERFLOW line_index=0 output=1", (k, out, err[-300:])
        rc, out, err = run(harness, se, kernel=k, threads=threads, env={"KH_OUT_CAP": "560000"})      # 558 KB of output: just fits
        assert rc == 0 and out.startswith("OK"), (k, out)
