"""The drop-in command line (bin/sickle, C++ host over the C ABI) against the reference's outputs.

Runs the golden.json cases through `sickle se|pe` exactly as the reference was run to produce them
(same flags, `-a N` included) and compares output files (md5), the summary counters on stdout, and
-- for the error fixtures -- exit code and stderr text.  Also: gzip input, -g output, -M.
"""
import gzip
import hashlib
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "sickle")


@pytest.fixture(scope="module")
def sickle():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-s", "-C", ROOT, "cli"])
    return BIN


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest() if os.path.exists(path) else None


def case_args(case, gdir, tmp, tag=""):
    """Arguments of one golden case (after the program name) and its output paths."""
    args = [case["mode"]]
    for k, v in case["inputs"].items():
        args += [k, os.path.join(gdir, v)]
    outs = {}
    for k in case["outputs"]:
        outs[k] = os.path.join(tmp, "out" + tag + k.strip("-"))
        args += [k, outs[k]]
    args += case["flags"]
    if case["threads"] > 1:
        args += ["-a", str(case["threads"])]
    return args, outs


def run_case(sickle, case, gdir, tmp, extra_env=None, tag=""):
    args, outs = case_args(case, gdir, tmp, tag)
    cmd = [sickle] + args
    env = dict(os.environ, SICKLE_B200_SLOT_MB="1", SICKLE_B200_HEADROOM_MB="1")
    env.update(extra_env or {})
    p = subprocess.run(cmd, capture_output=True, timeout=120, env=env)
    return p, outs


def counts(stdout):
    pats = {"kept": r"FastQ records kept: (\d+)", "discard": r"FastQ records discarded: (\d+)",
            "total": r"Total FastQ records: (\d+)", "kept_p": r"FastQ paired records kept: (\d+)",
            "discard_p": r"FastQ paired records discarded: (\d+)", "kept_s": r"FastQ single records kept: (\d+)",
            "discard_s": r"FastQ single records discarded: (\d+)"}
    return {k: int(m.group(1)) for k, pat in pats.items() if (m := re.search(pat, stdout))}


def test_golden_cases_through_cli(sickle, golden, tmp_path):
    from concurrent.futures import ThreadPoolExecutor

    # every seventh flag set plus all -a N and all edge/error cases keeps the run short; each case is a
    # fresh process (~2 s of CUDA start-up), so four run side by side
    picked = [(i, c) for i, c in enumerate(golden["cases"])
              if i % 7 == 0 or c["threads"] > 1 or "err_" in c["id"] or "ok_" in c["id"]]
    with ThreadPoolExecutor(max_workers=4) as pool:
        results = list(pool.map(lambda ic: run_case(sickle, ic[1], golden["dir"], str(tmp_path), tag="%d" % ic[0]), picked))
    bad = []
    for (i, case), (p, outs) in zip(picked, results):
        if p.returncode != case["rc"]:
            bad.append((case["id"], "rc", p.returncode, p.stderr[-300:]))
            continue
        if case["rc"] == 0:
            for k, o in case["outputs"].items():
                if md5(outs[k]) != o["md5"]:
                    bad.append((case["id"], "md5", k, os.path.getsize(outs[k]), o["bytes"]))
            got = counts(p.stdout.decode())
            for k, v in case["counts"].items():
                if k in got and k != "total" and got[k] != v:
                    bad.append((case["id"], "count", k, got[k], v))
            if case["mode"] == "se" and got.get("total") != case["counts"]["total"]:
                bad.append((case["id"], "total", got.get("total"), case["counts"]["total"]))
        elif p.stderr.decode("latin-1") != case["stderr"]:
            bad.append((case["id"], "stderr", p.stderr.decode("latin-1")[:400], case["stderr"][:400]))
    assert not bad, bad[:6]
    assert len(picked) > 50


def test_gzip_input_and_output(sickle, golden, tmp_path):
    src = os.path.join(golden["dir"], "se_r150.fastq")
    gz_in = str(tmp_path / "in.fastq.gz")
    with open(src, "rb") as f, gzip.open(gz_in, "wb") as g:
        g.write(f.read())
    plain, gz_out, gz_out2 = str(tmp_path / "p.fq"), str(tmp_path / "o.fq.gz"), str(tmp_path / "o2.fq.gz")
    subprocess.run([sickle, "se", "-f", src, "-t", "sanger", "-o", plain], check=True, capture_output=True)
    subprocess.run([sickle, "se", "-f", gz_in, "-t", "sanger", "-o", gz_out, "-g"], check=True, capture_output=True)
    subprocess.run([sickle, "se", "-f", src, "-t", "sanger", "-o", gz_out2, "-g"], check=True, capture_output=True)
    want = open(plain, "rb").read()
    case = [c for c in golden["cases"] if c["id"] == "se.se_r150.sanger.default"][0]
    assert hashlib.md5(want).hexdigest() == case["outputs"]["-o"]["md5"]
    assert gzip.open(gz_out, "rb").read() == want
    assert gzip.open(gz_out2, "rb").read() == want
    # blocked gzip (BGZF, what -g writes and what bgzip / Illumina's converters produce) as input:
    # blocks are inflated in parallel; the records come out the same
    from sickle_b200 import synth

    big = synth.fixed_length_records(60000, 150, "sanger", seed=21).tobytes()
    big_plain, big_gz, o1, o2 = (str(tmp_path / n) for n in ("big.fq", "big.fq.gz", "big1.fq", "big2.fq"))
    open(big_plain, "wb").write(big)
    subprocess.run([os.path.join(os.path.dirname(sickle), "io_tool"), big_plain, big_gz, str(1 << 22), "1"], check=True,
                   capture_output=True)
    subprocess.run([sickle, "se", "-f", big_plain, "-t", "sanger", "-o", o1], check=True, capture_output=True)
    p = subprocess.run([sickle, "se", "-f", big_gz, "-t", "sanger", "-o", o2, "-d"], check=True, capture_output=True)
    assert open(o1, "rb").read() == open(o2, "rb").read()
    assert b"(gzip)" in p.stderr


def test_M_mode_matches_oracle(sickle, golden, tmp_path):
    """-M (parity unpinned: absent from the reference fork) against the oracle's README-derived rule."""
    import oracle_py as orc

    src = os.path.join(golden["dir"], "il15_inter.fastq")
    out = str(tmp_path / "m.fq")
    p = subprocess.run([sickle, "pe", "-c", src, "-t", "illumina", "-M", out], capture_output=True)
    assert p.returncode == 0, p.stderr
    data = open(src, "rb").read()
    want = orc.run(orc.MODE_PE_INTER_M, orc.make_params("illumina"), data, batch_len=1 << 40)
    got = open(out, "rb").read()
    assert got == want["out"][0]
    assert got.count(b"\n") == data.count(b"\n")          # every record is still there
    assert b"\nN\n+\n@\n" in got or b"\nN\n+" in got       # some N records


def test_usage_and_argument_errors(sickle, tmp_path):
    p = subprocess.run([sickle], capture_output=True)
    assert p.returncode == 1 and b"Usage: sickle <command> [options]" in p.stdout
    p = subprocess.run([sickle, "se", "-t", "sanger"], capture_output=True)
    assert p.returncode == 1 and b"****Error: Must have quality type, input file, and output file." in p.stderr
    p = subprocess.run([sickle, "se", "-f", "a", "-o", "b", "-t", "phred"], capture_output=True)
    assert p.returncode == 1 and b"Error: Quality type 'phred' is not a valid type." in p.stderr
    p = subprocess.run([sickle, "se", "-f", "a", "-o", "a", "-t", "sanger"], capture_output=True)
    assert p.returncode == 1 and b"****Error: Input file is same as output file." in p.stderr
    p = subprocess.run([sickle, "se", "-f", "a", "-o", "b", "-t", "sanger", "-q", "-3"], capture_output=True)
    assert p.returncode == 1 and b"Quality threshold must be >= 0" in p.stderr
    p = subprocess.run([sickle, "pe", "-t", "sanger"], capture_output=True)
    assert p.returncode == 1 and b"****Error: Must have either -f OR -c argument." in p.stderr
    p = subprocess.run([sickle, "pe", "-f", "a", "-t", "sanger"], capture_output=True)
    assert p.returncode == 1 and b"you must have the -r, -o, -p, and -s options" in p.stderr
    p = subprocess.run([sickle, "--version"], capture_output=True)
    assert p.returncode == 0 and b"sickle version 1.33" in p.stdout


def test_batch_mode_runs_commands_in_one_process(sickle, golden, tmp_path):
    """`sickle batch`: one command per stdin line, context reused; outputs equal those of separate
    processes, a data error in one command is reported and does not disturb the next ones."""
    g = golden["dir"]
    se, inter, bad = (os.path.join(g, n) for n in ("se_r150.fastq", "pe_r150_inter.fastq", "err_len_mismatch.fastq"))
    f, r = os.path.join(g, "pe_r150_f.fastq"), os.path.join(g, "pe_r150_r.fastq")
    spaced = tmp_path / "with space"
    spaced.mkdir()
    o = {k: str(tmp_path / k) for k in ("a", "b", "bs", "c", "d", "e1", "e2", "es")}
    o["sp"] = str(spaced / "out file.fq")
    lines = ["se -f %s -t sanger -o %s" % (se, o["a"]),
             "pe -c %s -t sanger -m %s -s %s" % (inter, o["b"], o["bs"]),
             "se -f %s -t sanger -o %s" % (bad, o["c"]),
             "se -f %s -t sanger -o %s -q 30" % (se, o["d"]),
             "pe -f %s -r %s -t sanger -o %s -p %s -s %s" % (f, r, o["e1"], o["e2"], o["es"]),
             'se -f %s -t sanger -o "%s"' % (se, o["sp"])]
    p = subprocess.run([sickle, "batch"], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=300)
    rcs = [int(l.split()[1]) for l in p.stdout.splitlines() if l.startswith("##rc ")]
    assert rcs == [0, 0, 1, 0, 0, 0], (rcs, p.stderr[-500:])
    assert p.returncode == 1 and "different lengths" in p.stderr
    singles = {"a": ["se", "-f", se, "-t", "sanger", "-o"], "d": ["se", "-f", se, "-t", "sanger", "-q", "30", "-o"]}
    for k, cmd in singles.items():
        ref = str(tmp_path / ("ref_" + k))
        subprocess.run([sickle] + cmd + [ref], check=True, capture_output=True)
        assert md5(o[k]) == md5(ref), k
    assert md5(o["sp"]) == md5(o["a"])
    ref = [str(tmp_path / n) for n in ("rb", "rbs", "r1", "r2", "rs")]
    subprocess.run([sickle, "pe", "-c", inter, "-t", "sanger", "-m", ref[0], "-s", ref[1]], check=True, capture_output=True)
    subprocess.run([sickle, "pe", "-f", f, "-r", r, "-t", "sanger", "-o", ref[2], "-p", ref[3], "-s", ref[4]], check=True,
                   capture_output=True)
    assert [md5(o[k]) for k in ("b", "bs", "e1", "e2", "es")] == [md5(x) for x in ref]


# ---------------------------------------------------------------------------------------------
# several devices (SICKLE_B200_DEVICES): independent whole-record batches dealt to one context per
# device, outputs appended in batch order.  "0,0" = two contexts on one GPU, so the dealing, the
# host-side cutting (host/unit_cutter.h) and the ordered collection run against the real library on
# a one-GPU box; real device pairs are added when the box has them.  Every process pays 2-3 s of CUDA
# start-up, so the commands of one configuration go through one `sickle batch` process.  (The host
# logic itself is covered case by case on CPU in tests/test_host_logic.py.)
# ---------------------------------------------------------------------------------------------
def run_batch(sickle, commands, env):
    """Run argument lists through one `sickle batch` process.  Returns ([(rc, stdout of the command)], stderr)."""
    text = "".join(" ".join('"%s"' % a for a in args) + "\n" for args in commands)
    p = subprocess.run([sickle, "batch"], input=text.encode(), capture_output=True, timeout=600, env=dict(os.environ, **env))
    res, cur = [], []
    for line in p.stdout.decode().splitlines():
        if line.startswith("##rc "):
            res.append((int(line.split()[1]), "\n".join(cur)))
            cur = []
        else:
            cur.append(line)
    assert len(res) == len(commands), (len(res), len(commands), p.stderr[-600:])
    return res, p.stderr


def _device_lists():
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "sickle_b200", "libsickle_b200.so"))
    n = lib.sk_device_count()
    lists = [("0,0", "256"), ("0,0,0", "1000")]
    if n >= 2:
        lists.append(("0,1", "256"))
    if n >= 4:
        lists.append(("3,1,0,2", "512"))
    return lists


def test_several_devices_golden_cases(sickle, golden, tmp_path):
    """Golden cases (all -a N ones, all error fixtures, every fourth of the rest) with the input cut into
    64 KiB batches over two contexts: same files, counters, exit codes and messages as the reference."""
    picked = [c for i, c in enumerate(golden["cases"]) if i % 4 == 0 or c["threads"] > 1 or "err_" in c["id"] or "ok_" in c["id"]]
    assert len(picked) > 70
    cmds, outs = [], []
    for i, case in enumerate(picked):
        a, o = case_args(case, golden["dir"], str(tmp_path), tag="%d" % i)
        cmds.append(a)
        outs.append(o)
    res, stderr = run_batch(sickle, cmds, {"SICKLE_B200_DEVICES": "0,0", "SICKLE_B200_SLOT_KB": "64"})
    bad = []
    for case, o, (rc, out) in zip(picked, outs, res):
        if rc != case["rc"]:
            bad.append((case["id"], "rc", rc))
        elif rc == 0:
            for k, want in case["outputs"].items():
                if md5(o[k]) != want["md5"]:
                    bad.append((case["id"], "md5", k, os.path.getsize(o[k]), want["bytes"]))
            got = counts(out)
            for k, v in case["counts"].items():
                if k in got and k != "total" and got[k] != v:
                    bad.append((case["id"], "count", k, got[k], v))
    assert not bad, (bad[:6], stderr[-400:])
    # the error fixtures' messages, in the order the commands ran
    assert stderr.decode("latin-1") == "".join(c["stderr"] for c in picked if c["rc"] != 0)


def test_several_devices_many_batches(sickle, tmp_path):
    """Larger synthetic inputs, tens of batches in flight over 2-4 contexts: single end, interleaved
    pairs (+ singles, and -M), two files whose mates differ in length, -a 3 with the reference's
    batches; against whole-input runs of the CPU oracle."""
    import oracle_py as orc
    from sickle_b200 import synth

    se = synth.fixed_length_records(30000, 150, "sanger", seed=41).tobytes()
    f1, f2, inter = (a.tobytes() for a in synth.paired_records(12000, 150, "sanger", seed=42))
    # mate 2 cut to 90 bases: the two files then hold different numbers of records per megabyte
    recs = f2.split(b"\n")
    short = []
    for r in range(0, len(recs) - 1, 4):
        short += [recs[r], recs[r + 1][:90], recs[r + 2], recs[r + 3][:90]]
    f2s = b"\n".join(short) + b"\n"
    paths = {k: str(tmp_path / (k + ".fq")) for k in ("se", "f1", "f2s", "inter")}
    for k, d in (("se", se), ("f1", f1), ("f2s", f2s), ("inter", inter)):
        open(paths[k], "wb").write(d)
    pr = orc.make_params("sanger")
    want_se = orc.run(orc.MODE_SE, pr, se)
    want_il = orc.run(orc.MODE_PE_INTER, pr, inter, batch_len=1 << 40)
    want_m = orc.run(orc.MODE_PE_INTER_M, pr, inter, batch_len=1 << 40)
    want_2f = orc.run(orc.MODE_PE_2FILE, pr, f1, f2s, batch_len=1 << 40)
    want_a3 = orc.run(orc.MODE_SE, pr, se, threads=3, b_mib=1)
    assert want_2f["rc"] == 0 and want_il["rc"] == 0 and want_se["rc"] == 0 and want_a3["counters"]["n_batches"] > 5
    o = lambda name: str(tmp_path / name)
    cmds = [["se", "-f", paths["se"], "-t", "sanger", "-o", o("se.out"), "-d"],
            ["pe", "-c", paths["inter"], "-t", "sanger", "-m", o("il.out"), "-s", o("il.s")],
            ["pe", "-f", paths["f1"], "-r", paths["f2s"], "-t", "sanger", "-o", o("p1"), "-p", o("p2"), "-s", o("ps")],
            ["pe", "-c", paths["inter"], "-t", "sanger", "-M", o("m.out")],
            ["se", "-f", paths["se"], "-t", "sanger", "-o", o("a3.out"), "-a", "3", "-b", "1", "-d"]]
    want = {"se.out": want_se["out"][0], "il.out": want_il["out"][0], "il.s": want_il["out"][2], "p1": want_2f["out"][0],
            "p2": want_2f["out"][1], "ps": want_2f["out"][2], "m.out": want_m["out"][0], "a3.out": want_a3["out"][0]}
    for devs, slot_kb in _device_lists():
        for name in want:
            if os.path.exists(o(name)):
                os.remove(o(name))
        res, stderr = run_batch(sickle, cmds, {"SICKLE_B200_DEVICES": devs, "SICKLE_B200_SLOT_KB": slot_kb})
        assert [rc for rc, _ in res] == [0] * len(cmds), (devs, stderr[-600:])
        for name, data in want.items():
            assert open(o(name), "rb").read() == data, (devs, slot_kb, name)
        assert counts(res[0][1])["kept"] == want_se["counters"]["kept"]
        assert counts(res[2][1])["kept_p"] == want_2f["counters"]["kept_p"]
        nb = [int(x) for x in re.findall(rb"batches (\d+)", stderr)]
        assert nb[0] >= len(se) // (int(slot_kb) << 10) and nb[1] == want_a3["counters"]["n_batches"], (devs, nb)


def _nth_newline(data, n):
    pos = -1
    for _ in range(n):
        pos = data.index(b"\n", pos + 1)
    return pos


def test_two_files_of_unequal_length(sickle, tmp_path):
    """`pe -f -r` where one file holds fewer records than the other and ends inside a record: the pairs that exist
    are trimmed, the rest is dropped (the reference refuses such input) -- in particular a full slot
    of the longer file next to the other file's cut-off tail is not "a record that does not fit"."""
    import oracle_py as orc
    from sickle_b200 import synth

    f, r, _ = synth.paired_records(20000, 150, "sanger", seed=33)
    fb, rb = f.tobytes(), r.tobytes()
    o = lambda name: str(tmp_path / name)
    # (all three cuts fall inside a name or base line; one inside a quality line is a record with lines of
    #  different lengths, an error here as in the reference)
    for tag, a, b in (("b_short", fb, rb[:len(rb) * 2 // 3 + 11]), ("a_half", fb[:len(fb) // 2], rb), ("a_cut", fb[:_nth_newline(fb, 4 * 6001 + 1) + 60], rb)):
        npairs = min(a.count(b"\n"), b.count(b"\n")) // 4
        cut = lambda d: b"\n".join(d.split(b"\n")[:4 * npairs]) + b"\n"
        want = orc.run(orc.MODE_PE_2FILE, orc.make_params("sanger"), cut(a), cut(b), batch_len=1 << 40)
        assert want["rc"] == 0
        open(o("a.fq"), "wb").write(a)
        open(o("b.fq"), "wb").write(b)
        for env in ({"SICKLE_B200_SLOT_MB": "1"}, {"SICKLE_B200_DEVICES": "0,0", "SICKLE_B200_SLOT_KB": "512"}):
            p = subprocess.run([sickle, "pe", "-f", o("a.fq"), "-r", o("b.fq"), "-t", "sanger", "-o", o("p1"), "-p", o("p2"), "-s", o("ps")],
                               capture_output=True, env=dict(os.environ, **env), timeout=120)
            assert p.returncode == 0, (tag, env, p.stderr[-400:])
            for name, k in (("p1", 0), ("p2", 1), ("ps", 2)):
                assert open(o(name), "rb").read() == want["out"][k], (tag, env, name)
            assert counts(p.stdout.decode())["kept_p"] == want["counters"]["kept_p"]


def test_several_devices_errors(sickle, tmp_path):
    """A data error in a late batch: exit 1 with the reference's message and the record's true number
    (same text as with one context); a record that does not fit a slot; an unusable device number."""
    from sickle_b200 import synth

    lines = synth.fixed_length_records(6000, 150, "sanger", seed=43).tobytes().split(b"\n")
    rec = 5000
    lines[4 * rec + 3] = b"\x7f" + lines[4 * rec + 3][1:]   # 127 > Sanger's maximum (126)
    src, src2, src3 = (str(tmp_path / n) for n in ("bad.fq", "bad2.fq", "long.fq"))
    open(src, "wb").write(b"\n".join(lines))
    lines[4 * rec + 3] = lines[4 * rec + 3][:100]            # and a quality line shorter than its sequence
    open(src2, "wb").write(b"\n".join(lines))
    long_rec = b"@x\n" + b"A" * 70000 + b"\n+\n" + b"I" * 70000 + b"\n"
    open(src3, "wb").write(long_rec * 4)
    cmds = [["se", "-f", f, "-t", "sanger", "-o", str(tmp_path / "o.fq")] for f in (src, src2)]
    res, stderr = run_batch(sickle, cmds, {"SICKLE_B200_DEVICES": "0,0", "SICKLE_B200_SLOT_KB": "128"})
    res1, stderr1 = run_batch(sickle, cmds, {})
    assert [rc for rc, _ in res] == [1, 1] and [rc for rc, _ in res1] == [1, 1]
    assert stderr == stderr1 and b"Quality value (127)" in stderr and lines[4 * rec] in stderr and b"different lengths" in stderr
    cmd = [sickle, "se", "-f", src3, "-t", "sanger", "-o", str(tmp_path / "o.fq")]
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SICKLE_B200_DEVICES="0,0", SICKLE_B200_SLOT_KB="64"), timeout=120)
    assert p.returncode == 1 and b"does not fit" in p.stderr
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SICKLE_B200_DEVICES="0,99"), timeout=120)
    assert p.returncode == 1 and b"not available" in p.stderr
