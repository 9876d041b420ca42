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


def run_case(sickle, case, gdir, tmp, extra_env=None):
    cmd = [sickle, case["mode"]]
    for k, v in case["inputs"].items():
        cmd += [k, os.path.join(gdir, v)]
    outs = {}
    for k in case["outputs"]:
        outs[k] = os.path.join(tmp, "out" + k.strip("-"))
        cmd += [k, outs[k]]
    cmd += case["flags"]
    if case["threads"] > 1:
        cmd += ["-a", str(case["threads"])]
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
    bad = []
    n = 0
    for i, case in enumerate(golden["cases"]):
        # every seventh flag set plus all -a N and all edge/error cases keeps the run short
        # (each case is a fresh process: ~2 s of CUDA start-up)
        edge = "err_" in case["id"] or "ok_" in case["id"]
        if not (i % 7 == 0 or case["threads"] > 1 or edge):
            continue
        n += 1
        p, outs = run_case(sickle, case, golden["dir"], str(tmp_path))
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
    assert n > 50


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
