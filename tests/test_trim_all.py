"""Directory driver (trim_all.py; reference trim_all.py:62-108): file discovery, mate pairing and
output names on CPU (--dry-run), and one real run over a small directory on the GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trim_all  # noqa: E402


def _touch(d, *names):
    for n in names:
        (d / n).write_bytes(b"")


def test_se_plan(tmp_path):
    _touch(tmp_path, "a.fq", "b.fastq", "c.fastq.gz", "notes.txt")
    jobs = trim_all.plan("se", "sanger", str(tmp_path), "/out", 0, 0, "sickle")
    assert [j[0] for j in jobs] == ["a.fq", "b.fastq", "c.fastq.gz"]
    assert jobs[0][1] == ["sickle", "se", "-t", "sanger", "-f", str(tmp_path / "a.fq"), "-o", "/out/a.trim.fastq"]
    assert jobs[2][2] == ["/out/c.trim.fastq"]
    jobs = trim_all.plan("se", "illumina", str(tmp_path), "/out", 8, 256, "sickle")
    assert jobs[0][1][-4:] == ["-a", "8", "-b", "256"]


def test_pe_plan_and_separators(tmp_path):
    d = tmp_path / "dot"
    d.mkdir()
    _touch(d, "s1.1.fq", "s1.2.fq", "x.1.fastq", "x.2.fastq")
    jobs = trim_all.plan("pe", "sanger", str(d), "/o", 0, 0, "sickle")
    assert len(jobs) == 2
    cmd = jobs[0][1]
    assert cmd[cmd.index("-f") + 1].endswith("s1.1.fq") and cmd[cmd.index("-r") + 1].endswith("s1.2.fq")
    assert jobs[0][2] == ["/o/s1.1.trim.fastq", "/o/s1.2.trim.fastq", "/o/s1.s.trim.fastq"]
    u = tmp_path / "underscore"
    u.mkdir()
    _touch(u, "r_1.fastq", "r_2.fastq", "q_1.fastq", "q_2.fastq")
    jobs = trim_all.plan("pe", "sanger", str(u), "/o", 0, 0, "sickle")
    assert [j[2][2] for j in jobs] == ["/o/q_s.trim.fastq", "/o/r_s.trim.fastq"]
    m = tmp_path / "missing"
    m.mkdir()
    _touch(m, "a.1.fq", "b.1.fq", "a.2.fq")
    with pytest.raises(FileNotFoundError):
        trim_all.plan("pe", "sanger", str(m), "/o", 0, 0, "sickle")


def test_dry_run_skips_existing(tmp_path, capsys):
    i, o = tmp_path / "in", tmp_path / "out"
    i.mkdir()
    o.mkdir()
    _touch(i, "a.fq", "b.fq")
    _touch(o, "a.trim.fastq")
    assert trim_all.main(["se", "sanger", str(i), str(o), "--dry-run"]) == 0
    out = capsys.readouterr().out
    assert "a.trim.fastq already exists" in out and "b.fq" in out and out.count("\t> ") == 1


@pytest.mark.gpu
def test_directory_run_matches_single_runs(tmp_path, golden):
    import shutil

    i, o = tmp_path / "in", tmp_path / "out"
    i.mkdir()
    src = os.path.join(golden["dir"], "se_r150.fastq")
    for k in range(3):
        shutil.copy(src, i / ("lane%d.fastq" % k))
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "trim_all.py"), "se", "sanger", str(i), str(o)])
    assert rc == 0
    # against the reference binary's own output for this input and these flags (tests/golden/golden.json)
    import hashlib

    case = [c for c in golden["cases"] if c["id"] == "se.se_r150.sanger.default"][0]
    for k in range(3):
        assert hashlib.md5((o / ("lane%d.trim.fastq" % k)).read_bytes()).hexdigest() == case["outputs"]["-o"]["md5"]


@pytest.mark.gpu
def test_directory_run_paired(tmp_path, golden):
    import shutil

    i, o = tmp_path / "in", tmp_path / "out"
    i.mkdir()
    for k in ("a", "b"):
        shutil.copy(os.path.join(golden["dir"], "pe_r150_f.fastq"), i / ("%s_1.fastq" % k))
        shutil.copy(os.path.join(golden["dir"], "pe_r150_r.fastq"), i / ("%s_2.fastq" % k))
    assert subprocess.call([sys.executable, os.path.join(ROOT, "trim_all.py"), "pe", "sanger", str(i), str(o), "--procs-per-gpu", "2"]) == 0
    case = [c for c in golden["cases"] if c["id"] == "pe2.pe_r150.sanger.default"]
    import hashlib

    for k in ("a", "b"):
        got = [hashlib.md5((o / ("%s_%s.trim.fastq" % (k, x))).read_bytes()).hexdigest() for x in ("1", "2", "s")]
        if case:
            assert got == [case[0]["outputs"][f]["md5"] for f in ("-o", "-p", "-s")]
    # a second run finds every output in place and does nothing
    p = subprocess.run([sys.executable, os.path.join(ROOT, "trim_all.py"), "pe", "sanger", str(i), str(o)], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.count("already exists") == 2 and "> " not in p.stdout
