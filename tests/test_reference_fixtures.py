"""BASELINE.json configs[0]: the reference's own bundled fixtures (reproduced under tests/golden/ref/) and the
md5 table its unmodified binary produced on them (BASELINE.md calls it "the parity gate
before any timing counts").  CPU: the oracle.  GPU (-m gpu): the CUDA path through the C ABI, with the
single-pass kernel and with K1/K2/K3, and through the `bin/sickle` command line."""
import gzip
import hashlib
import os
import subprocess

import pytest

import oracle_py as orc
from test_oracle_golden import SURVEY_MD5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref")
IDS = ["%s-%s-%s-a%d" % (r[0], r[2], "".join("%s%s" % kv for kv in r[3].items()) or "default", r[4]) for r in SURVEY_MD5]


def fixture(name):
    return gzip.open(os.path.join(REF, name + ".gz"), "rb").read()


def md5(b):
    return hashlib.md5(b).hexdigest()


def test_copies_are_the_reference_files():
    """md5 of the input == md5 of `se -q 20 -l 20` output (first row): the copy is intact."""
    assert md5(fixture("test.fastq")) == "17960489e277c3f6839d834155fd3b79"
    assert len(fixture("test.f.fastq")) == 441216 and len(fixture("test.r.fastq")) == 441216
    ref = "/root/reference/test"
    if os.path.isdir(ref):
        for n in ("test.fastq", "test.f.fastq", "test.r.fastq"):
            assert fixture(n) == open(os.path.join(ref, n), "rb").read()


@pytest.mark.parametrize("row", SURVEY_MD5, ids=IDS)
def test_md5_table_oracle(row):
    kind, files, qt, kw, threads, want = row
    ins = [fixture(f) for f in files.split(",")]
    r = orc.run({"se": orc.MODE_SE, "pe2": orc.MODE_PE_2FILE, "pei": orc.MODE_PE_INTER}[kind], orc.make_params(qt, **kw),
                ins[0], ins[1] if len(ins) > 1 else b"", threads=threads)
    assert r["rc"] == 0
    for s, w in enumerate(want):
        if w is not None:
            assert md5(r["out"][s]).startswith(w), (s, len(r["out"][s]))


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["auto", "general"])
@pytest.mark.parametrize("row", SURVEY_MD5, ids=IDS)
def test_md5_table_cuda(row, path, monkeypatch):
    from sickle_b200 import capi, runner

    monkeypatch.setenv("SICKLE_B200_PATH", path)
    kind, files, qt, kw, threads, want = row
    ins = [fixture(f) for f in files.split(",")]
    mode = {"se": capi.MODE_SE, "pe2": capi.MODE_PE_2FILE, "pei": capi.MODE_PE_INTER}[kind]
    p = capi.make_params(qt, kw.get("q", 20), kw.get("l", 20), kw.get("x", False), kw.get("n", False), mode=mode,
                         emulate_threads=threads, has_singles=True)
    for slot in ((1 << 20,) if threads > 1 else (1 << 20, 1 << 16)):     # whole file in one batch; ~13 batches
        with capi.Context(p, slot, 1) as ctx:
            if threads > 1:
                r = runner.trim_stream_reference_order(ctx, ins[0], ins[1] if len(ins) > 1 else b"")
            else:
                r = runner.trim_stream(ctx, ins[0], ins[1] if len(ins) > 1 else b"")
        for s, w in enumerate(want):
            if w is not None:
                assert md5(r["out"][s]).startswith(w), (path, slot, s, len(r["out"][s]))


@pytest.mark.gpu
@pytest.mark.parametrize("row", SURVEY_MD5, ids=IDS)
def test_md5_table_command_line(row, tmp_path):
    kind, files, qt, kw, threads, want = row
    exe = os.path.join(ROOT, "bin", "sickle")
    paths = []
    for f in files.split(","):
        p = str(tmp_path / f)
        open(p, "wb").write(fixture(f))
        paths.append(p)
    outs = [str(tmp_path / n) for n in ("o1.fq", "o2.fq", "s.fq")]
    if kind == "se":
        cmd = [exe, "se", "-f", paths[0], "-o", outs[0]]
    elif kind == "pe2":
        cmd = [exe, "pe", "-f", paths[0], "-r", paths[1], "-o", outs[0], "-p", outs[1], "-s", outs[2]]
    else:
        cmd = [exe, "pe", "-c", paths[0], "-m", outs[0], "-s", outs[2]]
    cmd += ["-t", qt]
    if "q" in kw:
        cmd += ["-q", str(kw["q"])]
    if "l" in kw:
        cmd += ["-l", str(kw["l"])]
    if kw.get("x"):
        cmd.append("-x")
    if kw.get("n"):
        cmd.append("-n")
    if threads > 1:
        cmd += ["-a", str(threads)]
    p = subprocess.run(cmd, capture_output=True, timeout=120)
    assert p.returncode == 0, p.stderr[-500:]
    for s, w in enumerate(want):
        if w is not None:
            assert md5(open(outs[s], "rb").read()).startswith(w), (s, cmd)
