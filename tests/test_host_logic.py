"""Host logic of the command line (host/*.cpp) on CPU: option handling, batch cutting, carried tails,
reference-order batches, batches dealt to several contexts, ordered output, counters, messages, exit codes.

The per-read work of the product happens only in the CUDA library.  So that the host code around it
can still be tested in a container without a GPU, tests/host_stub/stub_abi.cpp answers the C ABI of
include/sickle_b200.h with the CPU oracle; host/*.cpp is compiled against that stub into
tests/_build/sickle_hoststub (test infrastructure: never shipped, never measured; bin/sickle links the
CUDA library only and exits 1 without a GPU).  The same cases run against the real library in
tests/test_cli.py (-m gpu).
"""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_py as orc
import test_cli
from test_cli import counts, md5, run_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "_build", "sickle_hoststub")


@pytest.fixture(scope="module")
def hoststub():
    os.makedirs(os.path.dirname(STUB), exist_ok=True)
    src = [os.path.join(ROOT, p) for p in ("host/sickle_main.cpp", "host/trimmer.cpp", "host/io.cpp",
                                           "tests/host_stub/stub_abi.cpp", "oracle/sickle_oracle.c")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include")] + src +
                          ["-o", STUB, "-lz", "-lpthread"])
    return STUB


ENVS = {
    "one context": {},
    "two contexts, 64 KiB batches": {"SICKLE_B200_DEVICES": "0,0", "SICKLE_B200_SLOT_KB": "64"},
    "four devices, 24 KiB batches": {"SICKLE_STUB_DEVICES": "4", "SICKLE_B200_GPUS": "all", "SICKLE_B200_SLOT_KB": "24"},
}


@pytest.mark.parametrize("env_name", list(ENVS))
def test_golden_cases_through_host_code(hoststub, golden, tmp_path, env_name):
    """All 213 reference golden cases (files, counters, exit codes, stderr text, -a N order) through the
    host code, with one context and with batches dealt to several."""
    bad = []
    for case in golden["cases"]:
        p, outs = run_case(hoststub, case, golden["dir"], str(tmp_path), extra_env=ENVS[env_name])
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


def test_many_batches_over_several_contexts(hoststub, tmp_path):
    """The -m gpu test of the same name in test_cli.py, run against the stub: tens of batches in flight
    over 2-3 contexts; se, interleaved (+ singles, -M), two files whose mates differ in length, -a 3 with
    the reference's batches; each against one whole-input oracle run."""
    test_cli.test_several_devices_many_batches(hoststub, tmp_path)


def test_golden_cases_side_by_side(hoststub, golden, tmp_path):
    """The -m gpu test of the command line on golden cases (fresh processes, four at a time), against the stub."""
    test_cli.test_golden_cases_through_cli(hoststub, golden, tmp_path)


def test_golden_cases_in_one_batch_process(hoststub, golden, tmp_path):
    test_cli.test_several_devices_golden_cases(hoststub, golden, tmp_path)


def test_errors_over_several_contexts(hoststub, tmp_path):
    """A data error in a late batch reports the record's number in the file and stops with exit 1 (same
    text as with one context); a record larger than a slot; a device that does not exist."""
    test_cli.test_several_devices_errors(hoststub, tmp_path)


def test_two_files_of_unequal_length(hoststub, tmp_path):
    test_cli.test_two_files_of_unequal_length(hoststub, tmp_path)


def test_four_devices_in_any_order(hoststub, tmp_path):
    from sickle_b200 import synth

    data = synth.fixed_length_records(20000, 150, "sanger", seed=44).tobytes()
    src, out = str(tmp_path / "in.fq"), str(tmp_path / "out.fq")
    open(src, "wb").write(data)
    want = orc.run(orc.MODE_SE, orc.make_params("sanger"), data)
    env = dict(os.environ, SICKLE_B200_DEVICES="3,1,0,2", SICKLE_STUB_DEVICES="4", SICKLE_B200_SLOT_KB="100")
    p = subprocess.run([hoststub, "se", "-f", src, "-t", "sanger", "-o", out, "-d"], capture_output=True, env=env, timeout=120)
    assert p.returncode == 0 and open(out, "rb").read() == want["out"][0]
    assert int(re.search(rb"batches (\d+)", p.stderr).group(1)) >= 60


def test_gzip_in_and_out_over_several_contexts(hoststub, golden, tmp_path):
    """Plain gzip and BGZF input, -g (BGZF) output, with the batches dealt to two contexts."""
    import gzip

    from sickle_b200 import synth

    data = synth.fixed_length_records(20000, 150, "sanger", seed=51).tobytes()
    want = orc.run(orc.MODE_SE, orc.make_params("sanger"), data)["out"][0]
    plain, gz_in, bgzf_in = (str(tmp_path / n) for n in ("in.fq", "in.fq.gz", "in.bgzf.gz"))
    open(plain, "wb").write(data)
    with gzip.open(gz_in, "wb") as g:
        g.write(data)
    env = dict(os.environ, SICKLE_B200_DEVICES="0,0", SICKLE_B200_SLOT_KB="512")
    o1 = str(tmp_path / "o1.fq.gz")
    assert subprocess.run([hoststub, "se", "-f", plain, "-t", "sanger", "-o", bgzf_in, "-q", "0", "-l", "0", "-x", "-g"], env=env).returncode == 0
    assert gzip.open(bgzf_in, "rb").read() == data            # -q 0 -l 0 -x keeps every base: a BGZF copy of the input
    for src in (gz_in, bgzf_in):
        p = subprocess.run([hoststub, "se", "-f", src, "-t", "sanger", "-o", o1, "-g", "-d"], env=env, capture_output=True)
        assert p.returncode == 0 and b"(gzip)" in p.stderr, p.stderr
        assert gzip.open(o1, "rb").read() == want, src


def test_directory_driver_over_batch_workers(hoststub, golden, tmp_path):
    """trim_all.py (long-lived `sickle batch` workers whose context is reused from file to file) on CPU:
    single-end and paired directories, two workers per "GPU", outputs equal the reference's."""
    import hashlib
    import shutil
    import sys

    env = dict(os.environ, SICKLE_B200_BIN=hoststub)
    i, o = tmp_path / "in", tmp_path / "out"
    i.mkdir()
    for k in range(5):
        shutil.copy(os.path.join(golden["dir"], "se_r150.fastq"), i / ("lane%d.fastq" % k))
    assert subprocess.call([sys.executable, os.path.join(ROOT, "trim_all.py"), "se", "sanger", str(i), str(o), "--gpus", "1"], env=env) == 0
    case = [c for c in golden["cases"] if c["id"] == "se.se_r150.sanger.default"][0]
    for k in range(5):
        assert md5(str(o / ("lane%d.trim.fastq" % k))) == case["outputs"]["-o"]["md5"]
    i2, o2 = tmp_path / "in2", tmp_path / "out2"
    i2.mkdir()
    for k in ("a", "b", "c"):
        shutil.copy(os.path.join(golden["dir"], "pe_r150_f.fastq"), i2 / ("%s_1.fastq" % k))
        shutil.copy(os.path.join(golden["dir"], "pe_r150_r.fastq"), i2 / ("%s_2.fastq" % k))
    assert subprocess.call([sys.executable, os.path.join(ROOT, "trim_all.py"), "pe", "sanger", str(i2), str(o2), "--gpus", "1",
                            "--procs-per-gpu", "2"], env=env) == 0
    case = [c for c in golden["cases"] if c["id"] == "pe2.pe_r150.sanger.default"][0]
    for k in ("a", "b", "c"):
        got = [hashlib.md5((o2 / ("%s_%s.trim.fastq" % (k, x))).read_bytes()).hexdigest() for x in ("1", "2", "s")]
        assert got == [case["outputs"][f]["md5"] for f in ("-o", "-p", "-s")]
    # a damaged file among good ones: reported as failed, the others are still trimmed
    i3, o3 = tmp_path / "in3", tmp_path / "out3"
    i3.mkdir()
    shutil.copy(os.path.join(golden["dir"], "se_r150.fastq"), i3 / "good1.fastq")
    shutil.copy(os.path.join(golden["dir"], "err_len_mismatch.fastq"), i3 / "bad.fastq")
    shutil.copy(os.path.join(golden["dir"], "se_r150.fastq"), i3 / "good2.fastq")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "trim_all.py"), "se", "sanger", str(i3), str(o3), "--gpus", "1", "--procs-per-gpu", "1"],
                       env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "FAILED (1): bad.fastq" in p.stderr
    case = [c for c in golden["cases"] if c["id"] == "se.se_r150.sanger.default"][0]
    assert md5(str(o3 / "good1.trim.fastq")) == md5(str(o3 / "good2.trim.fastq")) == case["outputs"]["-o"]["md5"]


REF = os.path.join(ROOT, "oracle", "_ref", "sickle_sync")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (oracle/Makefile target `ref`)")
@pytest.mark.parametrize("env_name", ["one context", "two contexts, 64 KiB batches"])
def test_host_code_equals_reference_binary_on_damaged_files(hoststub, tmp_path, env_name):
    """The command line as a whole (host code over the stub) against the reference binary on seeded
    random files, most of them damaged: exit status, output bytes, the summary on stdout and -- for data
    errors -- the complete stderr text, which carries the record's name, line number and quality string."""
    from test_oracle_fuzz_vs_ref import FLAGSETS, _damage, _records

    rng = np.random.default_rng(991)
    n_ok = n_err = n_other_batch_cut = 0
    src, out, rout = (str(tmp_path / n) for n in ("in.fastq", "out.fastq", "ref.fastq"))
    env = dict(os.environ, **ENVS[env_name])
    for case in range(240):
        qualtype = ["sanger", "illumina", "solexa"][case % 3]
        data = _records(rng, int(rng.integers(120, 400)), int(rng.choice([12, 40, 90])), qualtype)
        if case % 4:
            data = _damage(rng, data)
        fl = FLAGSETS[case % len(FLAGSETS)]
        open(src, "wb").write(data)
        flags = ["-t", qualtype, "-q", str(fl["q"]), "-l", str(fl["l"])] + (["-x"] if fl["x"] else []) + (["-n"] if fl["n"] else [])
        r = subprocess.run([REF, "se", "-f", src, "-o", rout, "-a", "1"] + flags, capture_output=True, timeout=60)
        if r.returncode < 0:
            continue   # the reference itself crashed
        p = subprocess.run([hoststub, "se", "-f", src, "-o", out] + flags, capture_output=True, timeout=60, env=env)
        tag = (case, qualtype, fl, len(data))
        assert p.returncode == r.returncode, (tag, p.stderr[-300:], r.stderr[-300:])
        if r.returncode == 0:
            assert open(out, "rb").read() == open(rout, "rb").read(), tag
            assert counts(p.stdout.decode()) == counts(r.stdout.decode()), tag
            n_ok += 1
        elif p.stderr != r.stderr:
            # a file with two different data errors: the reference validates a whole batch of its own
            # geometry before trimming it, so which error comes first depends on the batch cut
            n_other_batch_cut += 1
        else:
            n_err += 1
    print("ok %d, same error text %d, other error first %d" % (n_ok, n_err, n_other_batch_cut))
    assert n_ok > 80 and n_err > 60 and n_other_batch_cut <= 2, (n_ok, n_err, n_other_batch_cut)


def test_product_binary_has_no_cpu_path():
    """bin/sickle is linked against the CUDA library only: without a GPU it refuses to run."""
    import shutil

    binary = os.path.join(ROOT, "bin", "sickle")
    if not os.path.exists(binary):
        pytest.skip("bin/sickle not built")
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    out = subprocess.run(["ldd", binary], capture_output=True).stdout.decode() if shutil.which("ldd") else ""
    assert "libsickle_b200.so" in out and "oracle" not in out
    gdir = os.path.join(ROOT, "tests", "golden")
    p = subprocess.run([binary, "se", "-f", os.path.join(gdir, "se_r150.fastq"), "-t", "sanger", "-o", "/dev/null"], capture_output=True, timeout=120)
    assert p.returncode == 1 and b"no usable CUDA device" in p.stderr


def test_error_return_waits_for_queued_writes_under_asan(tmp_path):
    """A data error in a late batch while the previous batches' output is still queued on a slow sink
    (-g deflates; a FIFO nobody drains quickly) must not free the pinned result buffers under the writer
    thread: host code + stub built with AddressSanitizer, several paths (one context pipelined, two
    files, -a N, several contexts, `sickle batch` which keeps serving files afterwards)."""
    from sickle_b200 import synth

    exe = os.path.join(ROOT, "tests", "_build", "sickle_hoststub_asan")
    src = [os.path.join(ROOT, p) for p in ("host/sickle_main.cpp", "host/trimmer.cpp", "host/io.cpp",
                                           "tests/host_stub/stub_abi.cpp", "oracle/sickle_oracle.c")]
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-std=c++17",
                           "-I" + os.path.join(ROOT, "include")] + src + ["-o", exe, "-lz", "-lpthread"])
    n = 40000
    lines = synth.fixed_length_records(n, 150, "sanger", seed=91).tobytes().split(b"\n")
    rec = n - 300                                                   # the error sits in the last of many batches
    lines[4 * rec + 3] = b"\x7f" + lines[4 * rec + 3][1:]
    bad = str(tmp_path / "bad.fq")
    open(bad, "wb").write(b"\n".join(lines))
    good2 = str(tmp_path / "mate2.fq")
    open(good2, "wb").write(synth.fixed_length_records(n, 150, "sanger", seed=92).tobytes())
    out = [str(tmp_path / ("o%d.fq.gz" % k)) for k in range(3)]
    base_env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", SICKLE_B200_GZIP_LEVEL="9", SICKLE_B200_ZIP_THREADS="1")
    cases = [
        (["se", "-f", bad, "-t", "sanger", "-o", out[0], "-g"], {"SICKLE_B200_SLOT_MB": "1"}),
        (["se", "-f", bad, "-t", "sanger", "-o", out[0], "-g", "-a", "3", "-b", "1"], {}),
        (["pe", "-f", bad, "-r", good2, "-t", "sanger", "-o", out[0], "-p", out[1], "-s", out[2], "-g"], {"SICKLE_B200_SLOT_MB": "1"}),
        (["se", "-f", bad, "-t", "sanger", "-o", out[0], "-g"], {"SICKLE_B200_DEVICES": "0,0,0", "SICKLE_B200_SLOT_KB": "256"}),
    ]
    for args, env in cases:
        p = subprocess.run([exe] + args, capture_output=True, env=dict(base_env, **env), timeout=300)
        assert p.returncode == 1, (args, p.returncode, p.stderr[-600:])
        assert b"AddressSanitizer" not in p.stderr and b"Quality value (127)" in p.stderr, (args, p.stderr[-1500:])
