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


def _synthetic(tmp_path):
    from sickle_b200 import synth

    se = synth.fixed_length_records(30000, 150, "sanger", seed=41).tobytes()
    f1, f2, inter = (a.tobytes() for a in synth.paired_records(12000, 150, "sanger", seed=42))
    recs = f2.split(b"\n")
    short = []
    for r in range(0, len(recs) - 1, 4):      # mate 2 cut to 90 bases: different records per megabyte in the two files
        short += [recs[r], recs[r + 1][:90], recs[r + 2], recs[r + 3][:90]]
    f2s = b"\n".join(short) + b"\n"
    data = dict(se=se, f1=f1, f2s=f2s, inter=inter)
    paths = {k: str(tmp_path / (k + ".fq")) for k in data}
    for k, d in data.items():
        open(paths[k], "wb").write(d)
    return data, paths


def test_many_batches_over_several_contexts(hoststub, tmp_path):
    """Tens of batches in flight over 2-4 contexts; se, interleaved (+ singles, -M), two files whose mates
    differ in length, -a N with the reference's batches; each against one whole-input oracle run."""
    data, paths = _synthetic(tmp_path)
    pr = orc.make_params("sanger")
    want_se = orc.run(orc.MODE_SE, pr, data["se"])
    want_il = orc.run(orc.MODE_PE_INTER, pr, data["inter"], batch_len=1 << 40)
    want_m = orc.run(orc.MODE_PE_INTER_M, pr, data["inter"], batch_len=1 << 40)
    want_2f = orc.run(orc.MODE_PE_2FILE, pr, data["f1"], data["f2s"], batch_len=1 << 40)
    want_a3 = orc.run(orc.MODE_SE, pr, data["se"], threads=3, b_mib=1)
    o = lambda name: str(tmp_path / name)
    for devs, ndev in (("0,0", 1), ("0,1,2", 3), ("3,1,0,2", 4)):
        for slot_kb in ("256", "1000"):
            env = dict(os.environ, SICKLE_B200_DEVICES=devs, SICKLE_B200_SLOT_KB=slot_kb, SICKLE_STUB_DEVICES=str(ndev))
            tag = (devs, slot_kb)
            p = subprocess.run([hoststub, "se", "-f", paths["se"], "-t", "sanger", "-o", o("se.out"), "-d"], capture_output=True, env=env, timeout=300)
            assert p.returncode == 0, (tag, p.stderr)
            assert open(o("se.out"), "rb").read() == want_se["out"][0], tag
            assert counts(p.stdout.decode())["kept"] == want_se["counters"]["kept"]
            assert int(re.search(rb"batches (\d+)", p.stderr).group(1)) >= len(data["se"]) // (int(slot_kb) << 10), tag
            p = subprocess.run([hoststub, "pe", "-c", paths["inter"], "-t", "sanger", "-m", o("il.out"), "-s", o("il.s")], capture_output=True, env=env, timeout=300)
            assert p.returncode == 0, (tag, p.stderr)
            assert open(o("il.out"), "rb").read() == want_il["out"][0] and open(o("il.s"), "rb").read() == want_il["out"][2], tag
            p = subprocess.run([hoststub, "pe", "-f", paths["f1"], "-r", paths["f2s"], "-t", "sanger", "-o", o("p1"), "-p", o("p2"), "-s", o("ps")],
                               capture_output=True, env=env, timeout=300)
            assert p.returncode == 0, (tag, p.stderr)
            for k, name in ((0, "p1"), (1, "p2"), (2, "ps")):
                assert open(o(name), "rb").read() == want_2f["out"][k], (tag, name)
            got = counts(p.stdout.decode())
            assert got["kept_p"] == want_2f["counters"]["kept_p"] and got["discard_p"] == want_2f["counters"]["discard_p"]
        env = dict(os.environ, SICKLE_B200_DEVICES=devs, SICKLE_B200_SLOT_KB="300", SICKLE_STUB_DEVICES=str(ndev))
        p = subprocess.run([hoststub, "pe", "-c", paths["inter"], "-t", "sanger", "-M", o("m.out")], capture_output=True, env=env, timeout=300)
        assert p.returncode == 0 and open(o("m.out"), "rb").read() == want_m["out"][0], devs
        # -a 3 -b 1: the reference's batches (1 MiB limit -> file/8 = 1.2 MB -> 1 MiB), dealt to the devices
        p = subprocess.run([hoststub, "se", "-f", paths["se"], "-t", "sanger", "-o", o("a3.out"), "-a", "3", "-b", "1", "-d"], capture_output=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr
        assert open(o("a3.out"), "rb").read() == want_a3["out"][0], devs
        assert int(re.search(rb"batches (\d+)", p.stderr).group(1)) == want_a3["counters"]["n_batches"] > 5


def test_errors_over_several_contexts(hoststub, tmp_path):
    """A data error in a late batch reports the record's number in the file and stops with exit 1 (same
    text as with one context); a record larger than a slot; a device that does not exist."""
    from sickle_b200 import synth

    lines = synth.fixed_length_records(6000, 150, "sanger", seed=43).tobytes().split(b"\n")
    rec = 5000
    lines[4 * rec + 3] = b"\x7f" + lines[4 * rec + 3][1:]   # 127 > Sanger's maximum (126)
    src = str(tmp_path / "bad.fq")
    open(src, "wb").write(b"\n".join(lines))
    env = dict(os.environ, SICKLE_B200_DEVICES="0,0", SICKLE_B200_SLOT_KB="128")
    cmd = [hoststub, "se", "-f", src, "-t", "sanger", "-o", str(tmp_path / "o.fq")]
    p = subprocess.run(cmd, capture_output=True, env=env, timeout=120)
    single = subprocess.run(cmd, capture_output=True, timeout=120)
    assert p.returncode == 1 and single.returncode == 1
    assert p.stderr == single.stderr and b"Quality value (127)" in p.stderr and lines[4 * rec] in p.stderr
    # the same damage as an empty sequence line: the message carries no record text, only the exit code
    lines[4 * rec + 3] = lines[4 * rec + 3][:100]
    open(src, "wb").write(b"\n".join(lines))
    p = subprocess.run(cmd, capture_output=True, env=env, timeout=120)
    assert p.returncode == 1 and b"different lengths" in p.stderr
    long_rec = b"@x\n" + b"A" * 70000 + b"\n+\n" + b"I" * 70000 + b"\n"
    open(src, "wb").write(long_rec * 4)
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SICKLE_B200_DEVICES="0,0", SICKLE_B200_SLOT_KB="64"), timeout=120)
    assert p.returncode == 1 and b"does not fit" in p.stderr
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SICKLE_B200_DEVICES="0,0", SICKLE_B200_SLOT_KB="256"), timeout=120)
    assert p.returncode == 0 and open(str(tmp_path / "o.fq"), "rb").read() == long_rec * 4
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SICKLE_B200_DEVICES="0,99"), timeout=120)
    assert p.returncode == 1 and b"not available" in p.stderr


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
