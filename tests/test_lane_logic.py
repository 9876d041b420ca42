"""The device function that trims one read (sk::lane_sliding_window, sickle_b200/csrc/trim_lane.cuh --
shared by the single-pass kernel and by K2's short-read path) compiled for the HOST and checked against
the CPU oracle: CPU only, no CUDA.

tests/host_stub/lane_shim/ stands in for the two CUDA headers trim_lane.cuh includes (qualifiers become
empty macros, __funnelshift / __clz / __ffs / __dp4a are restated, and the two lanes that share a read
are two host threads whose __shfl_xor_sync is a rendezvous).  tests/host_stub/lane_harness.cpp then runs
seeded random reads -- 1 to 2500 bases, three quality encodings, -q / -l / -x / -n combinations, good /
bad stretches, decaying and uniform qualities, N / n bases, out-of-range quality bytes, every byte phase
-- through the function with one lane and with two, and compares keep / five / three / range error with
so_sliding_window.  This pins the integer restatement (dp4a window totals, sign-bit masks, the two-lane
split and merge) to the reference's scalar loop without a GPU; the -m gpu parity tests then cover the
same code as compiled by nvcc.  Test infrastructure only.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build", "lane")
HARNESS = os.path.join(ROOT, "tests", "_build", "lane_harness")


def build_harness(src_header, build_dir, out, defines=()):
    os.makedirs(build_dir, exist_ok=True)
    shutil.copy(src_header, os.path.join(build_dir, "trim_lane.cuh"))
    for name in ("sk_device.cuh", "k1_index.cuh"):
        shutil.copy(os.path.join(ROOT, "tests", "host_stub", "lane_shim", name), build_dir)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w"] + ["-D" + d for d in defines] + ["-I" + build_dir, "-I" + os.path.join(ROOT, "oracle"),
                           "-x", "c++", os.path.join(ROOT, "tests", "host_stub", "lane_harness.cpp"),
                           "-x", "c", os.path.join(ROOT, "oracle", "sickle_oracle.c"), "-o", out, "-lpthread"])
    return out


@pytest.fixture(scope="module")
def harness():
    return build_harness(os.path.join(ROOT, "sickle_b200", "csrc", "trim_lane.cuh"), BUILD, HARNESS)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_lane_sliding_window_equals_oracle(harness, seed):
    p = subprocess.run([harness, str(seed), "40000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    words = p.stdout.split()
    assert words[0] == "checked" and int(words[1]) == 40000 and int(words[7]) == 0
    assert int(words[3]) > 10000 and int(words[5]) > 500          # plenty of kept reads and of range errors


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_balanced_lane_split_variant_equals_oracle(tmp_path_factory, seed):
    """-DSK_LANE_SPLIT4 (experimental, off in the shipped build: the two lanes of a read get equal shares
    cut at a multiple of four windows, partial last step) takes the same decisions."""
    d = tmp_path_factory.getbasetemp() / "lane_split4"
    exe = str(d / "lane_harness")
    if not os.path.exists(exe):
        build_harness(os.path.join(ROOT, "sickle_b200", "csrc", "trim_lane.cuh"), str(d), exe, defines=("SK_LANE_SPLIT4",))
    p = subprocess.run([exe, str(seed), "40000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "mismatches 0" in p.stdout, p.stderr[-2000:]


def test_harness_notices_a_wrong_kernel(tmp_path):
    """A one-token change of the device function (the first good window of the third step is skipped)
    must show up as mismatches: the harness is not vacuous."""
    src = open(os.path.join(ROOT, "sickle_b200", "csrc", "trim_lane.cuh")).read()
    assert src.count("if (goodw) {") == 1
    mutated = tmp_path / "trim_lane_mutated.cuh"
    mutated.write_text(src.replace("if (goodw) {", "if (goodw && base != 64u) {"))
    exe = build_harness(str(mutated), str(tmp_path / "lane"), str(tmp_path / "lane_harness"))
    p = subprocess.run([exe, "1", "40000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 1 and "MISMATCH" in p.stderr
