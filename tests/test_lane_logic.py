"""The device function that trims one read (sk::lane_sliding_window, sickle_b200/csrc/trim_lane.cuh --
shared by the single-pass kernel and by K2's short-read path) compiled for the HOST and checked against
the CPU oracle: CPU only, no CUDA.

tests/host_stub/lane_shim/ stands in for the two CUDA headers trim_lane.cuh includes (qualifiers become
empty macros, __funnelshift / __clz / __ffs / __dp4a are restated).  tests/host_stub/lane_harness.cpp then runs
seeded random reads -- 1 to 2500 bases, three quality encodings, -q / -l / -x / -n combinations, good /
bad stretches, decaying and uniform qualities, N / n bases, out-of-range quality bytes, every byte phase
-- through the function and compares keep / five / three / range error with so_sliding_window (the
function may also *decline* a read with an out-of-range quality byte anywhere in it: its callers then use
the exact warp-wide path).  This pins the integer restatement (word-granular lower bounds, dp4a window
totals, sign-bit masks) to the reference's scalar loop without a GPU; the -m gpu parity tests then cover the
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


def test_harness_notices_a_wrong_kernel(tmp_path):
    """A one-token change of the device function (the exact scan starts one window after the last one the
    word-granular bound vouches for) must show up as mismatches: the harness is not vacuous."""
    src = open(os.path.join(ROOT, "sickle_b200", "csrc", "trim_lane.cuh")).read()
    assert src.count("(int)Q - 3;") == 1
    mutated = tmp_path / "trim_lane_mutated.cuh"
    mutated.write_text(src.replace("(int)Q - 3;", "(int)Q - 2;"))
    exe = build_harness(str(mutated), str(tmp_path / "lane"), str(tmp_path / "lane_harness"))
    p = subprocess.run([exe, "1", "40000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 1 and "MISMATCH" in p.stderr
