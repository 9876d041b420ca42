"""The byte movers of the single-pass kernel (sickle_b200/csrc/sk_copy.cuh: the staging copy inside
shared memory and the phase-shifting flush to global memory) compiled for the HOST and checked against
memcpy -- CPU only, no CUDA.  Every source / destination byte phase, lengths from 0 up, and the bytes
around the destination must stay untouched (they belong to other lanes / other tiles).  Same shim as
tests/test_lane_logic.py; test infrastructure only.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(src_header, build_dir, out):
    os.makedirs(build_dir, exist_ok=True)
    shutil.copy(src_header, os.path.join(build_dir, "sk_copy.cuh"))
    for name in ("sk_device.cuh", "k1_index.cuh"):
        shutil.copy(os.path.join(ROOT, "tests", "host_stub", "lane_shim", name), build_dir)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-I" + build_dir,
                           os.path.join(ROOT, "tests", "host_stub", "copy_harness.cpp"), "-o", out])
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_copies_equal_memcpy(tmp_path_factory, seed):
    d = tmp_path_factory.getbasetemp() / "copy_harness"
    exe = str(d / "h")
    if not os.path.exists(exe):
        build(os.path.join(ROOT, "sickle_b200", "csrc", "sk_copy.cuh"), str(d), exe)
    p = subprocess.run([exe, str(seed), "200000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.strip() == "smem_copy 200000 flush 10001 mismatches 0", (p.stdout, p.stderr[-1000:])


@pytest.mark.parametrize("old,new", [("const uint32_t sh = (src & 3u) * 8u;", "const uint32_t sh = (src & 3u) * 8u + (len == 77u ? 8u : 0u);"),
                                     ("const uint32_t c_hi = end >> 4;", "const uint32_t c_hi = (end >> 4) - (tot > 100u && tot % 64u == 3u ? 1u : 0u);")])
def test_harness_notices_a_wrong_copy(tmp_path, old, new):
    src = open(os.path.join(ROOT, "sickle_b200", "csrc", "sk_copy.cuh")).read()
    assert src.count(old) == 1
    mutated = tmp_path / "mutated.cuh"
    mutated.write_text(src.replace(old, new))
    exe = build(str(mutated), str(tmp_path / "b"), str(tmp_path / "h"))
    p = subprocess.run([exe, "1", "150000"], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "MISMATCH" in p.stderr
