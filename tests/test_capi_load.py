"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares (no compute calls -- those need a GPU), and the host-side batch logic."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sickle_b200.h")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "sickle_b200", "libsickle_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", ROOT, "lib"])
    return ctypes.CDLL(so)


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 12, names
    for n in names:
        assert hasattr(lib, n), "libsickle_b200.so does not export %s" % n


def test_python_binding_lists_every_export():
    from sickle_b200 import capi

    assert sorted(capi.EXPORTS) == declared_functions()


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and as C++, and a C program using every entry
    point links against the library (nothing runs: no GPU here)."""
    for std, lang in (("c99", "c"), ("c++17", "c++")):
        subprocess.run(["gcc", "-std=" + std, "-Wall", "-Werror", "-fsyntax-only", "-x", lang, HEADER], check=True)
    src = tmp_path / "use.c"
    src.write_text(
        '#include "sickle_b200.h"\n#include <stdio.h>\n'
        "int main(int argc, char **argv) {\n"
        "  if (argc < 99) { printf(\"%d\\n\", sk_abi_version()); return 0; }\n"
        "  sk_params p = {0}; sk_result r; void *o[3] = {0}; uint64_t c[3] = {0};\n"
        "  sk_ctx *x = sk_create(0, 1u << 20, 2, &p);\n"
        "  char *b = sk_in_buffer(x, 0, 0); (void)b; (void)sk_slot_bytes(x); (void)sk_device_count();\n"
        "  sk_upload(x, 0, 0, 0, 16); sk_submit(x, 0, 0, 16, 0, 0); sk_wait(x, 0, &r);\n"
        "  sk_trim_device(x, 0, 0, 0, 0, 0, o, c, 0); sk_result_device(x, 0, 0, &r);\n"
        "  puts(sk_last_error()); sk_destroy(x); return 0; }\n")
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", os.path.join(ROOT, "sickle_b200"), "-lsickle_b200", "-Wl,-rpath," + os.path.join(ROOT, "sickle_b200")],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, check=True).stdout
    assert int(out) >= 1


def test_abi_version(lib):
    lib.sk_abi_version.restype = ctypes.c_int
    assert lib.sk_abi_version() == 1


def test_create_fails_loudly_without_gpu(lib):
    """No CPU fallback: without a CUDA device sk_create returns NULL and says why."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sickle_b200 import capi

    with pytest.raises(capi.SickleError) as ei:
        capi.Context(capi.make_params("sanger"), 1 << 20, 1)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_struct_layout_matches_header():
    """sizeof checks via a tiny C program compiled against the header."""
    import tempfile

    from sickle_b200 import capi

    prog = r'''
    #include <stdio.h>
    #include "sickle_b200.h"
    int main(void){ printf("%zu %zu %zu\n", sizeof(sk_params), sizeof(sk_error_info), sizeof(sk_result)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, r = map(int, subprocess.check_output([exe]).split())
    assert (a, b, r) == (ctypes.sizeof(capi.Params), ctypes.sizeof(capi.ErrorInfo), ctypes.sizeof(capi.Result))


def test_reference_batches_match_oracle_batch_count():
    """Host batch-geometry logic (GZReader emulation) agrees with the oracle's batch count."""
    import oracle_py as orc
    from sickle_b200 import runner

    gdir = os.path.join(ROOT, "tests", "golden")
    for name, minlines, mode in (("se_r150.fastq", 4, orc.MODE_SE), ("il15_inter.fastq", 8, orc.MODE_PE_INTER),
                                 ("varlen_illumina.fastq", 4, orc.MODE_SE), ("tiny_reads.fastq", 4, orc.MODE_SE)):
        data = open(os.path.join(gdir, name), "rb").read()
        bl = runner.recommended_batch_len(len(data), 512, mode != orc.MODE_SE)
        assert bl == orc.lib().so_recommended_batch_len(len(data), 512, int(mode != orc.MODE_SE))
        rngs = runner.reference_batches(data, bl, minlines)
        qt = "illumina" if "il" in name else "sanger"
        r = orc.run(mode, orc.make_params(qt, l=0, q=0), data)
        assert r["rc"] == 0
        assert len(rngs) == r["counters"]["n_batches"], name
        assert rngs[0][0] == 0 and all(rngs[i][1] == rngs[i + 1][0] for i in range(len(rngs) - 1))
        for a, e in rngs:
            assert data[a:e].count(b"\n") % minlines == 0
