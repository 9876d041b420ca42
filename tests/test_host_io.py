"""Host I/O stages of the CLI (host/io.{h,cpp}) -- CPU only, no CUDA.

bin/io_tool copies a file through ByteSource -> ByteSink exactly as the trimmer's batch loop does
(bounded reads, two buffers in flight, ordered asynchronous writes).  Python's gzip module is the
independent check of the `-g` output (BGZF = ordinary multi-member gzip) and the producer of plain
single-member .gz inputs.
"""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "bin", "io_tool")


@pytest.fixture(scope="module")
def tool():
    subprocess.run(["make", "-s", "bin/io_tool"], cwd=ROOT, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert os.path.exists(TOOL)
    return TOOL


def _fastq_like(nbytes, seed=5):
    rng = np.random.default_rng(seed)
    rec = []
    size = 0
    i = 0
    while size < nbytes:
        L = int(rng.integers(30, 300))
        r = b"@r%d\n" % i + bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8)) + b"\n+\n" + \
            bytes((rng.integers(2, 41, L) + 33).astype(np.uint8)) + b"\n"
        rec.append(r)
        size += len(r)
        i += 1
    return b"".join(rec)[:nbytes]


def _run(tool, src, dst, chunk, gz, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([tool, src, dst, str(chunk), "1" if gz else "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    return p.returncode, p.stdout.decode().strip(), p.stderr.decode()


@pytest.mark.parametrize("chunk", [1 << 22, 999_983, 4096])
@pytest.mark.parametrize("threads", ["1", "8"])
def test_plain_copy_is_identical(tool, tmp_path, chunk, threads):
    data = _fastq_like(9_000_001)
    src, dst = tmp_path / "a.fastq", tmp_path / "b.fastq"
    src.write_bytes(data)
    rc, out, err = _run(tool, str(src), str(dst), chunk, False, {"SICKLE_B200_IO_THREADS": threads})
    assert rc == 0, err
    assert out == "bgzf=0 gzip=0 bytes=%d" % len(data)
    assert dst.read_bytes() == data


def test_mapped_output_is_identical(tool, tmp_path):
    """SICKLE_B200_MMAP_OUT=1: large writes go through ftruncate + shared mapping + parallel memcpy,
    small ones through pwrite at the same running offset."""
    data = _fastq_like(30_000_007, seed=9)
    src, dst = tmp_path / "a.fastq", tmp_path / "b.fastq"
    src.write_bytes(data)
    for chunk in (12_000_001, 9_000_000, 100_000):   # mapped + pwrite tail, unaligned offsets, pwrite only
        rc, out, err = _run(tool, str(src), str(dst), chunk, False, {"SICKLE_B200_MMAP_OUT": "1"})
        assert rc == 0, err
        assert dst.read_bytes() == data, chunk


def _bgzf_blocks(raw):
    """Walk BGZF blocks; returns [(compressed_size, uncompressed_size)]."""
    out, o = [], 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04", o
        xlen = struct.unpack_from("<H", raw, o + 10)[0]
        extra = raw[o + 12:o + 12 + xlen]
        assert extra[:4] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", extra, 4)[0] + 1
        out.append((bsize, struct.unpack_from("<I", raw, o + bsize - 4)[0]))
        o += bsize
    assert o == len(raw)
    return out


@pytest.mark.parametrize("threads", ["1", "8"])
def test_gzip_output_is_valid_bgzf_and_reads_back_in_parallel(tool, tmp_path, threads):
    data = _fastq_like(21_000_003, seed=6)
    src, gzf, back = tmp_path / "a.fastq", tmp_path / "a.fastq.gz", tmp_path / "back.fastq"
    src.write_bytes(data)
    env = {"SICKLE_B200_IO_THREADS": threads}
    rc, out, err = _run(tool, str(src), str(gzf), 5_000_000, True, env)
    assert rc == 0, err
    raw = gzf.read_bytes()
    assert gzip.decompress(raw) == data                 # any gunzip reads it
    blocks = _bgzf_blocks(raw)
    assert blocks[-1] == (28, 0)                        # BGZF end-of-file marker
    assert sum(u for _, u in blocks) == len(data) and max(u for _, u in blocks) <= 0xff00
    assert len(raw) < 0.6 * len(data)
    # ... and back through the parallel block reader, with read sizes that split blocks
    for chunk in (1 << 22, 1_000_003, 1000):
        rc, out, err = _run(tool, str(gzf), str(back), chunk, False, env)
        assert rc == 0, err
        assert out == "bgzf=1 gzip=1 bytes=%d" % len(data)
        assert back.read_bytes() == data
    # the same file through the generic single-stream zlib reader
    rc, out, err = _run(tool, str(gzf), str(back), 1 << 20, False, dict(env, SICKLE_B200_NO_BGZF="1"))
    assert rc == 0 and out == "bgzf=0 gzip=1 bytes=%d" % len(data) and back.read_bytes() == data


def test_plain_gzip_and_concatenated_members(tool, tmp_path):
    data = _fastq_like(3_000_000, seed=7)
    gzf, back = tmp_path / "p.gz", tmp_path / "p.fastq"
    gzf.write_bytes(gzip.compress(data[:1_000_000]) + gzip.compress(data[1_000_000:]))   # cat a.gz b.gz
    rc, out, err = _run(tool, str(gzf), str(back), 700_001, False)
    assert rc == 0, err
    assert out == "bgzf=0 gzip=1 bytes=%d" % len(data)
    assert back.read_bytes() == data


def test_empty_and_damaged_inputs(tool, tmp_path):
    empty, gzf, back = tmp_path / "e.fastq", tmp_path / "e.gz", tmp_path / "e.back"
    empty.write_bytes(b"")
    rc, out, err = _run(tool, str(empty), str(gzf), 4096, True)
    assert rc == 0 and out.endswith("bytes=0")
    assert gzip.decompress(gzf.read_bytes()) == b""
    rc, out, err = _run(tool, str(gzf), str(back), 4096, False)
    assert rc == 0 and out == "bgzf=1 gzip=1 bytes=0" and back.read_bytes() == b""
    # a BGZF file cut in the middle of a block, and one with a flipped payload bit: read errors
    data = _fastq_like(400_000, seed=8)
    src, good = tmp_path / "d.fastq", tmp_path / "d.gz"
    src.write_bytes(data)
    assert _run(tool, str(src), str(good), 1 << 20, True)[0] == 0
    raw = bytearray(good.read_bytes())
    cut = tmp_path / "cut.gz"
    cut.write_bytes(bytes(raw[:len(raw) // 2]))
    assert _run(tool, str(cut), str(back), 1 << 20, False)[0] == 1
    raw[len(raw) // 3] ^= 0x10
    bad = tmp_path / "bad.gz"
    bad.write_bytes(bytes(raw))
    assert _run(tool, str(bad), str(back), 1 << 20, False)[0] == 1
    assert _run(tool, str(tmp_path / "missing"), str(back), 4096, False)[0] == 1


def test_reference_batch_cutter_matches_the_python_model(tool, tmp_path):
    """host/ref_batcher.h (block-wise newline counting) cuts exactly the batches of
    runner.reference_batches (the line-by-line model of GZReader::read_lines, itself checked against the
    reference's -a N outputs by the golden tests)."""
    import sys

    sys.path.insert(0, ROOT)
    from sickle_b200 import runner

    rng = np.random.default_rng(12)
    datas = [_fastq_like(300_000, seed=20), _fastq_like(300_001, seed=21)[:-7],        # cut mid-line, no final newline
             b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"A" * 3, b"I" * 3) for i in range(5000)),   # tiny records
             b"@only\nACGT\n+\nIIII\n", b"", b"\n\n\n\n\n", b"@x\n" + b"A" * 200_000 + b"\n+\n" + b"I" * 200_000 + b"\n"]
    big = _fastq_like(2_000_000, seed=22) * 7              # > 8 MB: the multi-threaded pre-scan takes part
    src = tmp_path / "big.fastq"
    src.write_bytes(big)
    for batch_len, minlines in ((9_000_000, 4), (3_100_000, 8), (20_000_000, 4)):
        want = [b - a for a, b in runner.reference_batches(big, batch_len, minlines)]
        p = subprocess.run([tool, "refbatch", str(src), str(batch_len), str(minlines), str(32_000_000)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0 and [int(x) for x in p.stdout.split()] == want, (batch_len, minlines)
    for di, data in enumerate(datas):
        src = tmp_path / ("d%d.fastq" % di)
        src.write_bytes(data)
        for batch_len in (20, 777, 70_000, int(rng.integers(1000, 50_000)), 10_000_000):
            for minlines in (4, 8):
                want = [b - a for a, b in runner.reference_batches(data, batch_len, minlines)]
                p = subprocess.run([tool, "refbatch", str(src), str(batch_len), str(minlines), str(2_000_000)],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                got = [int(x) for x in p.stdout.split()]
                assert p.returncode == 0 and got == want, (di, batch_len, minlines, got[:5], want[:5], len(got), len(want))


def _unit_batches_model(datas, cap, lpu):
    """Plain-Python model of host/unit_cutter.h's UnitStream loop: returns ([(bytes0, bytes1, units)], starved)."""
    st = [dict(data=d, pos=0, carry=b"", eof=False) for d in datas]
    per_unit = [lpu] + [4] * (len(datas) - 1)
    out = []
    while True:
        bufs = []
        for s in st:
            buf = s["carry"]
            want = cap - len(buf)
            if not s["eof"] and want > 0:
                got = s["data"][s["pos"]:s["pos"] + want]
                s["pos"] += len(got)
                if len(got) < want:
                    s["eof"] = True
                buf += got
            if s["eof"] and buf and buf[-1:] != b"\n":
                buf = buf[:-1] + b"\n"
            bufs.append(buf)
        units = [b.count(b"\n") // u for b, u in zip(bufs, per_unit)]
        u = min(units)
        if u == 0:
            starved = any(n == 0 and len(b) == cap and not s["eof"] for n, b, s in zip(units, bufs, st))
            return out, starved
        cuts = []
        for s, b, per in zip(st, bufs, per_unit):
            p = 0
            for _ in range(u * per):
                p = b.index(b"\n", p) + 1
            s["carry"] = b[p:]
            cuts.append(p)
        out.append((cuts[0], cuts[1] if len(cuts) > 1 else 0, u))


def test_unit_cutter_matches_the_python_model(tool, tmp_path):
    """host/unit_cutter.h: the multi-GPU driver's batches hold whole records (pairs), cover the input
    without gaps, and pair two files by record number even when their records differ in size."""
    def run(cap, lpu, *paths):
        p = subprocess.run([tool, "units", str(cap), str(lpu)] + [str(x) for x in paths], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        return p.returncode, [tuple(int(v) for v in ln.split()) for ln in p.stdout.decode().splitlines()]

    big = _fastq_like(3_000_000, seed=31) * 4                      # > 4 MiB buffers: several counting threads
    big = big[:big.rfind(b"\n@r") + 1]
    datas = [_fastq_like(300_000, seed=30), _fastq_like(300_001, seed=32)[:-7],   # cut mid-line, no final newline
             b"".join(b"@r%d\nACG\n+\nIII\n" % i for i in range(5000)), b"@only\nACGT\n+\nIIII\n", b"", b"\n\n\n\n\n", big]
    for di, data in enumerate(datas):
        src = tmp_path / ("u%d.fastq" % di)
        src.write_bytes(data)
        for cap in ((6 << 20, 2_500_000) if data is big else (1_000_000, 65_536, 4_099, 700)):
            for lpu in (4, 8):
                want, starved = _unit_batches_model([data], cap, lpu)
                rc, got = run(cap, lpu, src)
                assert rc == (3 if starved else 0) and got == want, (di, cap, lpu)
                # whole units, no gaps: the batches are a prefix of the input, line count a multiple of lpu
                if not starved and data.endswith(b"\n"):         # (an unterminated last line gets patched)
                    covered = sum(g[0] for g in got)
                    assert data[:covered].count(b"\n") == sum(g[2] for g in got) * lpu
                    assert covered == 0 or data[covered - 1:covered] == b"\n"
                    assert data[covered:].count(b"\n") < lpu        # only an incomplete unit is left over
    # two files whose records differ in size (mate 2 reads are shorter): same record count per batch
    rng = np.random.default_rng(33)
    n = 6000
    f1 = b"".join(b"@p%d/1\n%s\n+\n%s\n" % (i, b"A" * L, b"I" * L) for i, L in enumerate(rng.integers(50, 250, n)))
    f2 = b"".join(b"@p%d/2\n%s\n+\n%s\n" % (i, b"C" * L, b"5" * L) for i, L in enumerate(rng.integers(20, 90, n)))
    p1, p2 = tmp_path / "m1.fastq", tmp_path / "m2.fastq"
    p1.write_bytes(f1)
    p2.write_bytes(f2 + b"@extra\nAC\n+\nII\n")               # one record too many in file 2: dropped
    for cap in (300_000, 70_000, 1 << 20):
        want, starved = _unit_batches_model([f1, p2.read_bytes()], cap, 4)
        rc, got = run(cap, 4, p1, p2)
        assert rc == 0 and not starved and got == want
        assert sum(g[2] for g in got) == n and sum(g[0] for g in got) == len(f1) and sum(g[1] for g in got) == len(f2)
    # a record longer than the buffer
    long_rec = b"@x\n" + b"A" * 5000 + b"\n+\n" + b"I" * 5000 + b"\n"
    src = tmp_path / "long.fastq"
    src.write_bytes(long_rec * 3)
    assert run(4096, 4, src)[0] == 3
    assert run(20_000, 4, src) == (0, [(len(long_rec), 0, 1)] * 3)
