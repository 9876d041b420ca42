"""Small workload for `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python profiles/sanitize_run.py`:
every kernel of the library (kf_fused<5|7|9|11> + kf_finalize, k1_line_index, k2_trim_route incl. the warp-wide
long-read path, k3_emit, k_finalize) on small inputs, each result compared with the CPU oracle, plus a few damaged
inputs (fused kernel -> hand-over -> general path -> data error).  Prints one line per case and "SANITIZE_RUN_OK"."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as orc  # noqa: E402
from sickle_b200 import capi, runner, synth  # noqa: E402


def one(tag, cmode, omode, data, data2=b"", qualtype="sanger", q=20, l=20, x=False, n=False, slot=1 << 18, env=None, threads=1):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        # (threads > 1: the reference's own batch geometry; else one batch -- the reference mis-parses files whose
        #  size / 8 is below the longest line, SURVEY.md 9-D11, which this implementation does not imitate)
        want = orc.run(omode, orc.make_params(qualtype, q, l, x, n), data, data2, threads=threads) if threads > 1 else \
            orc.run(omode, orc.make_params(qualtype, q, l, x, n), data, data2, batch_len=1 << 40)
        p = capi.make_params(qualtype, q, l, x, n, mode=cmode, emulate_threads=threads)
        got = err = None
        with capi.Context(p, slot, 1) as ctx:
            try:
                got = runner.trim_stream_reference_order(ctx, data, data2) if threads > 1 else runner.trim_stream(ctx, data, data2)
            except runner.DataError as e:
                err = e
        if want["rc"] == 0:
            assert err is None, (tag, err)
            for s in range(3):
                assert got["out"][s] == want["out"][s], (tag, s)
            print("%-44s ok   out %d+%d+%d bytes, fused batches %d/%d" % (tag, *[len(o) for o in got["out"]], got.get("fused_batches", 0), got.get("batches", 0)))
        else:
            assert err is not None and err.kind == want["rc"] and err.record == want["err"]["record"], (tag, err, want)
            print("%-44s ok   data error kind %d at record %d" % (tag, err.kind, err.record))
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)


def main():
    se = synth.fixed_length_records(3000, 150, "sanger", seed=21).tobytes()
    f, r, inter = synth.paired_records(1200, 150, "sanger", seed=22)
    f, r, inter = f.tobytes(), r.tobytes(), inter.tobytes()
    for ch in ("5", "7", "9", "11"):
        one("se R150 fused CH=%s" % ch, capi.MODE_SE, orc.MODE_SE, se, env={"SICKLE_B200_FUSED_CH": ch})
    one("se R150 -x -n q30 fused", capi.MODE_SE, orc.MODE_SE, se, q=30, l=5, x=True, n=True)
    one("se R150 general path", capi.MODE_SE, orc.MODE_SE, se, env={"SICKLE_B200_PATH": "general"})
    one("se R150 -a 3 (reference order)", capi.MODE_SE, orc.MODE_SE, se, threads=3, slot=1 << 21)
    one("pe interleaved fused", capi.MODE_PE_INTER, orc.MODE_PE_INTER, inter)
    one("pe interleaved -M fused", capi.MODE_PE_INTER_M, orc.MODE_PE_INTER_M, inter)
    one("pe two files general", capi.MODE_PE_2FILE, orc.MODE_PE_2FILE, f, r)
    short = synth.variable_length_records(3000, 20, 90, "illumina", 23)
    one("se 20-90 bases illumina", capi.MODE_SE, orc.MODE_SE, short, qualtype="illumina", n=True, slot=1 << 17)
    for qt, seed in (("illumina", 24), ("solexa", 25)):
        long_ = synth.variable_length_records(40, 1000, 20000, qt, seed)
        one("se 1-20 kb %s -x -n (warp-wide path)" % qt, capi.MODE_SE, orc.MODE_SE, long_, qualtype=qt, x=True, n=True, slot=1 << 20)
    # damaged inputs: a bad quality byte late in the file, a short quality line, a missing line
    lines = se.split(b"\n")
    lines[4 * 2500 + 3] = b"\x7f" + lines[4 * 2500 + 3][1:]
    one("se bad quality byte (hand-over + error)", capi.MODE_SE, orc.MODE_SE, b"\n".join(lines))
    lines = se.split(b"\n")
    lines[4 * 1000 + 3] = lines[4 * 1000 + 3][:70]
    one("se short quality line", capi.MODE_SE, orc.MODE_SE, b"\n".join(lines))
    lines = se.split(b"\n")
    del lines[4 * 700 + 2]
    one("se missing '+' line", capi.MODE_SE, orc.MODE_SE, b"\n".join(lines))
    big = np.frombuffer(synth.variable_length_records(3, 1500, 6000, "sanger", 26), dtype=np.uint8).tobytes()
    bad = bytearray(big)
    q0 = big.index(b"\n+", 0)
    bad[big.index(b"\n", q0 + 2) + 1 + 700] = 0x1f
    one("se long read, bad byte (warp-wide path)", capi.MODE_SE, orc.MODE_SE, bytes(bad))
    print("SANITIZE_RUN_OK")


if __name__ == "__main__":
    main()
