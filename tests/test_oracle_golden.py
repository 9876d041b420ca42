"""The CPU oracle (oracle/sickle_oracle.c) against the reference's own outputs.

(1) tests/golden/golden.json -- md5s produced by the unmodified reference binary on the committed
    synthetic fixtures (tests/golden/make_golden.py), incl. -a N order and the error fixtures.
(2) the md5 table of SURVEY.md section 4 on the reference's bundled test/ fixtures -- only when
    /root/reference is present (build container); skipped elsewhere.
"""
import os

import pytest

import golden_util as gu
import oracle_py as orc

MODES = {"se": orc.MODE_SE, "pei": orc.MODE_PE_INTER, "pe2": orc.MODE_PE_2FILE}
ERRKIND = {"Sequence ID is to short.": 1, "Invalid char at the beggining of ID.": 2,
           "Sequence line is empty": 3, "Quality line is empty.": 4,
           "Sequence and quality lines have different lengths:": 5, "ERROR: Quality value": 6}


def _run_case(case, gdir):
    kind, in1, in2 = gu.load_inputs(case, gdir)
    f = gu.parse_flags(case["flags"])
    p = orc.make_params(f["qualtype"], f["q"], f["l"], f["x"], f["n"])
    has_singles = "-s" in case["outputs"]
    return orc.run(MODES[kind], p, in1, in2, threads=case["threads"], has_singles=has_singles)


def test_all_golden_cases(golden):
    bad = []
    for case in golden["cases"]:
        r = _run_case(case, golden["dir"])
        exp = gu.expected_streams(case)
        if case["rc"] == 0:
            if r["rc"] != 0:
                bad.append((case["id"], "oracle rc", r["rc"]))
                continue
            for s in range(3):
                if exp[s] is not None and (gu.md5(r["out"][s]), len(r["out"][s])) != exp[s]:
                    bad.append((case["id"], "stream", s, len(r["out"][s]), exp[s][1]))
            c, g = r["counters"], case["counts"]
            if case["mode"] == "se":
                if (c["kept"], c["discard"]) != (g["kept"], g["discard"]):
                    bad.append((case["id"], "counts", c, g))
            else:
                if (c["kept_p"], c["discard_p"], c["kept_s1"] + c["kept_s2"],
                        c["discard_s1"] + c["discard_s2"]) != (g["kept_p"], g["discard_p"], g["kept_s"], g["discard_s"]):
                    bad.append((case["id"], "counts", c, g))
                if "kept_s1" in g and (c["kept_s1"], c["kept_s2"]) != (g["kept_s1"], g["kept_s2"]):
                    bad.append((case["id"], "s1/s2", c, g))
        else:
            want = [k for msg, k in ERRKIND.items() if msg in case["stderr"]]
            assert len(want) == 1, case["id"]
            if r["rc"] != want[0]:
                bad.append((case["id"], "err kind", r["rc"], want[0]))
            if want[0] == 6:
                import re
                pos = int(re.search(r"Quality position: (\d+)", case["stderr"]).group(1))
                val = int(re.search(r"Quality value \((-?\d+)\)", case["stderr"]).group(1))
                rec = re.search(r"FastQ record: (\S+)", case["stderr"]).group(1)
                if (r["err"]["position"] + 1, r["err"]["byte"]) != (pos, val):
                    bad.append((case["id"], "err pos/byte", r["err"], pos, val))
                in1 = gu.load_inputs(case, golden["dir"])[1]
                name = in1.split(b"\n")[4 * r["err"]["record"]].decode()
                if name != rec:
                    bad.append((case["id"], "err record", name, rec))
    assert not bad, bad[:10]
    assert len(golden["cases"]) > 200


REF_TEST = "/root/reference/test"
SURVEY_MD5 = [  # SURVEY.md section 4 (reference binary, -a 1 unless stated)
    ("se", "test.fastq", "sanger", dict(q=20, l=20), 1, ["17960489e277c3f6839d834155fd3b79"]),
    ("se", "test.fastq", "sanger", dict(q=60), 1, ["7496a067308eee736d3676d3c544030d"]),
    ("se", "test.fastq", "illumina", dict(), 1, ["0ca1dd1d8a7161883d8ac301c4b03491"]),
    ("se", "test.fastq", "illumina", dict(x=True), 1, ["132bc2ff1f914110d2821799a01402be"]),
    ("se", "test.fastq", "illumina", dict(n=True), 1, ["9fb90619c365cb4fc73be7dbb1c43286"]),
    ("se", "test.fastq", "illumina", dict(x=True, n=True), 1, ["82901542b34fe4cb0100fea423ee0ca5"]),
    ("se", "test.fastq", "solexa", dict(q=30, l=50), 1, ["07f962aba570371ea180639ba63ff4bf"]),
    ("se", "test.fastq", "illumina", dict(q=35, l=0), 1, ["e5029523a071706a814cbb365b4a3462"]),
    ("pe2", "test.f.fastq,test.r.fastq", "sanger", dict(q=60), 1,
     ["12bdd85ffdaeeeb7ca08ad59365287f0", "306a7f40a19903bbb967eeceb988b5a6", "5e1d30c504ac80e042cd4f1e42b004cf"]),
    ("pe2", "test.f.fastq,test.r.fastq", "illumina", dict(), 1,
     ["f2478e9cea050cb2690808f2badf435e", "9107333e1b8e68f0cbf017f75a4b0bca", "6cc7df35ef9d21ade41e85126ec63be8"]),
    ("pe2", "test.f.fastq,test.r.fastq", "illumina", dict(x=True, n=True), 1,
     ["26da0a9bbc2e663ae73eda9f2b2ce08f", "acd9b987b59075d1139cc68e9fc9b2a8", "61f717857b6877c399b6dce287ed4cc9"]),
    ("pe2", "test.f.fastq,test.r.fastq", "solexa", dict(q=30, l=50), 1,
     ["b26efae9658cb18a3b5d0eb3f3c26b56", "b38e1ca5f24a1abd36a17a6ade62c905", "a0ae11d2c51c7de8e87ce7c4bf428e0b"]),
    ("pei", "test.fastq", "sanger", dict(q=60), 1,
     ["aea939b5d014c9f522644ebb161bf09b", None, "9475fab6060d303623b5d041b3e8e771"]),
    ("pei", "test.fastq", "illumina", dict(), 1,
     ["f672c74961557089767ab3d8d917fd77", None, "3db5d57ca6dbdc830632a357d7dc3927"]),
    ("pei", "test.fastq", "illumina", dict(n=True), 1,
     ["09feaf4fe0c0ee1bcad8077233cd170f", None, "23d2e5ba7c24e83aa2edd823bc692a0e"]),
    ("se", "test.fastq", "sanger", dict(), 2, ["21b0cfd7"]),
    ("se", "test.fastq", "sanger", dict(), 3, ["9b7c1c81"]),
    ("se", "test.fastq", "sanger", dict(), 4, ["6f1cc220"]),
    ("se", "test.fastq", "sanger", dict(), 8, ["bd7bb9b8"]),
    ("pei", "test.fastq", "sanger", dict(q=60), 4, ["ff5f101b", None, "6dbf28a3"]),
]


@pytest.mark.skipif(not os.path.isdir(REF_TEST), reason="reference fixtures only exist in the build container")
@pytest.mark.parametrize("row", SURVEY_MD5, ids=lambda r: "%s-%s-%s-a%d" % (r[0], r[2], "".join(map(str, r[3].values())), r[4]))
def test_survey_md5_table(row):
    kind, files, qt, kw, threads, want = row
    ins = [open(os.path.join(REF_TEST, f), "rb").read() for f in files.split(",")]
    p = orc.make_params(qt, **kw)
    r = orc.run(MODES[kind], p, ins[0], ins[1] if len(ins) > 1 else b"", threads=threads)
    assert r["rc"] == 0
    for s, w in enumerate(want):
        if w is not None:
            assert gu.md5(r["out"][s]).startswith(w), (s, len(r["out"][s]))


def test_visited_prefix_rule():
    """A bad quality byte only matters inside [0, min(L, i_break + ws))."""
    p = orc.make_params("sanger")
    q = bytearray(b"I" * 40 + b"#" * 60)
    seq = b"A" * 100
    rc, five, three, _, _, vis = orc.sliding_window(seq, bytes(q), p)
    assert (rc, five, three, vis) == (0, 0, 40, 46)
    q[45] = 32
    assert orc.sliding_window(seq, bytes(q), p)[0] == 6
    q[45] = ord("#"); q[46] = 32
    assert orc.sliding_window(seq, bytes(q), p)[:3] == (0, 0, 40)
    # shorter than -l: discarded before any byte is looked at
    assert orc.sliding_window(b"ACGT", b"\x01\x01\x01\x01", p)[:3] == (0, -1, -1)
