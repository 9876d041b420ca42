"""Helpers shared by the golden-vector tests (oracle on CPU, CUDA path on GPU)."""
from __future__ import annotations

import hashlib
import os

QUALTYPES = ("sanger", "solexa", "illumina")


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def parse_flags(flags):
    """['-t','sanger','-q','30','-x'] -> dict(qualtype=..., q=..., l=..., x=..., n=...)"""
    d = dict(qualtype=None, q=20, l=20, x=False, n=False)
    i = 0
    while i < len(flags):
        f = flags[i]
        if f == "-t":
            d["qualtype"] = flags[i + 1]; i += 2
        elif f == "-q":
            d["q"] = int(flags[i + 1]); i += 2
        elif f == "-l":
            d["l"] = int(flags[i + 1]); i += 2
        elif f == "-x":
            d["x"] = True; i += 1
        elif f == "-n":
            d["n"] = True; i += 1
        else:
            raise ValueError(f)
    return d


def load_inputs(case, gdir):
    ins = {k: open(os.path.join(gdir, v), "rb").read() for k, v in case["inputs"].items()}
    if case["mode"] == "se":
        return "se", ins["-f"], b""
    if "-c" in ins:
        return "pei", ins["-c"], b""
    return "pe2", ins["-f"], ins["-r"]


def expected_streams(case):
    """golden outputs -> [(md5, bytes) or None] * 3 in stream order (0: -o/-m, 1: -p, 2: -s)."""
    o = case["outputs"]
    s = [None, None, None]
    for k, idx in (("-o", 0), ("-m", 0), ("-p", 1), ("-s", 2)):
        if k in o:
            s[idx] = (o[k]["md5"], o[k]["bytes"])
    return s
