"""ctypes binding of the CPU oracle (oracle/sickle_oracle.c) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

QUALTYPE = {"sanger": 1, "solexa": 2, "illumina": 3}
MODE_SE, MODE_PE_2FILE, MODE_PE_INTER, MODE_PE_INTER_M = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [("qualtype", C.c_int), ("qual_threshold", C.c_int), ("length_threshold", C.c_int),
                ("no_fiveprime", C.c_int), ("trunc_n", C.c_int)]


class Cut(C.Structure):
    _fields_ = [("five", C.c_int), ("three", C.c_int)]


class Error(C.Structure):
    _fields_ = [("kind", C.c_int), ("record", C.c_int64), ("file", C.c_int), ("position", C.c_int),
                ("byte", C.c_int)]


class Counters(C.Structure):
    _fields_ = [("kept", C.c_int64), ("discard", C.c_int64), ("kept_p", C.c_int64), ("discard_p", C.c_int64),
                ("kept_s1", C.c_int64), ("kept_s2", C.c_int64), ("discard_s1", C.c_int64),
                ("discard_s2", C.c_int64), ("records_in", C.c_int64 * 2), ("n_batches", C.c_int64)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_build", "libsickle_oracle.so")
        src = os.path.join(ROOT, "oracle", "sickle_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        _LIB = C.CDLL(so)
        _LIB.so_run.restype = C.c_int
        _LIB.so_sliding_window.restype = C.c_int
        _LIB.so_recommended_batch_len.restype = C.c_int64
        _LIB.so_recommended_batch_len.argtypes = [C.c_int64, C.c_int64, C.c_int]
    return _LIB


def make_params(qualtype="sanger", q=20, l=20, x=False, n=False):
    return Params(QUALTYPE[qualtype], q, l, int(x), int(n))


def sliding_window(seq: bytes, qual: bytes, p: Params):
    """Returns (rc, five, three, err_position, err_byte, visited)."""
    cut, err, vis = Cut(), Error(), C.c_int(0)
    rc = lib().so_sliding_window(seq, C.c_size_t(len(seq)), qual, C.c_size_t(len(qual)), C.byref(p),
                                 C.byref(cut), C.byref(err), C.byref(vis))
    return rc, cut.five, cut.three, err.position, err.byte, vis.value


def run(mode: int, p: Params, in1: bytes, in2: bytes = b"", threads: int = 1, batch_len: int | None = None,
        has_singles: bool = True, b_mib: int = 512):
    """Returns dict(rc, out=[bytes]*3, counters=dict, err=dict)."""
    L = lib()
    if batch_len is None:
        batch_len = L.so_recommended_batch_len(len(in1), b_mib, int(mode != MODE_SE))
    cap = len(in1) + len(in2) + 64
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    outp = (C.c_char_p * 3)(*[C.cast(b, C.c_char_p) for b in bufs])
    caps = (C.c_size_t * 3)(cap, cap, cap)
    lens = (C.c_size_t * 3)()
    ctr, err = Counters(), Error()
    rc = L.so_run(C.c_int(mode), C.byref(p), C.c_int(threads), C.c_int64(batch_len), C.c_int(int(has_singles)),
                  in1, C.c_size_t(len(in1)), in2, C.c_size_t(len(in2)), outp, caps, lens, C.byref(ctr), C.byref(err))
    out = [bufs[i].raw[:lens[i]] for i in range(3)]
    counters = {k: getattr(ctr, k) for k, _ in Counters._fields_ if k != "records_in"}
    counters["records_in"] = list(ctr.records_in)
    return dict(rc=rc, out=out, counters=counters,
                err=dict(kind=err.kind, record=err.record, file=err.file, position=err.position, byte=err.byte))
