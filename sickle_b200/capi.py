"""ctypes binding of the C ABI in include/sickle_b200.h (libsickle_b200.so).

Thin by design: the product's host side is the C++ `sickle` CLI (host/); Python is used by the
tests and by bench.py.  There is no fallback: if the shared library is missing or CUDA is not
usable, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SICKLE_B200_LIB", os.path.join(_HERE, "libsickle_b200.so"))   # override: A/B builds

QUAL_SANGER, QUAL_SOLEXA, QUAL_ILLUMINA = 1, 2, 3
QUALTYPE = {"sanger": QUAL_SANGER, "solexa": QUAL_SOLEXA, "illumina": QUAL_ILLUMINA}
MODE_SE, MODE_PE_2FILE, MODE_PE_INTER, MODE_PE_INTER_M = 0, 1, 2, 3
OUT_MAIN, OUT_MATE2, OUT_SINGLES = 0, 1, 2
SK_OK, SK_E_ARG, SK_E_CUDA, SK_E_NOMEM, SK_E_CAPACITY = 0, -1, -2, -3, -4

EXPORTS = ["sk_abi_version", "sk_device_count", "sk_last_error", "sk_create", "sk_destroy", "sk_in_buffer",
           "sk_slot_bytes", "sk_upload", "sk_submit", "sk_wait", "sk_trim_device", "sk_result_device"]


class Params(C.Structure):
    _fields_ = [("qualtype", C.c_int32), ("qual_threshold", C.c_int32), ("length_threshold", C.c_int32),
                ("no_fiveprime", C.c_int32), ("trunc_n", C.c_int32), ("mode", C.c_int32),
                ("emulate_threads", C.c_int32), ("has_singles", C.c_int32)]


class ErrorInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("file", C.c_int32), ("record", C.c_int64), ("position", C.c_int32),
                ("byte", C.c_int32), ("line_off", C.c_uint64 * 4), ("line_len", C.c_uint64 * 4)]


class Result(C.Structure):
    _fields_ = [("out", C.c_void_p * 3), ("out_bytes", C.c_uint64 * 3), ("consumed", C.c_uint64 * 2),
                ("records", C.c_uint64 * 2), ("kept", C.c_int64), ("discard", C.c_int64), ("kept_p", C.c_int64),
                ("discard_p", C.c_int64), ("kept_s1", C.c_int64), ("kept_s2", C.c_int64),
                ("discard_s1", C.c_int64), ("discard_s2", C.c_int64), ("error", ErrorInfo),
                ("kernel_ms", C.c_float), ("stage_ms", C.c_float * 4), ("kernel_launches", C.c_uint32),
                ("fused", C.c_uint32)]

    COUNTERS = ("kept", "discard", "kept_p", "discard_p", "kept_s1", "kept_s2", "discard_s1", "discard_s2")

    def counters(self):
        return {k: getattr(self, k) for k in self.COUNTERS}


class SickleError(RuntimeError):
    pass


_lib = None


def load():
    """Load libsickle_b200.so (built in-tree by `make lib` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SickleError("%s not found: build it with `make lib` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sk_abi_version.restype = C.c_int
    lib.sk_device_count.restype = C.c_int
    lib.sk_last_error.restype = C.c_char_p
    lib.sk_create.restype = C.c_void_p
    lib.sk_create.argtypes = [C.c_int, C.c_uint64, C.c_int, C.POINTER(Params)]
    lib.sk_destroy.argtypes = [C.c_void_p]
    lib.sk_destroy.restype = None
    lib.sk_in_buffer.restype = C.c_void_p
    lib.sk_in_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.sk_slot_bytes.restype = C.c_uint64
    lib.sk_slot_bytes.argtypes = [C.c_void_p]
    lib.sk_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
    lib.sk_submit.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.sk_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(Result)]
    lib.sk_trim_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                   C.POINTER(C.c_void_p * 3), C.POINTER(C.c_uint64 * 3), C.c_void_p]
    lib.sk_result_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Result)]
    _lib = lib
    return lib


def make_params(qualtype="sanger", q=20, l=20, x=False, n=False, mode=MODE_SE, emulate_threads=1, has_singles=True):
    qt = QUALTYPE[qualtype] if isinstance(qualtype, str) else int(qualtype)
    return Params(qt, q, l, int(x), int(n), mode, emulate_threads, int(has_singles))


class Context:
    """One GPU context (sk_ctx).  Not thread-safe; one per GPU."""

    def __init__(self, params: Params, slot_bytes: int, n_slots: int = 2, device: int = 0):
        self.lib = load()
        self.params = params
        self.n_slots = n_slots
        self.handle = self.lib.sk_create(device, slot_bytes, n_slots, C.byref(params))
        if not self.handle:
            raise SickleError("sk_create failed: %s" % self.lib.sk_last_error().decode())
        self.slot_bytes = self.lib.sk_slot_bytes(self.handle)

    def close(self):
        if self.handle:
            self.lib.sk_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != SK_OK:
            raise SickleError("%s failed (%d): %s" % (what, rc, self.lib.sk_last_error().decode()))

    def in_buffer(self, slot: int, which: int = 0):
        """The slot's pinned input buffer as a writable ctypes char array."""
        p = self.lib.sk_in_buffer(self.handle, slot, which)
        if not p:
            raise SickleError("sk_in_buffer: %s" % self.lib.sk_last_error().decode())
        return (C.c_char * self.slot_bytes).from_address(p)

    def in_buffer_address(self, slot: int, which: int = 0) -> int:
        p = self.lib.sk_in_buffer(self.handle, slot, which)
        if not p:
            raise SickleError("sk_in_buffer: %s" % self.lib.sk_last_error().decode())
        return p

    def upload(self, slot, which, offset, nbytes):
        self._check(self.lib.sk_upload(self.handle, slot, which, offset, nbytes), "sk_upload")

    def submit(self, slot, start0, end0, start1=0, end1=0):
        self._check(self.lib.sk_submit(self.handle, slot, start0, end0, start1, end1), "sk_submit")

    def wait(self, slot) -> Result:
        r = Result()
        self._check(self.lib.sk_wait(self.handle, slot, C.byref(r)), "sk_wait")
        return r

    @staticmethod
    def out_bytes(res: Result, stream: int) -> bytes:
        n = res.out_bytes[stream]
        if not n or not res.out[stream]:
            return b""
        return C.string_at(res.out[stream], n)

    def trim_device(self, in0_ptr, n0, in1_ptr, n1, out_ptrs, out_caps, stream_ptr=None, slot=0):
        outs = (C.c_void_p * 3)(*[C.c_void_p(p) if p else None for p in out_ptrs])
        caps = (C.c_uint64 * 3)(*out_caps)
        self._check(self.lib.sk_trim_device(self.handle, slot, C.c_void_p(in0_ptr), n0,
                                            C.c_void_p(in1_ptr) if in1_ptr else None, n1, C.byref(outs),
                                            C.byref(caps), C.c_void_p(stream_ptr) if stream_ptr else None),
                    "sk_trim_device")

    def result_device(self, stream_ptr=None, slot=0) -> Result:
        r = Result()
        self._check(self.lib.sk_result_device(self.handle, slot, C.c_void_p(stream_ptr) if stream_ptr else None,
                                              C.byref(r)), "sk_result_device")
        return r
