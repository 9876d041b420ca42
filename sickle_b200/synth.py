"""Synthetic FASTQ generators for tests and bench.py.

Not part of the trimming path: this only fabricates inputs.  The record shape the
bench is quoted on is **R150**: a 20-char name line ``@SRR000001.%09d``, 150 bases, a bare
``+`` and 150 Sanger qualities = 325 bytes including the four newlines.

Quality model (per read of length L): plateau Q37, decay starting at ``onset ~ U[L/3, 5L/4]``
with ``slope ~ U[0.15, 0.6]`` Q/base, Gaussian noise sigma 3, clipped to [2, 41], the first three
bases min'ed with U[2, 41]; ``bad_frac`` of the reads instead use ``onset ~ U[0, L/4]`` and
``slope ~ U[1, 3]`` so that discards and PE singles occur.  P(N) = 0.002 per base.
"""
from __future__ import annotations

import numpy as np

R150_NAME_FMT = "@SRR000001.%09d"
R150_RECORD_BYTES = 325
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)

QUAL_OFFSET = {"sanger": 33, "illumina": 64, "solexa": 64}
QUAL_RANGE = {"sanger": (2, 41), "illumina": (0, 40), "solexa": (-5, 40)}


def _quals(rng: np.random.Generator, n: int, L: int, lo: int, hi: int, bad_frac: float) -> np.ndarray:
    pos = np.arange(L, dtype=np.float32)[None, :]
    bad = rng.random(n) < bad_frac
    onset = np.where(bad, rng.uniform(0, L / 4, n), rng.uniform(L / 3, 5 * L / 4, n)).astype(np.float32)
    slope = np.where(bad, rng.uniform(1.0, 3.0, n), rng.uniform(0.15, 0.6, n)).astype(np.float32)
    q = 37.0 - np.maximum(pos - onset[:, None], 0.0) * slope[:, None]
    q += rng.normal(0.0, 3.0, (n, L)).astype(np.float32)
    q = np.clip(np.rint(q), lo, hi).astype(np.int16)
    k = min(3, L)
    q[:, :k] = np.minimum(q[:, :k], rng.integers(lo, hi + 1, (n, k), dtype=np.int16))
    return q


def fixed_length_records(n: int, L: int = 150, qualtype: str = "sanger", seed: int = 2, start: int = 0,
                         p_N: float = 0.002, bad_frac: float = 0.10, name_fmt: str = R150_NAME_FMT,
                         suffix: str = "") -> np.ndarray:
    """Return an ``[n, record_bytes]`` uint8 matrix of fixed-width FASTQ records (bare '+')."""
    rng = np.random.default_rng([seed, start])
    off = QUAL_OFFSET[qualtype]
    lo, hi = QUAL_RANGE[qualtype]
    names = np.array([(name_fmt % (start + i) + suffix).encode() for i in range(n)], dtype="S")
    name_w = names.dtype.itemsize
    rec = np.empty((n, name_w + 1 + L + 1 + 1 + 1 + L + 1), dtype=np.uint8)
    rec[:, :name_w] = names.view(np.uint8).reshape(n, name_w)
    c = name_w
    rec[:, c] = 10
    seq = _BASES[rng.integers(0, 4, (n, L))]
    seq[rng.random((n, L)) < p_N] = ord("N")
    rec[:, c + 1:c + 1 + L] = seq
    c += 1 + L
    rec[:, c] = 10
    rec[:, c + 1] = ord("+")
    rec[:, c + 2] = 10
    rec[:, c + 3:c + 3 + L] = (_quals(rng, n, L, lo, hi, bad_frac) + off).astype(np.uint8)
    rec[:, c + 3 + L] = 10
    return rec


def variable_length_records(n: int, len_lo: int, len_hi: int, qualtype: str, seed: int,
                            p_N: float = 3e-4, p_n: float = 2e-4, plus_name_every: int = 5,
                            log_uniform: bool = True, bad_frac: float = 0.10,
                            name_prefix: str = "@read") -> bytes:
    """Records with L ~ (log-)uniform[len_lo, len_hi]; every ``plus_name_every``-th has ``+name``."""
    rng = np.random.default_rng([seed, len_lo, len_hi])
    off = QUAL_OFFSET[qualtype]
    lo, hi = QUAL_RANGE[qualtype]
    out = []
    for i in range(n):
        if log_uniform and len_lo > 0:
            L = int(round(np.exp(rng.uniform(np.log(len_lo), np.log(len_hi)))))
        else:
            L = int(rng.integers(len_lo, len_hi + 1))
        L = max(1, L)
        name = ("%s.%d len=%d" % (name_prefix, i, L)).encode()
        seq = _BASES[rng.integers(0, 4, L)].copy()
        r = rng.random(L)
        seq[r < p_N] = ord("N")
        seq[(r >= p_N) & (r < p_N + p_n)] = ord("n")
        q = (_quals(rng, 1, L, lo, hi, bad_frac)[0] + off).astype(np.uint8)
        plus = b"+" + name[1:] if (plus_name_every and i % plus_name_every == 0) else b"+"
        out.append(name + b"\n" + seq.tobytes() + b"\n" + plus + b"\n" + q.tobytes() + b"\n")
    return b"".join(out)


def paired_records(n_pairs: int, L: int = 150, qualtype: str = "sanger", seed: int = 3, start: int = 0,
                   **kw):
    """Return (forward, reverse, interleaved) uint8 matrices; names differ only in /1 /2."""
    f = fixed_length_records(n_pairs, L, qualtype, seed, start, suffix="/1", **kw)
    r = fixed_length_records(n_pairs, L, qualtype, seed + 7919, start, suffix="/2", **kw)
    inter = np.empty((2 * n_pairs, f.shape[1]), dtype=np.uint8)
    inter[0::2] = f
    inter[1::2] = r
    return f, r, inter


def r150_records_torch(n: int, start: int, device, seed: int = 2, p_N: float = 0.002,
                       bad_frac: float = 0.10, suffix: bytes = b""):
    """Same R150 model generated with torch on `device` (for bench.py's large inputs).

    Returns a uint8 tensor ``[n, 325 + len(suffix)]``.  The streams differ from the numpy
    generator's (different RNG) -- only the distribution is the same.
    """
    import torch

    L = 150
    g = torch.Generator(device=device)
    g.manual_seed((seed << 32) + start)
    name_w = 20 + len(suffix)
    W = name_w + 1 + L + 1 + 1 + 1 + L + 1
    rec = torch.empty((n, W), dtype=torch.uint8, device=device)
    prefix = torch.tensor(list(b"@SRR000001."), dtype=torch.uint8, device=device)
    rec[:, :11] = prefix
    idx = torch.arange(start, start + n, device=device, dtype=torch.int64)
    for d in range(9):
        rec[:, 11 + d] = ((idx // (10 ** (8 - d))) % 10 + 48).to(torch.uint8)
    if suffix:
        rec[:, 20:name_w] = torch.tensor(list(suffix), dtype=torch.uint8, device=device)
    c = name_w
    rec[:, c] = 10
    bases = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    seq = bases[torch.randint(0, 4, (n, L), device=device, generator=g)]
    seq[torch.rand((n, L), device=device, generator=g) < p_N] = ord("N")
    rec[:, c + 1:c + 1 + L] = seq
    del seq
    c += 1 + L
    rec[:, c] = 10
    rec[:, c + 1] = ord("+")
    rec[:, c + 2] = 10
    pos = torch.arange(L, device=device, dtype=torch.float32)[None, :]
    bad = torch.rand(n, device=device, generator=g) < bad_frac
    u1 = torch.rand(n, device=device, generator=g)
    u2 = torch.rand(n, device=device, generator=g)
    onset = torch.where(bad, u1 * (L / 4), L / 3 + u1 * (5 * L / 4 - L / 3))
    slope = torch.where(bad, 1.0 + 2.0 * u2, 0.15 + 0.45 * u2)
    q = 37.0 - torch.clamp(pos - onset[:, None], min=0.0) * slope[:, None]
    q += 3.0 * torch.randn((n, L), device=device, generator=g)
    q = torch.clamp(torch.round(q), 2, 41)
    head = torch.randint(2, 42, (n, 3), device=device, generator=g).to(torch.float32)
    q[:, :3] = torch.minimum(q[:, :3], head)
    rec[:, c + 3:c + 3 + L] = (q + 33).to(torch.uint8)
    rec[:, c + 3 + L] = 10
    return rec
