"""K-mer count pyramid of the indexed text -- host-side DEFINITION of the tables the fast seeding kernel reads.

The device library builds the same tables itself from the 2-bit text (``smem_gpu_build_kmer_tables``); this numpy
version is the executable definition used by the CPU model test (tools/fast_model.c) and to check the device builder.

For a pattern P of length L, in the row numbering of ``bwt_t`` (row 0 = the ``$`` suffix, bwt.h:80):
  size  x[2] = number of occurrences of P in T                     (T = forward + reverse complement, bntseq.c:268-273)
  start x[0] = 1 + number of suffixes of T that sort before P      (a suffix shorter than P that is a prefix of P sorts before it)
  start x[1] = x[0] of revcomp(P)

Tables (code of an L-mer: 2 bits per base, first base in the most significant position):
  cnt[L]  L = 1..DL     uint32   occurrences of every L-mer
  cum[L]  L = 1..DL+1   uint64   x[0] of every L-mer
  pyr     L = DL+4      uint8    occurrences, saturating; 255 = unknown (saturated, or a suffix of T of length DL+1..DL+3 has this prefix)
  top     L = DL+5      uint8    occurrences, saturating; 255 = unknown (saturated, or the last DL+4 symbols of T are its prefix)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class KmerTables:
    DL: int
    cnt: list          # cnt[L] for L in 0..DL (cnt[0] unused)
    cum: list          # cum[L] for L in 0..DL+1 (cum[0] unused)
    pyr: np.ndarray
    top: "np.ndarray | None"


def _codes(T: np.ndarray, L: int) -> np.ndarray:
    n = T.size
    c = np.zeros(n - L + 1, np.int64)
    for k in range(L):
        c = (c << 2) | T[k:n - L + 1 + k].astype(np.int64)
    return c


def _code_of(sym: np.ndarray) -> int:
    c = 0
    for v in sym:
        c = (c << 2) | int(v)
    return c


def build_kmer_tables(T: np.ndarray, DL: int, top: bool = True) -> KmerTables:
    """``T``: the whole indexed text (uint8 symbols 0..3), ``DL``: deepest direct level."""
    T = np.ascontiguousarray(T, np.uint8)
    n = T.size
    K, LP, LT = DL + 1, DL + 4, DL + 5
    assert n > LT, "text shorter than the deepest table level"
    cnt = [None] * (DL + 1)
    cum = [None] * (DL + 2)
    for L in range(1, K + 1):
        c = np.bincount(_codes(T, L), minlength=4 ** L).astype(np.int64)
        x0 = np.empty(4 ** L, np.int64)
        x0[0] = 0
        np.cumsum(c[:-1], out=x0[1:])
        x0 += 1                                            # row 0 is the '$' suffix
        for a in range(1, L):                              # the suffix of T of length a < L sorts before every L-mer whose a-prefix is >= it
            x0[_code_of(T[n - a:]) << (2 * (L - a)):] += 1
        cum[L] = x0.astype(np.uint64)
        if L <= DL:
            cnt[L] = c.astype(np.uint32)
    pyr = np.minimum(np.bincount(_codes(T, LP), minlength=4 ** LP), 255).astype(np.uint8)
    for a in range(K, LP):                                 # sums over these entries would miss the short suffix itself
        lo = _code_of(T[n - a:]) << (2 * (LP - a))
        pyr[lo:lo + 4 ** (LP - a)] = 255
    tp = None
    if top:
        tp = np.minimum(np.bincount(_codes(T, LT), minlength=4 ** LT), 255).astype(np.uint8)
        lo = _code_of(T[n - LP:]) << 2
        tp[lo:lo + 4] = 255
    return KmerTables(DL, cnt, cum, pyr, tp)
