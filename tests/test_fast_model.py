"""CPU: the executable specification of the k-mer count pyramid fast path (tools/fast_model.c, both the entry form and
the bit-mask form the CUDA kernel runs) against the oracle, bit for bit, on seeded sets; plus the table definition."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle.binding import Oracle, SeedOpt

fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
sy = importlib.import_module("bwa-mem-harp2_b200.synth")
kt = importlib.import_module("bwa-mem-harp2_b200.kmer_tables")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Tab(C.Structure):
    _fields_ = [("DL", C.c_int), ("has_top", C.c_int), ("cnt", C.c_void_p * 14), ("cum", C.c_void_p * 15), ("pyr", C.c_void_p), ("top", C.c_void_p)]


class St(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in "calls rounds fm_fwd fm_bwd mat_emit mat_grow lookups_direct lookups_pyr lookups_top escapes reads".split()]


@pytest.fixture(scope="module")
def model():
    so = os.path.join(ROOT, "tools", "libfast_model.so")
    Oracle  # noqa: B018  (liboracle.so is built by importing / constructing the oracle below)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    subprocess.run(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "fast_model.c"),
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(so)
    lib.fast_collect.restype = C.c_int64
    return lib


def run_model(lib, o, tb, seq, offs, opt, variant):
    tab = Tab()
    tab.DL, tab.has_top = tb.DL, int(tb.top is not None)
    for L in range(1, tb.DL + 1):
        tab.cnt[L] = tb.cnt[L].ctypes.data
    for L in range(1, tb.DL + 2):
        tab.cum[L] = tb.cum[L].ctypes.data
    tab.pyr = tb.pyr.ctypes.data
    tab.top = tb.top.ctypes.data if tb.top is not None else None
    C.c_int.in_dll(lib, "fast_model_variant").value = variant
    n = len(offs) - 1
    cap = 64 * n + 64
    intv = np.zeros((cap, 4), np.uint64)
    ro = np.zeros(n + 1, np.int64)
    step = np.zeros(cap, np.uint16)
    esc = np.zeros(max(n, 1), np.uint8)
    st = St()
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    tot = lib.fast_collect(C.byref(o.ix), C.byref(tab), C.c_int64(n), P(seq, C.c_uint8), P(offs, C.c_int64), C.byref(opt), P(intv, C.c_uint64),
                           C.c_int64(cap), P(ro, C.c_int64), P(step, C.c_uint16), P(esc, C.c_uint8), C.byref(st))
    assert tot <= cap
    return intv[:tot], ro, step[:tot], esc[:n], st


@pytest.mark.parametrize("ref_bp,read_len,err,DL,top,opt", [
    (200_000, 101, 0.01, 4, True, SeedOpt()),
    (200_000, 250, 0.02, 5, True, SeedOpt()),
    (60_000, 101, 0.03, 3, False, SeedOpt(start_width=2)),
    (500_000, 76, 0.01, 5, True, SeedOpt(min_seed_len=15, split_factor=1.2, split_width=4)),
])
def test_model_matches_oracle(model, ref_bp, read_len, err, DL, top, opt):
    ref = sy.make_reference(ref_bp, 21)
    ix = fm.build_index(ref)
    tb = kt.build_kmer_tables(fm.text_from_forward(ref).numpy(), DL, top)
    seq, offs = sy.to_batch(sy.simulate_reads(ref, 1500, read_len, err, seed=9, n_frac=0.08, paired=True))
    o = Oracle(ix)
    want = o.collect(seq, offs, opt, nthreads=4)
    for variant in (0, 1):
        intv, ro, step, esc, st = run_model(model, o, tb, seq, offs, opt, variant)
        assert esc.sum() < len(esc)          # the tables answered for some reads ...
        n_ok = 0
        for r in range(len(esc)):
            if esc[r]:
                assert ro[r + 1] == ro[r]
                continue
            a, b = intv[ro[r]:ro[r + 1]], want["intv"][want["read_off"][r]:want["read_off"][r + 1]]
            assert a.shape == b.shape and np.array_equal(a, b), f"read {r} variant {variant}"
            assert np.array_equal(step[ro[r]:ro[r + 1]], want["step"][want["read_off"][r]:want["read_off"][r + 1]])
            n_ok += 1
        assert n_ok > len(esc) // 5           # ... and for a good share of them even at these toy table depths
        assert st.fm_fwd + st.fm_bwd > 0 and st.mat_grow > 0


def test_tables_against_brute_force():
    """cnt / cum / pyr / top of a tiny text against direct counting and sorting of its suffixes."""
    rng = np.random.default_rng(5)
    fwd = rng.integers(0, 4, 300).astype(np.uint8)
    T = np.concatenate([fwd, 3 - fwd[::-1]])
    n = T.size
    DL = 2
    tb = kt.build_kmer_tables(T, DL, True)
    sufs = sorted(range(n), key=lambda p: bytes(T[p:]))     # '$' (end of text) sorts first: a prefix sorts before its extensions
    for L in range(1, DL + 6):
        for code in rng.integers(0, 4 ** L, 60):
            pat = bytes((code >> (2 * (L - 1 - k))) & 3 for k in range(L))
            occ = sum(bytes(T[p:p + L]) == pat for p in range(n - L + 1))
            x0 = 1 + sum(bytes(T[p:]) < pat for p in sufs)
            if L <= DL:
                assert tb.cnt[L][code] == occ
            if L <= DL + 1:
                assert tb.cum[L][code] == x0
            if L == DL + 4 and tb.pyr[code] != 255:
                assert tb.pyr[code] == occ
            if L == DL + 5 and tb.top[code] != 255:
                assert tb.top[code] == occ
    assert (tb.pyr == 255).sum() == 64 + 16 + 4 and (tb.top == 255).sum() == 4
