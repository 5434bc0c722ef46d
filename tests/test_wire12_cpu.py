"""The 12-byte interval record (smem_intv12_t, include/smem_gpu.h): the header's inline decoder and the Python mirror agree with the
documented bit layout, for every pos_bits the format admits, including escaped sizes.  No GPU: records are packed here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack11(intv, pos_bits):
    """bwtintv_t rows -> (records uint8[n,11], exceptions) following the comment above smem_intv11_t."""
    rec12, exc = pack12(intv, pos_bits, total_bits=22)
    b = np.empty((len(rec12), 11), np.uint8)
    for k in range(4):
        b[:, k] = (rec12[:, 0] >> np.uint32(8 * k)) & np.uint32(255)
        b[:, 4 + k] = (rec12[:, 1] >> np.uint32(8 * k)) & np.uint32(255)
    for k in range(3):
        b[:, 8 + k] = (rec12[:, 2] >> np.uint32(8 * k)) & np.uint32(255)
    assert np.all(rec12[:, 2] < (1 << 24))
    return b, exc


def pack12(intv, pos_bits, total_bits=30):
    """bwtintv_t rows -> (records uint32[n,3], exceptions) following the comment above smem_intv12_t."""
    sg = pkg("smem_gpu")
    x0, x1, x2, info = (intv[:, k].astype(np.uint64) for k in range(4))
    qb, qe = (info >> np.uint64(32)).astype(np.uint64), (info & np.uint64(0xffffffff)).astype(np.uint64)
    fbits = total_bits - 2 * pos_bits
    esc = np.uint64((1 << fbits) - 1)
    is_exc = (x2 - np.uint64(1)) >= esc
    field = np.where(is_exc, esc, x2 - np.uint64(1))
    w2 = ((x0 >> np.uint64(32)) & np.uint64(1)) | (((x1 >> np.uint64(32)) & np.uint64(1)) << np.uint64(1)) | (qb << np.uint64(2)) \
        | ((qe - np.uint64(1)) << np.uint64(2 + pos_bits)) | (field << np.uint64(2 + 2 * pos_bits))
    rec = np.stack([x0 & np.uint64(0xffffffff), x1 & np.uint64(0xffffffff), w2], axis=1).astype(np.uint32)
    idx = np.nonzero(is_exc)[0]
    exc = np.zeros(len(idx), sg.EXC_DTYPE)
    exc["index"] = idx; exc["x2_lo"] = (x2[idx] & np.uint64(0xffffffff)).astype(np.uint32); exc["x2_hi"] = (x2[idx] >> np.uint64(32)).astype(np.uint32)
    return rec, exc[::-1].copy()                      # (the list is unordered by contract)


@pytest.mark.parametrize("pos_bits", [1, 7, 8, 10, 13])
def test_record_roundtrip_python_and_c(tmp_path, pos_bits):
    sg = pkg("smem_gpu")
    rng = np.random.default_rng(pos_bits)
    n = 5000
    max_len = 1 << pos_bits
    fbits = 30 - 2 * pos_bits
    intv = np.zeros((n, 4), np.uint64)
    intv[:, 0] = rng.integers(0, 1 << 33, n, dtype=np.uint64)
    intv[:, 1] = rng.integers(0, 1 << 33, n, dtype=np.uint64)
    x2 = rng.integers(1, (1 << fbits) + 2, n, dtype=np.uint64)              # around the escape value ...
    x2[::7] = rng.integers(1, 1 << 33, len(x2[::7]), dtype=np.uint64)       # ... and anywhere up to 2^33 - 1
    x2[:4] = [1, (1 << fbits) - 1, 1 << fbits, (1 << 33) - 1]
    intv[:, 2] = x2
    qb = rng.integers(0, max_len, n, dtype=np.uint64)
    qe = np.minimum(qb + rng.integers(1, max_len + 1, n, dtype=np.uint64), np.uint64(max_len))
    intv[:, 3] = (qb << np.uint64(32)) | qe
    rec, exc = pack12(intv, pos_bits)
    assert np.array_equal(sg.unpack_intv12(rec, pos_bits, exc), intv)
    if len(exc):
        with pytest.raises(ValueError):
            sg.unpack_intv12(rec, pos_bits, exc[:0])
    # the header's own decoder
    src = tmp_path / "u12.c"
    src.write_text('#include "smem_gpu.h"\nint u12(const smem_intv12_t *r, long n, int pb, smem_intv_t *o) { int e = 0; for (long i = 0; i < n; ++i) e += smem_intv12_unpack(r + i, pb, o + i); return e; }\n')
    so = tmp_path / "u12.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    out = np.zeros((n, 4), np.uint64)
    n_esc = lib.u12(C.c_void_p(rec.ctypes.data), C.c_long(n), C.c_int(pos_bits), C.c_void_p(out.ctypes.data))
    assert n_esc == len(exc)
    out[exc["index"], 2] = exc["x2_lo"].astype(np.uint64) | (exc["x2_hi"].astype(np.uint64) << np.uint64(32))
    assert np.array_equal(out, intv)


@pytest.mark.parametrize("pos_bits", [1, 7, 9])
def test_record11_roundtrip_python_and_c(tmp_path, pos_bits):
    sg = pkg("smem_gpu")
    rng = np.random.default_rng(100 + pos_bits)
    n = 5000
    max_len = 1 << pos_bits
    fbits = 22 - 2 * pos_bits
    intv = np.zeros((n, 4), np.uint64)
    intv[:, 0] = rng.integers(0, 1 << 33, n, dtype=np.uint64)
    intv[:, 1] = rng.integers(0, 1 << 33, n, dtype=np.uint64)
    x2 = rng.integers(1, (1 << fbits) + 2, n, dtype=np.uint64)
    x2[::5] = rng.integers(1, 1 << 33, len(x2[::5]), dtype=np.uint64)
    x2[:4] = [1, (1 << fbits) - 1, 1 << fbits, (1 << 33) - 1]
    intv[:, 2] = x2
    qb = rng.integers(0, max_len, n, dtype=np.uint64)
    qe = np.minimum(qb + rng.integers(1, max_len + 1, n, dtype=np.uint64), np.uint64(max_len))
    intv[:, 3] = (qb << np.uint64(32)) | qe
    rec, exc = pack11(intv, pos_bits)
    assert np.array_equal(sg.unpack_intv11(rec, pos_bits, exc), intv)
    src = tmp_path / "u11.c"
    src.write_text('#include "smem_gpu.h"\nint u11(const smem_intv11_t *r, long n, int pb, smem_intv_t *o) { int e = 0; for (long i = 0; i < n; ++i) e += smem_intv11_unpack(r + i, pb, o + i); return e; }\n')
    so = tmp_path / "u11.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    out = np.zeros((n, 4), np.uint64)
    n_esc = lib.u11(C.c_void_p(rec.ctypes.data), C.c_long(n), C.c_int(pos_bits), C.c_void_p(out.ctypes.data))
    assert n_esc == len(exc)
    out[exc["index"], 2] = exc["x2_lo"].astype(np.uint64) | (exc["x2_hi"].astype(np.uint64) << np.uint64(32))
    assert np.array_equal(out, intv)
