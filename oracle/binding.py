"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

* ``Oracle``    -> oracle/liboracle.so   (the C restatement, smem_oracle.c)
* ``Reference`` -> oracle/_ref/libbwaref.so (the reference's own objects + ref_harness.c);
  present only where ``make -C oracle ref`` ran or where the prebuilt file travelled.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/software"


class SeedOpt(C.Structure):
    """mem_opt_t seeding fields (bwamem.h:33-60), defaults from mem_opt_init (bwamem.c:45-75)."""
    _fields_ = [("min_seed_len", C.c_int), ("split_factor", C.c_double), ("split_width", C.c_int), ("start_width", C.c_int)]

    def __init__(self, min_seed_len=19, split_factor=1.5, split_width=10, start_width=1):
        super().__init__(min_seed_len, split_factor, split_width, start_width)


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("extends", "blocks", "smem1_calls", "steps", "intervals", "max_curr", "max_mem")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


SEED_DT = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4")])     # == mem_seed_t (bwamem.c:316-319)


class ChainOpt(C.Structure):
    """Chaining / chain-filter fields of mem_opt_t (bwamem.h:33-60), defaults from mem_opt_init (bwamem.c:45-75)."""
    _fields_ = [("w", C.c_int), ("max_chain_gap", C.c_int), ("min_seed_len", C.c_int), ("mask_level", C.c_float), ("chain_drop_ratio", C.c_float)]

    def __init__(self, w=100, max_chain_gap=10000, min_seed_len=19, mask_level=0.5, chain_drop_ratio=0.5):
        super().__init__(w, max_chain_gap, min_seed_len, mask_level, chain_drop_ratio)


class _RefChainOpt(C.Structure):
    _fields_ = [("seed", SeedOpt), ("w", C.c_int), ("max_chain_gap", C.c_int), ("max_occ", C.c_int), ("mask_level", C.c_float),
                ("chain_drop_ratio", C.c_float)]


class _OrcIndex(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.c_void_p)]


def build_oracle() -> str:
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return os.path.join(HERE, "liboracle.so")


def build_reference() -> "str | None":
    """Compile the reference where its sources lie; returns the .so path or None if unavailable."""
    so = os.path.join(HERE, "_ref", "libbwaref.so")
    if os.path.isdir(REF_SRC):
        subprocess.run(["make", "-s", "-C", HERE, "ref", "-j8"], check=True)
        if os.path.exists(os.path.join(HERE, "..", "bwa-mem-harp2_b200", "libsmem_gpu.so")):
            subprocess.run(["make", "-s", "-C", HERE, "gpu", "harp", "timed", "-j8"], check=True)   # reference `bwa mem` + GPU adapter / + protocol shim / timed CPU twin
    return so if os.path.exists(so) else None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class _Base:
    """Shared flat-array plumbing: both checkers expose the same collect/smem1/time signatures."""

    def _collect(self, fn, handle, seq, offs, opt, nthreads, extra=()):
        n = len(offs) - 1
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        cap = max(64, 24 * n)
        read_off = np.zeros(n + 1, np.int64)
        n_steps = np.zeros(max(n, 1), np.int32)
        last = np.zeros(max(n, 1), np.int32)
        while True:
            intv = np.zeros((cap, 4), np.uint64)
            step = np.zeros(cap, np.uint16)
            tot = fn(handle, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), C.byref(opt), nthreads, _p(intv, C.c_uint64), C.c_int64(cap),
                     _p(read_off, C.c_int64), _p(step, C.c_uint16), _p(n_steps, C.c_int32), _p(last, C.c_int32), *extra)
            if tot <= cap:
                return dict(intv=intv[:tot], read_off=read_off, step=step[:tot], n_steps=n_steps[:n], last_start=last[:n])
            cap = int(tot)

    def _smem1(self, fn, handle, seq, offs, x, min_intv):
        n = len(offs) - 1
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        x = np.ascontiguousarray(x, np.int32)
        mi = np.ascontiguousarray(min_intv, np.int32)
        cap = max(64, 32 * n)
        read_off = np.zeros(n + 1, np.int64)
        ret = np.zeros(max(n, 1), np.int32)
        while True:
            intv = np.zeros((cap, 4), np.uint64)
            tot = fn(handle, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), _p(x, C.c_int32), _p(mi, C.c_int32),
                     _p(intv, C.c_uint64), C.c_int64(cap), _p(read_off, C.c_int64), _p(ret, C.c_int32))
            if tot <= cap:
                return dict(intv=intv[:tot], read_off=read_off, ret=ret[:n])
            cap = int(tot)

    def _time(self, fn, handle, seq, offs, opt, nthreads):
        n = len(offs) - 1
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        ck = C.c_uint64(0)
        ni = C.c_int64(0)
        sec = fn(handle, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), C.byref(opt), nthreads, C.byref(ck), C.byref(ni))
        return dict(seconds=float(sec), checksum=int(ck.value), n_intervals=int(ni.value), reads_per_s=n / sec if sec > 0 else 0.0)


class Oracle(_Base):
    """The C restatement.  ``index`` is any object with primary/L2/seq_len/bwt_size/words_numpy()."""

    def __init__(self, index):
        self.lib = C.CDLL(build_oracle())
        self.words = np.ascontiguousarray(index.words_numpy(), np.uint32)
        self.ix = _OrcIndex(int(index.primary), (C.c_uint64 * 5)(*[int(v) for v in index.L2]), int(index.seq_len),
                            int(index.bwt_size), self.words.ctypes.data)
        L = self.lib
        L.orc_collect.restype = C.c_int64
        L.orc_smem1.restype = C.c_int64
        L.orc_time_collect.restype = C.c_double
        L.orc_checksum.restype = C.c_uint64

    def collect(self, seq, offs, opt=None, nthreads=1, stats=False):
        opt = opt or SeedOpt()
        st = Stats()
        out = self._collect(self.lib.orc_collect, C.byref(self.ix), seq, offs, opt, C.c_int(nthreads), (C.byref(st) if stats else None,))
        if stats:
            out["stats"] = st.asdict()
        return out

    def smem1(self, seq, offs, x, min_intv):
        return self._smem1(self.lib.orc_smem1, C.byref(self.ix), seq, offs, x, min_intv)

    def time_collect(self, seq, offs, opt=None, nthreads=1):
        return self._time(self.lib.orc_time_collect, C.byref(self.ix), seq, offs, opt or SeedOpt(), C.c_int(nthreads))

    def occ4(self, k: int):
        cnt = (C.c_uint64 * 4)()
        self.lib.orc_occ4(C.byref(self.ix), C.c_uint64(k & 0xFFFFFFFFFFFFFFFF), cnt)
        return [int(v) for v in cnt]

    def extend(self, ik3, is_back: int):
        ok = (C.c_uint64 * 12)()
        self.lib.orc_extend(C.byref(self.ix), (C.c_uint64 * 3)(*[int(v) for v in ik3]), C.c_int(is_back), ok)
        return np.array(list(ok), dtype=np.uint64).reshape(4, 3)

    def sa(self, index, k):
        """bwt_sa for every k; ``index`` carries sa_intv / sa samples."""
        k = np.ascontiguousarray(k, np.uint64)
        sa = np.ascontiguousarray(index.sa_numpy(), np.uint64)
        out = np.zeros(len(k), np.uint64)
        self.lib.orc_sa(C.byref(self.ix), C.c_int(int(index.sa_intv)), _p(sa, C.c_uint64), C.c_int64(len(k)), _p(k, C.c_uint64), _p(out, C.c_uint64))
        return out

    def seeds(self, index, intv, read_off, min_seed_len=19, max_occ=10000):
        """Intervals -> seeds (bwamem.c:462-476); ``index`` carries the SA samples."""
        intv = np.ascontiguousarray(intv, np.uint64)
        read_off = np.ascontiguousarray(read_off, np.int64)
        sa = np.ascontiguousarray(index.sa_numpy(), np.uint64)
        n = len(read_off) - 1
        seed_off = np.zeros(n + 1, np.int64)
        self.lib.orc_seeds.restype = C.c_int64
        cap = max(64, 4 * n)
        while True:
            seeds = np.zeros(cap, SEED_DT)
            tot = self.lib.orc_seeds(C.byref(self.ix), C.c_int(int(index.sa_intv)), _p(sa, C.c_uint64), C.c_int64(n), _p(intv, C.c_uint64),
                                     _p(read_off, C.c_int64), C.c_int(min_seed_len), C.c_int64(max_occ), C.c_void_p(seeds.ctypes.data),
                                     C.c_int64(cap), _p(seed_off, C.c_int64))
            if tot <= cap:
                return dict(seeds=seeds[:tot], seed_off=seed_off)
            cap = int(tot)

    def chains(self, seeds, seed_off, l_pac, copt=None, flt=True):
        """Seeds -> chains: mem_chain's insertion loop + (flt) mem_chain_flt."""
        copt = copt or ChainOpt()
        seeds = np.ascontiguousarray(seeds, SEED_DT)
        seed_off = np.ascontiguousarray(seed_off, np.int64)
        n = len(seed_off) - 1
        self.lib.orc_chains.restype = C.c_int64
        cap = max(len(seeds), 1)
        chain_off = np.zeros(n + 1, np.int64)
        chain = np.zeros((cap, 2), np.int64)
        out = np.zeros(cap, SEED_DT)
        ns = C.c_int64(0)
        nc = self.lib.orc_chains(C.c_int64(n), C.c_void_p(seeds.ctypes.data), _p(seed_off, C.c_int64), C.c_int64(int(l_pac)), C.byref(copt),
                                 C.c_int(int(flt)), _p(chain_off, C.c_int64), _p(chain, C.c_int64), C.c_int64(cap), C.c_void_p(out.ctypes.data),
                                 C.c_int64(cap), C.byref(ns))
        assert nc <= cap and ns.value <= cap
        return dict(chain_off=chain_off, chain=chain[:nc], seeds=out[:ns.value])

    def checksum(self, intv, read_off):
        intv = np.ascontiguousarray(intv, np.uint64)
        read_off = np.ascontiguousarray(read_off, np.int64)
        return int(self.lib.orc_checksum(C.c_int64(len(read_off) - 1), _p(intv, C.c_uint64), _p(read_off, C.c_int64)))


class Reference(_Base):
    """The reference's own compiled objects.  Raises FileNotFoundError when they are not available."""

    def __init__(self, index=None, path: "str | None" = None):
        so = build_reference()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libbwaref.so is not built and /root/reference is absent")
        self.lib = L = C.CDLL(so)
        L.ref_bwt_load.restype = C.c_void_p
        L.ref_bwt_view.restype = C.c_void_p
        L.ref_collect.restype = C.c_int64
        L.ref_smem1.restype = C.c_int64
        L.ref_time_collect.restype = C.c_double
        self._view = path is None
        if path is not None:
            self.h = C.c_void_p(L.ref_bwt_load(path.encode()))
        else:
            self.words = np.ascontiguousarray(index.words_numpy(), np.uint32)
            self.h = C.c_void_p(L.ref_bwt_view(C.c_uint64(int(index.primary)), (C.c_uint64 * 5)(*[int(v) for v in index.L2]),
                                               C.c_uint64(int(index.seq_len)), C.c_uint64(int(index.bwt_size)),
                                               C.c_void_p(self.words.ctypes.data)))

    def __del__(self):
        try:
            (self.lib.ref_bwt_free_view if self._view else self.lib.ref_bwt_free)(self.h)
        except Exception:
            pass

    def collect(self, seq, offs, opt=None, nthreads=1):
        return self._collect(self.lib.ref_collect, self.h, seq, offs, opt or SeedOpt(), C.c_int(nthreads))

    def smem1(self, seq, offs, x, min_intv):
        return self._smem1(self.lib.ref_smem1, self.h, seq, offs, x, min_intv)

    def time_collect(self, seq, offs, opt=None, nthreads=1):
        return self._time(self.lib.ref_time_collect, self.h, seq, offs, opt or SeedOpt(), C.c_int(nthreads))

    def occ4(self, k: int):
        cnt = (C.c_uint64 * 4)()
        self.lib.ref_occ4(self.h, C.c_uint64(k & 0xFFFFFFFFFFFFFFFF), cnt)
        return [int(v) for v in cnt]

    def sa(self, index, k):
        k = np.ascontiguousarray(k, np.uint64)
        sa = np.ascontiguousarray(index.sa_numpy(), np.uint64)
        out = np.zeros(len(k), np.uint64)
        self.lib.ref_bwt_set_sa(self.h, C.c_int(int(index.sa_intv)), C.c_uint64(len(sa)), C.c_void_p(sa.ctypes.data))
        self.lib.ref_sa(self.h, C.c_int64(len(k)), _p(k, C.c_uint64), _p(out, C.c_uint64))
        self.lib.ref_bwt_clear_sa(self.h)
        return out

    def chains(self, index, seq, offs, opt=None, copt=None, max_occ=10000, flt=True):
        """The reference's own mem_chain (+ mem_chain_flt) per read; ``index`` carries the SA samples."""
        opt, copt = opt or SeedOpt(), copt or ChainOpt()
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        sa = np.ascontiguousarray(index.sa_numpy(), np.uint64)
        n = len(offs) - 1
        ro = _RefChainOpt(opt, copt.w, copt.max_chain_gap, max_occ, copt.mask_level, copt.chain_drop_ratio)
        ro.seed.min_seed_len = copt.min_seed_len = opt.min_seed_len
        self.lib.ref_chains.restype = C.c_int64
        self.lib.ref_bwt_set_sa(self.h, C.c_int(int(index.sa_intv)), C.c_uint64(len(sa)), C.c_void_p(sa.ctypes.data))
        ccap, scap = max(64, 4 * n), max(64, 8 * n)
        chain_off = np.zeros(n + 1, np.int64)
        try:
            while True:
                chain = np.zeros((ccap, 2), np.int64)
                seeds = np.zeros(scap, SEED_DT)
                ns = C.c_int64(0)
                nc = self.lib.ref_chains(self.h, C.c_int64(int(index.seq_len) // 2), C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64),
                                         C.byref(ro), C.c_int(int(flt)), _p(chain_off, C.c_int64), _p(chain, C.c_int64), C.c_int64(ccap),
                                         C.c_void_p(seeds.ctypes.data), C.c_int64(scap), C.byref(ns))
                if nc <= ccap and ns.value <= scap:
                    return dict(chain_off=chain_off, chain=chain[:nc], seeds=seeds[:ns.value])
                ccap, scap = max(ccap, int(nc)), max(scap, int(ns.value))
        finally:
            self.lib.ref_bwt_clear_sa(self.h)

    def extend(self, ik3, is_back: int):
        ok = (C.c_uint64 * 12)()
        self.lib.ref_extend(self.h, (C.c_uint64 * 3)(*[int(v) for v in ik3]), C.c_int(is_back), ok)
        return np.array(list(ok), dtype=np.uint64).reshape(4, 3)
