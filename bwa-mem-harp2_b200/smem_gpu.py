"""ctypes binding of the C-ABI CUDA library (include/smem_gpu.h) -- the host-side mirror used by
tests/ and bench.py.  There is deliberately no CPU fallback: if ``libsmem_gpu.so`` is missing or
no CUDA device is usable, constructing :class:`SmemGpu` raises.

Mirrors the reference interface for the path:
  * ``SmemGpu.upload_index``   <- bwa_idx_load_bwt's upload step (bwa.c:289-301)
  * ``SmemGpu.collect``        <- mem_insert_seed's enumeration loop over smem_next2 (bwamem.c:453-460)
  * ``SmemGpu.smem1``          <- one DO call of bwt_smem1_batched / bwt_smem1 (bwt.c:444, bwt.c:776)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMEM_GPU_LIB") or os.path.join(HERE, "libsmem_gpu.so")   # (SMEM_GPU_LIB: kernel-variant builds, tools/variants.sh)

EXPORTS = [
    "smem_gpu_create", "smem_gpu_destroy", "smem_gpu_resize", "smem_gpu_upload_index", "smem_gpu_upload_index_device",
    "smem_gpu_collect", "smem_gpu_smem1", "smem_gpu_upload_sa", "smem_gpu_sa", "smem_gpu_seeds", "smem_gpu_chains", "smem_gpu_trace", "smem_gpu_share_index", "smem_gpu_build_repeat_filter", "smem_gpu_get_repeat_filter", "smem_gpu_build_text_index", "smem_gpu_get_text_index", "smem_gpu_stage_reads", "smem_gpu_run_collect", "smem_gpu_fetch",
    "smem_gpu_collect_packed", "smem_gpu_stage_reads_packed", "smem_gpu_fetch_packed", "smem_gpu_pack_reads", "smem_gpu_collect_packed12", "smem_gpu_fetch_packed12", "smem_gpu_collect_packed11", "smem_gpu_fetch_packed11",
    "smem_gpu_host_alloc", "smem_gpu_host_free", "smem_gpu_last_timing", "smem_gpu_set_param",
    "smem_gpu_get_param", "smem_gpu_gather_roofline", "smem_gpu_strerror", "smem_gpu_last_error",
    "smem_gpu_device_count",
]


class SeedOpt(C.Structure):
    """smem_seed_opt_t == seeding fields of mem_opt_t (bwamem.h:33-60; defaults bwamem.c:45-75)."""
    _fields_ = [("min_seed_len", C.c_int), ("split_factor", C.c_double), ("split_width", C.c_int), ("start_width", C.c_int)]

    def __init__(self, min_seed_len=19, split_factor=1.5, split_width=10, start_width=1):
        super().__init__(min_seed_len, split_factor, split_width, start_width)


class IndexDesc(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("seed_kernel_ms", C.c_double), ("total_device_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("overflow_reads", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Reads2(C.Structure):
    """smem_reads2_t: the compact read format (two bits per base, fixed stride, ambiguous bases as an exception list)."""
    _fields_ = [("n_reads", C.c_int64), ("seq2", C.c_void_p), ("stride", C.c_int32), ("read_len", C.c_int32), ("lens", C.c_void_p),
                ("amb", C.c_void_p), ("n_amb", C.c_int64)]


AMB_DTYPE = np.dtype([("read", np.uint32), ("pos", np.uint16), ("reserved", np.uint16)])


def unpack_intv16(rec: np.ndarray) -> np.ndarray:
    """smem_intv16_t records (uint64[n, 2]) -> bwtintv_t rows (uint64[n, 4]); mirrors smem_intv16_unpack of smem_gpu.h."""
    rec = np.asarray(rec, np.uint64).reshape(-1, 2)
    w0, w1 = rec[:, 0], rec[:, 1]
    m33 = np.uint64((1 << 33) - 1)
    out = np.empty((len(rec), 4), np.uint64)
    out[:, 0] = w0 & m33
    out[:, 1] = (w0 >> np.uint64(33)) | ((w1 & np.uint64(3)) << np.uint64(31))
    out[:, 2] = (w1 >> np.uint64(2)) & m33
    out[:, 3] = (((w1 >> np.uint64(35)) & np.uint64(0x3fff)) << np.uint64(32)) | ((w1 >> np.uint64(49)) & np.uint64(0x3fff))
    return out


EXC_DTYPE = np.dtype([("index", np.uint32), ("x2_lo", np.uint32), ("x2_hi", np.uint32)])


def unpack_intv12(rec: np.ndarray, pos_bits: int, exc: "np.ndarray | None" = None) -> np.ndarray:
    """smem_intv12_t records (uint32[n, 3]) + exception list -> bwtintv_t rows (uint64[n, 4]); mirrors smem_intv12_unpack of smem_gpu.h."""
    rec = np.asarray(rec, np.uint32).reshape(-1, 3)
    w2 = rec[:, 2].astype(np.uint64)
    P = np.uint64(pos_bits)
    pm = np.uint64((1 << pos_bits) - 1)
    esc = np.uint64((1 << (30 - 2 * pos_bits)) - 1)
    f = w2 >> (np.uint64(2) + P + P)
    out = np.empty((len(rec), 4), np.uint64)
    out[:, 0] = rec[:, 0].astype(np.uint64) | ((w2 & np.uint64(1)) << np.uint64(32))
    out[:, 1] = rec[:, 1].astype(np.uint64) | (((w2 >> np.uint64(1)) & np.uint64(1)) << np.uint64(32))
    out[:, 2] = np.where(f == esc, np.uint64(0), f + np.uint64(1))
    out[:, 3] = (((w2 >> np.uint64(2)) & pm) << np.uint64(32)) | (((w2 >> (np.uint64(2) + P)) & pm) + np.uint64(1))
    n_esc = int(np.count_nonzero(f == esc))
    if n_esc:
        if exc is None or len(exc) == 0:
            raise ValueError("records refer to the exception list, which is missing")
        out[exc["index"], 2] = exc["x2_lo"].astype(np.uint64) | (exc["x2_hi"].astype(np.uint64) << np.uint64(32))
        if np.any(out[:, 2] == 0):
            raise ValueError("an escaped record has no entry in the exception list")
    return out


def unpack_intv11(rec: np.ndarray, pos_bits: int, exc: "np.ndarray | None" = None) -> np.ndarray:
    """smem_intv11_t records (uint8[n, 11]) + exception list -> bwtintv_t rows; mirrors smem_intv11_unpack of smem_gpu.h."""
    b = np.asarray(rec, np.uint8).reshape(-1, 11).astype(np.uint32)
    w = np.empty((len(b), 3), np.uint32)
    w[:, 0] = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | (b[:, 3] << 24)
    w[:, 1] = b[:, 4] | (b[:, 5] << 8) | (b[:, 6] << 16) | (b[:, 7] << 24)
    w2 = b[:, 8] | (b[:, 9] << 8) | (b[:, 10] << 16)
    # the third word of the 12-byte record with a size field of 22 - 2P instead of 30 - 2P bits: widen an all-ones field
    fbits = 22 - 2 * pos_bits
    f = w2 >> (2 + 2 * pos_bits)
    wide = np.where(f == (1 << fbits) - 1, np.uint32((1 << (30 - 2 * pos_bits)) - 1), f)
    w[:, 2] = (w2 & np.uint32((1 << (2 + 2 * pos_bits)) - 1)) | (wide.astype(np.uint32) << np.uint32(2 + 2 * pos_bits))
    return unpack_intv12(w, pos_bits, exc)


class PackedReads:
    """Host buffers of one batch in the compact format (optionally pinned) + the smem_reads2_t that describes them."""

    def __init__(self, lib, seq, offs, stride=None, uniform=None, pinned=False, threads=8):
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        lens = np.diff(offs)
        max_len = int(lens.max()) if n else 0
        self.stride = int(stride or max(1, (max_len + 3) // 4))
        if uniform is None:
            uniform = bool(n == 0 or (lens == lens[0]).all())
        self.n = n
        alloc = (lambda shape, dt: PinnedArray(lib, shape, dt)) if pinned else None
        self._keep = []

        def mk(shape, dt):
            if alloc:
                a = alloc(shape, dt); self._keep.append(a); return a.array
            return np.zeros(shape, dt)
        self.seq2 = mk((max(n * self.stride, 1),), np.uint8)
        self.lens = None if uniform else mk((max(n, 1),), np.uint16)
        n_amb_guess = int(np.count_nonzero(seq > 3))
        self.amb = mk((max(n_amb_guess, 1),), AMB_DTYPE)
        n_amb = C.c_int64(0)
        rc = lib.smem_gpu_pack_reads(C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), C.c_int32(self.stride), C.c_void_p(self.seq2.ctypes.data),
                                     C.c_void_p(self.lens.ctypes.data) if self.lens is not None else None, C.c_void_p(self.amb.ctypes.data),
                                     C.c_int64(len(self.amb)), C.byref(n_amb), C.c_int(threads))
        if rc:
            raise SmemGpuError(rc, "smem_gpu_pack_reads failed")
        self.n_amb = int(n_amb.value)
        self.desc = Reads2(n, self.seq2.ctypes.data, self.stride, int(lens[0]) if (uniform and n) else 0,
                           self.lens.ctypes.data if self.lens is not None else None, self.amb.ctypes.data if self.n_amb else None, self.n_amb)

    @property
    def h2d_bytes(self):
        return self.n * self.stride + (2 * self.n if self.lens is not None else 0) + 8 * self.n_amb


class SmemGpuError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__(f"smem_gpu: {what} ({code}){': ' + detail if detail else ''}")
        self.code = code


class ChainOpt(C.Structure):
    """smem_chain_opt_t: chaining fields of mem_opt_t (bwamem.h:33-60), defaults from mem_opt_init (bwamem.c:45-75)."""
    _fields_ = [("w", C.c_int), ("max_chain_gap", C.c_int), ("min_seed_len", C.c_int), ("mask_level", C.c_float),
                ("chain_drop_ratio", C.c_float), ("filter", C.c_int)]


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (in-tree .so, travels to the GPU box)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, "csrc", f)) > os.path.getmtime(LIB_PATH)
            for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh"))):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "csrc")], check=True)
    return LIB_PATH


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    lib.smem_gpu_strerror.restype = C.c_char_p
    lib.smem_gpu_last_error.restype = C.c_char_p
    lib.smem_gpu_get_param.restype = C.c_int64
    return lib


def pack_pac(fwd):
    """Forward text (uint8 symbols 0..3; numpy or torch) -> the reference's 2-bit .pac layout: 4 bases per byte, first
    base in the top bits (bntseq.c:_set_pac)."""
    if hasattr(fwd, "is_cuda"):
        import torch
        n = fwd.numel()
        out = torch.zeros((n + 3) // 4, dtype=torch.uint8, device=fwd.device)
        step = 1 << 30
        for s0 in range(0, n, step):
            s1 = min(n, s0 + step)
            seg = torch.zeros(((s1 - s0 + 3) // 4) * 4, dtype=torch.uint8, device=fwd.device)
            seg[: s1 - s0] = fwd[s0:s1] & 3
            q = seg.view(-1, 4)
            out[s0 // 4: s0 // 4 + q.shape[0]] = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        return out
    f = np.asarray(fwd, np.uint8) & 3
    pad = np.zeros(((f.size + 3) // 4) * 4, np.uint8)
    pad[: f.size] = f
    q = pad.reshape(-1, 4)
    return ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class PinnedArray:
    """numpy view over cudaHostAlloc memory (smem_gpu_host_alloc)."""

    def __init__(self, lib, shape, dtype):
        self.lib = lib
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        rc = lib.smem_gpu_host_alloc(C.byref(self.ptr), C.c_size_t(max(self.nbytes, 1)))
        if rc:
            raise SmemGpuError(rc, "host_alloc failed")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.smem_gpu_host_free(self.ptr)
            self.ptr = None


class SmemGpu:
    def __init__(self, max_batch_reads: int, max_read_len: int, devices: "list[int] | None" = None):
        self.lib = load_library()
        n_avail = self.lib.smem_gpu_device_count()
        if n_avail < 1:
            raise SmemGpuError(-7, "no CUDA device visible; the seeding path has no CPU fallback")
        devices = list(devices) if devices is not None else [0]
        self.h = C.c_void_p()
        ids = (C.c_int * len(devices))(*devices)
        rc = self.lib.smem_gpu_create(C.byref(self.h), len(devices), ids, C.c_int64(max_batch_reads), C.c_int(max_read_len))
        if rc:
            raise SmemGpuError(rc, self.lib.smem_gpu_strerror(rc).decode())
        self.max_batch = max_batch_reads
        self.devices = devices
        self._keep = None

    # -- helpers
    def _check(self, rc):
        if rc:
            raise SmemGpuError(rc, self.lib.smem_gpu_strerror(rc).decode(), self.lib.smem_gpu_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.smem_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, name: str, value: int):
        self._check(self.lib.smem_gpu_set_param(self.h, name.encode(), C.c_int64(value)))

    def get_param(self, name: str) -> int:
        return int(self.lib.smem_gpu_get_param(self.h, name.encode()))

    def timing(self) -> dict:
        t = Timing()
        self._check(self.lib.smem_gpu_last_timing(self.h, C.byref(t)))
        return t.asdict()

    # -- index
    def upload_index(self, index):
        """``index``: object with primary / L2 / seq_len / bwt_size / bwt (numpy array or torch tensor)."""
        bwt = index.bwt
        desc = IndexDesc(int(index.primary), (C.c_uint64 * 5)(*[int(v) for v in index.L2]), int(index.seq_len),
                         int(index.bwt_size), None)
        if hasattr(bwt, "is_cuda") and bwt.is_cuda:
            desc.bwt = bwt.data_ptr()
            import torch
            torch.cuda.synchronize(bwt.device)
            self._check(self.lib.smem_gpu_upload_index_device(self.h, C.byref(desc), C.c_int(bwt.device.index or 0)))
        else:
            words = np.ascontiguousarray(index.words_numpy(), np.uint32)
            desc.bwt = words.ctypes.data
            self._check(self.lib.smem_gpu_upload_index(self.h, C.byref(desc)))

    def share_index_from(self, other: "SmemGpu"):
        """Alias ``other``'s index / SA copies (one handle per host thread, one index per GPU)."""
        self._check(self.lib.smem_gpu_share_index(self.h, other.h))
        self._shared_from = other        # keep the owner alive

    def build_repeat_filter(self, fwd, kmer_len: int = 0, log2_bits: int = 0):
        """Repeat filter of the re-seeding pass (smem_gpu_build_repeat_filter) from the forward text ``fwd`` (uint8 symbols
        0..3, numpy array or torch tensor on any device, or an already packed .pac given as ``(pac, l_pac)``)."""
        if isinstance(fwd, tuple):
            pac, l_pac = fwd
        else:
            pac = pack_pac(fwd)
            l_pac = int(fwd.numel() if hasattr(fwd, "numel") else fwd.size)
        if hasattr(pac, "is_cuda") and pac.is_cuda:
            import torch
            torch.cuda.synchronize(pac.device)
            self._check(self.lib.smem_gpu_build_repeat_filter(self.h, C.c_void_p(pac.data_ptr()), C.c_int64(l_pac), C.c_int(pac.device.index or 0),
                                                              C.c_int(kmer_len), C.c_int(log2_bits)))
        else:
            a = np.ascontiguousarray(pac.numpy() if hasattr(pac, "numpy") else pac, np.uint8)
            self._check(self.lib.smem_gpu_build_repeat_filter(self.h, C.c_void_p(a.ctypes.data), C.c_int64(l_pac), C.c_int(-1), C.c_int(kmer_len), C.c_int(log2_bits)))

    def build_text_index(self, fwd):
        """Unique-walk tables (smem_gpu_build_text_index): text (4 bits per base) + full SA + inverse SA; needs upload_index and upload_sa.
        ``fwd`` as for :meth:`build_repeat_filter`."""
        if isinstance(fwd, tuple):
            pac, l_pac = fwd
        else:
            pac = pack_pac(fwd)
            l_pac = int(fwd.numel() if hasattr(fwd, "numel") else fwd.size)
        if hasattr(pac, "is_cuda") and pac.is_cuda:
            import torch
            torch.cuda.synchronize(pac.device)
            self._check(self.lib.smem_gpu_build_text_index(self.h, C.c_void_p(pac.data_ptr()), C.c_int64(l_pac), C.c_int(pac.device.index or 0)))
        else:
            a = np.ascontiguousarray(pac.numpy() if hasattr(pac, "numpy") else pac, np.uint8)
            self._check(self.lib.smem_gpu_build_text_index(self.h, C.c_void_p(a.ctypes.data), C.c_int64(l_pac), C.c_int(-1)))
        self._text_len = 2 * l_pac

    def text_index(self, which: int) -> np.ndarray:
        """Test hook: 0 = full suffix array (seq_len + 1 rows), 1 = the sampled inverse (entry e = row of the suffix at
        seq_len - e * 2^unique_walk_isa_shift)."""
        out = np.empty(self._text_len + 1 if which == 0 else (self._text_len >> self.get_param("unique_walk_isa_shift")) + 1, np.uint64)
        self._check(self.lib.smem_gpu_get_text_index(self.h, C.c_int(which), C.c_void_p(out.ctypes.data), C.c_int64(out.size)))
        return out

    def repeat_filter_bits(self) -> np.ndarray:
        """Test hook: the bit table of device 0 as uint32 words."""
        out = np.empty(1 << (self.get_param("rf_log2_bits") - 5), np.uint32)
        self._check(self.lib.smem_gpu_get_repeat_filter(self.h, C.c_void_p(out.ctypes.data), C.c_int64(out.size)))
        return out

    def upload_sa(self, index):
        """Suffix-array samples of ``index`` (sa_intv, sa) -> HBM, for :meth:`sa`."""
        sa = index.sa
        if hasattr(sa, "is_cuda") and sa.is_cuda:
            import torch
            torch.cuda.synchronize(sa.device)
            self._check(self.lib.smem_gpu_upload_sa(self.h, C.c_int(int(index.sa_intv)), C.c_uint64(sa.numel()), C.c_void_p(sa.data_ptr()),
                                                    C.c_int(sa.device.index or 0)))
        else:
            a = np.ascontiguousarray(index.sa_numpy(), np.uint64)
            self._check(self.lib.smem_gpu_upload_sa(self.h, C.c_int(int(index.sa_intv)), C.c_uint64(a.size), C.c_void_p(a.ctypes.data), C.c_int(-1)))

    def sa(self, k):
        """bwt_sa (bwt.c:104) for a batch of suffix-array rows."""
        k = np.ascontiguousarray(k, np.uint64)
        out = np.zeros(len(k), np.uint64)
        self._check(self.lib.smem_gpu_sa(self.h, C.c_int64(len(k)), _p(k, C.c_uint64), _p(out, C.c_uint64)))
        return out

    SEED_DTYPE = np.dtype([("rbeg", np.int64), ("qbeg", np.int32), ("len", np.int32)])

    def seeds(self, n_reads: int, min_seed_len=19, max_occ=10000):
        """mem_seed_t list of the last run's intervals (bwamem.c:408-424): structured array + CSR offsets per read."""
        seed_off = np.zeros(n_reads + 1, np.int64)
        tot = C.c_int64(0)
        cap = max(64, 4 * n_reads)
        while True:
            out = np.zeros(cap, self.SEED_DTYPE)
            rc = self.lib.smem_gpu_seeds(self.h, C.c_int(min_seed_len), C.c_int64(max_occ), C.c_void_p(out.ctypes.data), C.c_int64(cap),
                                         _p(seed_off, C.c_int64), C.byref(tot))
            if rc == -5 and tot.value > cap:
                cap = int(tot.value)
                continue
            self._check(rc)
            return dict(seeds=out[:int(tot.value)], seed_off=seed_off)

    CHAIN_DTYPE = np.dtype([("pos", np.int64), ("seed_first", np.int64), ("n_seeds", np.int32), ("weight", np.int32)])

    def chains(self, n_reads: int, l_pac: int, w=100, max_chain_gap=10000, min_seed_len=19, mask_level=0.5, chain_drop_ratio=0.5,
               flt=True, fetch=True):
        """mem_chain [+ mem_chain_flt] (bwamem.c:593, 629) on the resident seeds of the last :meth:`seeds` call."""
        opt = ChainOpt(w, max_chain_gap, min_seed_len, mask_level, chain_drop_ratio, int(flt))
        chain_off = np.zeros(n_reads + 1, np.int64)
        nc, ns = C.c_int64(0), C.c_int64(0)
        ccap = scap = 0
        chains = seeds = None
        while True:
            rc = self.lib.smem_gpu_chains(self.h, C.byref(opt), C.c_int64(int(l_pac)), C.c_void_p(chains.ctypes.data) if chains is not None else None,
                                          C.c_int64(ccap), C.c_void_p(seeds.ctypes.data) if seeds is not None else None, C.c_int64(scap),
                                          _p(chain_off, C.c_int64), C.byref(nc), C.byref(ns))
            if rc == -5 and fetch and (nc.value > ccap or ns.value > scap):
                ccap, scap = max(int(nc.value), 1), max(int(ns.value), 1)
                chains, seeds = np.zeros(ccap, self.CHAIN_DTYPE), np.zeros(scap, self.SEED_DTYPE)
                continue
            if rc == -5 and not fetch:
                return dict(chain_off=chain_off, n_chains=int(nc.value), n_seeds=int(ns.value))
            self._check(rc)
            if chains is None:
                chains, seeds = np.zeros(0, self.CHAIN_DTYPE), np.zeros(0, self.SEED_DTYPE)
            return dict(chain_off=chain_off, chains=chains[:int(nc.value)], seeds=seeds[:int(ns.value)])

    # -- one-call forms (host buffers in, host buffers out)
    def collect(self, seq, offs, opt: "SeedOpt | None" = None, want_step=True, cap_hint: "int | None" = None):
        opt = opt or SeedOpt()
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        cap = cap_hint or max(64, 16 * n)
        read_off = np.zeros(n + 1, np.int64)
        tot = C.c_int64(0)
        while True:
            intv = np.empty((cap, 4), np.uint64)
            step = np.empty(cap, np.uint16) if want_step else None
            rc = self.lib.smem_gpu_collect(self.h, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), C.byref(opt),
                                           _p(intv, C.c_uint64), C.c_int64(cap), _p(read_off, C.c_int64),
                                           _p(step, C.c_uint16), C.byref(tot))
            if rc == -5 and tot.value > cap:
                cap = int(tot.value)
                continue
            self._check(rc)
            t = int(tot.value)
            return dict(intv=intv[:t], read_off=read_off, step=step[:t] if want_step else None)

    def collect_packed(self, reads: "PackedReads", opt: "SeedOpt | None" = None, out=None, read_off=None, unpack=True):
        """smem_gpu_collect_packed: compact reads in, 16-byte interval records + uint32 CSR offsets out."""
        opt = opt or SeedOpt()
        n = reads.n
        cap = out.shape[0] if out is not None else max(64, 16 * n)
        read_off = read_off if read_off is not None else np.zeros(n + 1, np.uint32)
        tot = C.c_int64(0)
        while True:
            rec = out if out is not None else np.empty((cap, 2), np.uint64)
            rc = self.lib.smem_gpu_collect_packed(self.h, C.byref(reads.desc), C.byref(opt), C.c_void_p(rec.ctypes.data), C.c_int64(cap),
                                                  C.c_void_p(read_off.ctypes.data), C.byref(tot))
            if rc == -5 and tot.value > cap and out is None:
                cap = int(tot.value)
                continue
            self._check(rc)
            t = int(tot.value)
            return dict(rec=rec[:t], intv=unpack_intv16(rec[:t]) if unpack else None, read_off=read_off.astype(np.int64) if unpack else read_off)

    def collect_packed11(self, reads: "PackedReads", opt: "SeedOpt | None" = None, **kw):
        """smem_gpu_collect_packed11: like collect_packed12 with byte-packed 11-byte records."""
        return self.collect_packed12(reads, opt, rec_bytes=11, **kw)

    def collect_packed12(self, reads: "PackedReads", opt: "SeedOpt | None" = None, out=None, read_off=None, exc=None, unpack=True, rec_bytes=12):
        """smem_gpu_collect_packed12: compact reads in, 12-byte interval records + exception list + uint32 CSR offsets out."""
        opt = opt or SeedOpt()
        fn = self.lib.smem_gpu_collect_packed12 if rec_bytes == 12 else self.lib.smem_gpu_collect_packed11
        shape, dt, unp = ((3,), np.uint32, unpack_intv12) if rec_bytes == 12 else ((11,), np.uint8, unpack_intv11)
        n = reads.n
        cap = out.shape[0] if out is not None else max(64, 16 * n)
        exc_cap = exc.shape[0] if exc is not None else max(64, n)
        read_off = read_off if read_off is not None else np.zeros(n + 1, np.uint32)
        tot, n_exc, pb = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        while True:
            rec = out if out is not None else np.empty((cap,) + shape, dt)
            ex = exc if exc is not None else np.empty(exc_cap, EXC_DTYPE)
            rc = fn(self.h, C.byref(reads.desc), C.byref(opt), C.c_void_p(rec.ctypes.data), C.c_int64(cap),
                                                    C.c_void_p(read_off.ctypes.data), C.c_void_p(ex.ctypes.data), C.c_int64(exc_cap), C.byref(n_exc),
                                                    C.byref(pb), C.byref(tot))
            if rc == -5 and out is None and exc is None and (tot.value > cap or n_exc.value > exc_cap):
                cap, exc_cap = max(cap, int(tot.value)), max(exc_cap, int(n_exc.value))
                continue
            self._check(rc)
            t, ne = int(tot.value), int(n_exc.value)
            return dict(rec=rec[:t], exc=ex[:ne], pos_bits=int(pb.value), intv=unp(rec[:t], int(pb.value), ex[:ne]) if unpack else None,
                        read_off=read_off.astype(np.int64) if unpack else read_off)

    def fetch_packed11(self, total: int):
        return self.fetch_packed12(total, rec_bytes=11)

    def fetch_packed12(self, total: int, rec_bytes=12):
        n = self._n
        fn = self.lib.smem_gpu_fetch_packed12 if rec_bytes == 12 else self.lib.smem_gpu_fetch_packed11
        shape, dt, unp = ((3,), np.uint32, unpack_intv12) if rec_bytes == 12 else ((11,), np.uint8, unpack_intv11)
        read_off = np.zeros(n + 1, np.uint32)
        tot, n_exc, pb = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        exc_cap = 1024
        while True:
            rec = np.empty((max(total, 1),) + shape, dt)
            ex = np.empty(exc_cap, EXC_DTYPE)
            rc = fn(self.h, C.c_void_p(rec.ctypes.data), C.c_int64(rec.shape[0]), C.c_void_p(read_off.ctypes.data),
                                                  C.c_void_p(ex.ctypes.data), C.c_int64(exc_cap), C.byref(n_exc), C.byref(pb), C.byref(tot))
            if rc == -5 and n_exc.value > exc_cap:
                exc_cap = int(n_exc.value)
                continue
            self._check(rc)
            t, ne = int(tot.value), int(n_exc.value)
            return dict(rec=rec[:t], exc=ex[:ne], pos_bits=int(pb.value), intv=unp(rec[:t], int(pb.value), ex[:ne]), read_off=read_off.astype(np.int64))

    def stage_packed(self, reads: "PackedReads"):
        self._keep = reads
        self._n = reads.n
        self._check(self.lib.smem_gpu_stage_reads_packed(self.h, C.byref(reads.desc)))

    def fetch_packed(self, total: int):
        n = self._n
        rec = np.empty((max(total, 1), 2), np.uint64)
        read_off = np.zeros(n + 1, np.uint32)
        tot = C.c_int64(0)
        self._check(self.lib.smem_gpu_fetch_packed(self.h, C.c_void_p(rec.ctypes.data), C.c_int64(rec.shape[0]), C.c_void_p(read_off.ctypes.data), C.byref(tot)))
        t = int(tot.value)
        return dict(rec=rec[:t], intv=unpack_intv16(rec[:t]), read_off=read_off.astype(np.int64))

    def trace(self, seq, offs, opt: "SeedOpt | None" = None):
        """Per-call raw lists of the whole iterator walk (smem_gpu_trace): intv, read_off, tag (2*step+pass), ret."""
        opt = opt or SeedOpt()
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        cap = max(64, 32 * n)
        read_off = np.zeros(n + 1, np.int64)
        tot = C.c_int64(0)
        while True:
            intv = np.empty((cap, 4), np.uint64)
            tag = np.empty(cap, np.uint16)
            ret = np.empty(cap, np.uint16)
            rc = self.lib.smem_gpu_trace(self.h, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), C.byref(opt), _p(intv, C.c_uint64),
                                         C.c_int64(cap), _p(read_off, C.c_int64), _p(tag, C.c_uint16), _p(ret, C.c_uint16), C.byref(tot))
            if rc == -5 and tot.value > cap:
                cap = int(tot.value)
                continue
            self._check(rc)
            t = int(tot.value)
            return dict(intv=intv[:t], read_off=read_off, tag=tag[:t], ret=ret[:t])

    def smem1(self, seq, offs, x, min_intv):
        seq = np.ascontiguousarray(seq, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        x = np.ascontiguousarray(x, np.int32)
        mi = np.ascontiguousarray(min_intv, np.int32)
        n = len(offs) - 1
        cap = max(64, 16 * n)
        read_off = np.zeros(n + 1, np.int64)
        ret = np.zeros(max(n, 1), np.int32)
        tot = C.c_int64(0)
        while True:
            intv = np.empty((cap, 4), np.uint64)
            rc = self.lib.smem_gpu_smem1(self.h, C.c_int64(n), _p(seq, C.c_uint8), _p(offs, C.c_int64), _p(x, C.c_int32),
                                         _p(mi, C.c_int32), _p(intv, C.c_uint64), C.c_int64(cap), _p(read_off, C.c_int64),
                                         _p(ret, C.c_int32), C.byref(tot))
            if rc == -5 and tot.value > cap:
                cap = int(tot.value)
                continue
            self._check(rc)
            return dict(intv=intv[:int(tot.value)], read_off=read_off, ret=ret[:n])

    # -- split form (device-resident timing)
    def stage(self, seq, offs):
        seq = np.ascontiguousarray(seq, np.uint8) if not isinstance(seq, np.ndarray) or not seq.flags.c_contiguous else seq
        offs = np.ascontiguousarray(offs, np.int64)
        self._keep = (seq, offs)
        self._n = len(offs) - 1
        self._check(self.lib.smem_gpu_stage_reads(self.h, C.c_int64(self._n), _p(seq, C.c_uint8), _p(offs, C.c_int64)))

    def run_collect(self, opt: "SeedOpt | None" = None) -> int:
        opt = opt or SeedOpt()
        tot = C.c_int64(0)
        self._check(self.lib.smem_gpu_run_collect(self.h, C.byref(opt), C.byref(tot)))
        return int(tot.value)

    def fetch(self, total: int, want_step=False, intv=None, read_off=None):
        n = self._n
        intv = intv if intv is not None else np.empty((max(total, 1), 4), np.uint64)
        read_off = read_off if read_off is not None else np.zeros(n + 1, np.int64)
        step = np.empty(max(total, 1), np.uint16) if want_step else None
        tot = C.c_int64(0)
        self._check(self.lib.smem_gpu_fetch(self.h, _p(intv, C.c_uint64), C.c_int64(intv.shape[0]), _p(read_off, C.c_int64),
                                            _p(step, C.c_uint16), C.byref(tot)))
        t = int(tot.value)
        return dict(intv=intv[:t], read_off=read_off, step=step[:t] if want_step else None)

    def gather_roofline(self, block_bytes=64, span_bytes=0, chains_per_sm=2048, steps=2000) -> float:
        g = C.c_double(0)
        self._check(self.lib.smem_gpu_gather_roofline(self.h, C.c_int(block_bytes), C.c_uint64(span_bytes), C.c_int(chains_per_sm),
                                                      C.c_int(steps), C.byref(g)))
        return float(g.value)
