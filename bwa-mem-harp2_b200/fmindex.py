"""FM-index construction in the reference's packed ``bwt_t`` layout.

This is workload tooling (SURVEY.md section 2.1 row S9 is OUT OF SCOPE for the hot path):
``bwa index`` needs ~2 min for 100 Mbp and >1 h for 3.1 Gbp single-threaded
(bwtindex.c:239-246), which the benchmark box cannot afford, so the synthetic
indices are built here with sort primitives (torch, on whatever device the
sequence tensor lives on) and verified bit-for-bit against the reference's own
``bwa index`` output at the sizes where that is affordable (tests/test_fmindex.py).

Layout produced (what ``bwt_restore_bwt`` loads, bwt.c:899-918; built by
``bwt_bwtupdate_core``, bwtindex.c:128-150):
  * text T = forward pac + its reverse complement (bntseq.c:268-273), ``seq_len = len(T)``
  * BWT of T$ with the ``$`` row removed; ``primary`` = row of the suffix starting at 0
  * per 128 symbols one 64-byte block: 4 x uint64 running counts (A,C,G,T before the
    block) followed by 8 x uint32 words of 16 symbols, first symbol in the top bits;
    the last block is truncated to ceil(rem/16) words and followed by the final counts.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np
import torch

KEY_SYMS = 31            # symbols per sort key (62 bits, keeps int64 keys non-negative)
_NONZERO_CHUNK = 1 << 30


@dataclass
class BwtIndex:
    """Host-side mirror of the fields of ``bwt_t`` (bwt.h:46-58) that the seeding path reads."""

    primary: int
    L2: np.ndarray        # uint64[5]
    seq_len: int
    bwt_size: int         # in uint32 words
    bwt: "np.ndarray | torch.Tensor"   # uint32/int32 words, length bwt_size
    sa_intv: int = 0      # 0 = no suffix-array samples
    sa: "np.ndarray | torch.Tensor | None" = None   # uint64/int64 [n_sa]: SA of every sa_intv-th row, sa[0] = -1 (bwt.c:79-101)

    def sa_numpy(self) -> np.ndarray:
        if isinstance(self.sa, torch.Tensor):
            return self.sa.cpu().numpy().view(np.uint64)
        return np.asarray(self.sa).view(np.uint64)

    def save_sa(self, path: str) -> None:
        """The reference's ``.sa`` file format (bwt_dump_sa, bwt.c:852-864)."""
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", self.primary))
            f.write(np.asarray(self.L2[1:5], dtype="<u8").tobytes())
            f.write(struct.pack("<QQ", self.sa_intv, self.seq_len))
            self.sa_numpy()[1:].tofile(f)

    @staticmethod
    def load_sa(path: str):
        """Returns (sa_intv, sa uint64[n_sa]) from a reference ``.sa`` file (bwt_restore_sa, bwt.c:877-897)."""
        with open(path, "rb") as f:
            f.read(40)
            intv, seq_len = struct.unpack("<QQ", f.read(16))
            rest = np.fromfile(f, dtype=np.uint64)
        sa = np.empty(rest.size + 1, dtype=np.uint64)
        sa[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
        sa[1:] = rest
        assert sa.size == (seq_len + intv) // intv
        return int(intv), sa

    def words_numpy(self) -> np.ndarray:
        if isinstance(self.bwt, torch.Tensor):
            return self.bwt.cpu().numpy().view(np.uint32)
        return self.bwt.view(np.uint32)

    def save(self, path: str) -> None:
        """Write the reference's ``.bwt`` file format (bwt_dump_bwt, bwt.c:841-850)."""
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", self.primary))
            f.write(np.asarray(self.L2[1:5], dtype="<u8").tobytes())
            self.words_numpy().tofile(f)

    @staticmethod
    def load(path: str) -> "BwtIndex":
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            primary = struct.unpack("<Q", f.read(8))[0]
            l2 = np.zeros(5, dtype=np.uint64)
            l2[1:] = np.frombuffer(f.read(32), dtype="<u8")
            words = np.fromfile(f, dtype=np.uint32)
        assert words.size == (size - 40) // 4
        return BwtIndex(primary, l2, int(l2[4]), int(words.size), words)


def text_from_forward(fwd: torch.Tensor) -> torch.Tensor:
    """T = fwd + revcomp(fwd), symbols 0..3 (bntseq.c:268-273)."""
    return torch.cat([fwd, (3 - fwd).flip(0)])


def _pack64(T: torch.Tensor, n: int) -> torch.Tensor:
    """32 symbols per int64 word, first symbol in the top bits; two zero words of padding."""
    dev = T.device
    nw = (n + 31) // 32 + 2
    out = torch.empty(nw, dtype=torch.int64, device=dev)
    chunk = 1 << 28  # words per pass bound the transient memory
    for w0 in range(0, nw, chunk):
        w1 = min(nw, w0 + chunk)
        seg = torch.zeros((w1 - w0) * 32, dtype=torch.uint8, device=dev)
        s0, s1 = w0 * 32, min(n, w1 * 32)
        if s1 > s0:
            seg[: s1 - s0] = T[s0:s1]
        q = seg.view(-1, 4)
        b = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        del seg, q
        out[w0:w1] = b.view(-1, 8).flip(1).contiguous().view(torch.int64).view(-1)
        del b
    return out


def _key31(P: torch.Tensor, pos: torch.Tensor, n: int) -> torch.Tensor:
    """Sort key of the suffix at ``pos``: its first 31 symbols (zero padded past the end);
    suffixes that start at or beyond ``n`` get negative keys, earlier end = smaller."""
    inside = pos < n
    p = torch.where(inside, pos, torch.zeros_like(pos))
    w = p >> 5
    s2 = (p & 31) << 1
    hi = P[w] << s2
    lo = ((P[w + 1] >> 1) & 0x7FFFFFFFFFFFFFFF) >> (63 - s2)
    key = ((hi | lo) >> 2) & 0x3FFFFFFFFFFFFFFF
    return torch.where(inside, key, -(pos - n + 1))


def _refine_ties(P: torch.Tensor, pos: torch.Tensor, keys: torch.Tensor, n: int) -> torch.Tensor:
    """``pos`` is sorted by ``keys``; order the runs of equal keys by the symbols that follow
    (31 more per round).  ``eq[i]`` says elements i and i+1 are still tied."""
    eq = keys[1:] == keys[:-1]
    depth = 1
    while bool(eq.any()):
        tied = torch.zeros(pos.numel(), dtype=torch.bool, device=pos.device)
        tied[1:] |= eq
        tied[:-1] |= eq
        idx = torch.nonzero(tied).view(-1)
        first = torch.ones(idx.numel(), dtype=torch.bool, device=pos.device)
        nz = idx > 0
        first[nz] = ~eq[idx[nz] - 1]
        gid = torch.cumsum(first.to(torch.int64), 0)
        sub = pos[idx]
        k2 = _key31(P, sub + KEY_SYMS * depth, n)
        o1 = torch.sort(k2, stable=True).indices
        o2 = torch.sort(gid[o1], stable=True).indices
        order = o1[o2]
        pos[idx] = sub[order]
        g, k = gid[order], k2[order]
        same = (g[1:] == g[:-1]) & (k[1:] == k[:-1])
        eq = torch.zeros_like(eq)
        eq[idx[:-1][same]] = True
        depth += 1
        if depth > n // KEY_SYMS + 3:
            raise RuntimeError("suffix refinement did not converge")
    return pos


def bwt_string(T: torch.Tensor, bucket_syms: int | None = None, sa_intv: int = 0):
    """BWT of T$ with the ``$`` removed (uint8 symbols) and ``primary`` (is_bwt / bwt_bwtgen semantics).
    With ``sa_intv`` also returns the suffix-array samples of bwt_cal_sa (bwt.c:79-101): SA of every
    sa_intv-th row of the sorted rotations of T$ (row 0 = "$"), sa[0] = -1."""
    n = int(T.numel())
    dev = T.device
    if bucket_syms is None:
        bucket_syms = 0 if n < (1 << 22) else (2 if n < (1 << 30) else 3)
    P = _pack64(T, n)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    out[0] = T[n - 1]            # row 0 is the empty suffix "$"
    primary = -1
    row = 1                      # next output index to fill (= SA row until the '$' row is skipped)
    sa_row = 1                   # SA row of the next sorted suffix
    sa = None
    if sa_intv:
        sa = torch.zeros((n + sa_intv) // sa_intv, dtype=torch.int64, device=dev)
        sa[0] = -1
    nb = 4 ** bucket_syms
    if bucket_syms:
        code = torch.zeros(n, dtype=torch.uint8, device=dev)
        for j in range(bucket_syms):
            code[: n - j] |= T[j:] << (2 * (bucket_syms - 1 - j))
    for b in range(nb):
        if bucket_syms:
            parts = []
            for c0 in range(0, n, _NONZERO_CHUNK):
                nz = torch.nonzero(code[c0:c0 + _NONZERO_CHUNK] == b).view(-1)
                parts.append(nz + c0)
            pos = torch.cat(parts) if len(parts) > 1 else parts[0]
            del parts
        else:
            pos = torch.arange(n, dtype=torch.int64, device=dev)
        if pos.numel() == 0:
            continue
        keys = _key31(P, pos, n)
        keys, order = torch.sort(keys)
        pos = pos[order]
        del order
        pos = _refine_ties(P, pos, keys, n)
        del keys
        m = int(pos.numel())
        if sa_intv:
            first = (-sa_row) % sa_intv
            if first < m:
                sel = torch.arange(first, m, sa_intv, device=dev, dtype=torch.int64)
                sa[(sa_row + sel) // sa_intv] = pos[sel]
        sa_row += m
        zero = torch.nonzero(pos == 0).view(-1)
        ch = T[torch.clamp(pos - 1, min=0)]
        if zero.numel():
            z = int(zero[0])
            primary = row + z
            out[row:row + z] = ch[:z]
            out[row + z:row + m - 1] = ch[z + 1:]
            row += m - 1
        else:
            out[row:row + m] = ch
            row += m
        del pos, ch
    assert primary >= 0 and row == n
    if sa_intv:
        return out, primary, sa
    return out, primary


def pack_bwt(bstr: torch.Tensor, primary: int) -> BwtIndex:
    """Interleave occ checkpoints every 128 symbols (bwt_bwtupdate_core, bwtindex.c:128-150)."""
    n = int(bstr.numel())
    dev = bstr.device
    nblk = (n + 127) // 128
    blocks = torch.zeros((nblk, 16), dtype=torch.int32, device=dev)
    totals = torch.zeros(4, dtype=torch.int64, device=dev)
    step = 1 << 21  # blocks per pass
    for b0 in range(0, nblk, step):
        b1 = min(nblk, b0 + step)
        seg = torch.zeros((b1 - b0) * 128, dtype=torch.uint8, device=dev)
        s0, s1 = b0 * 128, min(n, b1 * 128)
        seg[: s1 - s0] = bstr[s0:s1]
        valid = None
        if s1 - s0 < seg.numel():
            valid = torch.zeros(seg.numel(), dtype=torch.bool, device=dev)
            valid[: s1 - s0] = True
        cnt = torch.empty((b1 - b0, 4), dtype=torch.int64, device=dev)
        for c in range(4):
            hit = seg == c
            if valid is not None:
                hit &= valid
            per = hit.view(-1, 128).sum(1, dtype=torch.int64)
            cnt[:, c] = torch.cumsum(per, 0) - per + totals[c]
            totals[c] += per.sum()
        blocks[b0:b1, :8] = cnt.view(torch.int32).view(-1, 8)
        q = seg.view(-1, 4)
        by = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        blocks[b0:b1, 8:] = by.view(-1, 4).flip(1).contiguous().view(torch.int32).view(-1, 8)
        del seg, q, by, cnt
    nwords = (n + 15) >> 4
    bwt_size = nwords + (nblk + 1) * 8
    flat = blocks.view(-1)
    used = (nblk - 1) * 16 + 8 + (nwords - (nblk - 1) * 8) if nblk else 0
    words = torch.empty(bwt_size, dtype=torch.int32, device=dev)
    words[:used] = flat[:used]
    words[used:] = totals.view(torch.int32)
    l2 = np.zeros(5, dtype=np.uint64)
    l2[1:] = np.cumsum(totals.cpu().numpy().astype(np.uint64))
    assert int(l2[4]) == n
    return BwtIndex(int(primary), l2, n, int(bwt_size), words)


def build_index(fwd: "torch.Tensor | np.ndarray", device: "str | torch.device | None" = None, sa_intv: int = 0) -> BwtIndex:
    """Forward reference symbols (0..3, all contigs concatenated) -> packed FM-index (+ SA samples if sa_intv)."""
    if isinstance(fwd, np.ndarray):
        fwd = torch.from_numpy(np.ascontiguousarray(fwd, dtype=np.uint8))
    if device is not None:
        fwd = fwd.to(device)
    T = text_from_forward(fwd)
    if sa_intv:
        bstr, primary, sa = bwt_string(T, sa_intv=sa_intv)
    else:
        (bstr, primary), sa = bwt_string(T), None
    del T
    ix = pack_bwt(bstr, primary)
    ix.sa_intv, ix.sa = sa_intv, sa
    return ix
