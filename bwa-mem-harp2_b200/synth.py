"""Seeded synthetic references and simulated reads.

Reference: i.i.d. uniform A/C/G/T, symbols 0..3, contigs concatenated (seeding never looks at
contig boundaries).  Reads: uniform start positions on the forward strand, i.i.d. substitution
errors at rate ``err`` (always to one of the three other bases), half of them reverse-complemented,
optionally one ambiguous base (code 4) in a fraction of the reads for edge coverage.
Everything is generated with ``torch`` on the device of the reference tensor, so the 3.1 Gbp
workload is produced on the GPU in seconds; small parity sets use the CPU generator and are
therefore identical in the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch


def make_reference(n_bp: int, seed: int, device: "str | torch.device" = "cpu") -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n_bp, dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, n_bp, step):
        e = min(n_bp, s + step)
        out[s:e] = torch.randint(0, 4, (e - s,), generator=g, device=device, dtype=torch.uint8)
    return out


def add_repeat_families(fwd: torch.Tensor, frac: float, seed: int, n_families: int = 4, unit_len: int = 300,
                        divergence: float = 0.05) -> torch.Tensor:
    """Overwrite about ``frac`` of ``fwd`` in place with diverged copies (either strand) of ``n_families`` consensus
    sequences of ``unit_len`` bases -- a crude model of interspersed repeats (Alu-like) for robustness runs: seeds with
    thousands of occurrences, many chains per read, a well filled repeat filter."""
    dev = fwd.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = int(fwd.numel())
    cons = torch.randint(0, 4, (n_families, unit_len), generator=g, device=dev, dtype=torch.uint8)
    n_ins = int(frac * L / unit_len)
    ar = torch.arange(unit_len, device=dev, dtype=torch.int64)
    step = max(1, (1 << 24) // unit_len)
    for s0 in range(0, n_ins, step):
        m = min(step, n_ins - s0)
        pos = (torch.rand(m, generator=g, device=dev, dtype=torch.float64) * (L - unit_len)).to(torch.int64)
        fam = torch.randint(0, n_families, (m,), generator=g, device=dev)
        unit = cons[fam]
        hit = torch.rand((m, unit_len), generator=g, device=dev) < divergence
        delta = torch.randint(1, 4, (m, unit_len), generator=g, device=dev, dtype=torch.uint8)
        unit = torch.where(hit, (unit + delta) & 3, unit)
        rc = torch.rand(m, generator=g, device=dev) < 0.5
        unit = torch.where(rc[:, None], 3 - unit.flip(1), unit)
        fwd[(pos[:, None] + ar[None, :]).view(-1)] = unit.reshape(-1)
    return fwd


def simulate_reads(fwd: torch.Tensor, n_reads: int, read_len: int, err: float, seed: int,
                   n_frac: float = 0.0, paired: bool = False, insert_mean: float = 400.0,
                   insert_sd: float = 50.0) -> torch.Tensor:
    """Returns uint8 [n_reads, read_len] (codes 0..3, 4 = N).  With ``paired`` reads 2i and 2i+1
    are mates (FR orientation, insert ~ N(mean, sd)), as in BASELINE config 3."""
    dev = fwd.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = int(fwd.numel())
    if paired:
        n_pairs = (n_reads + 1) // 2
        ins = (torch.randn(n_pairs, generator=g, device=dev) * insert_sd + insert_mean).round().to(torch.int64)
        ins = ins.clamp(min=read_len, max=max(read_len, L - 1))
        p1 = (torch.rand(n_pairs, generator=g, device=dev, dtype=torch.float64) * (L - ins).clamp(min=1).to(torch.float64)).to(torch.int64)
        p2 = p1 + ins - read_len
        flip = torch.rand(n_pairs, generator=g, device=dev) < 0.5   # which mate is forward
        pos = torch.stack([p1, p2], 1).view(-1)[:n_reads]
        rc = torch.stack([flip, ~flip], 1).view(-1)[:n_reads]
    else:
        pos = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * (L - read_len + 1)).to(torch.int64)
        rc = torch.rand(n_reads, generator=g, device=dev) < 0.5
    pos = pos.clamp(min=0, max=L - read_len)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    reads = torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    step = max(1, (1 << 26) // read_len)
    for s in range(0, n_reads, step):
        e = min(n_reads, s + step)
        r = fwd[pos[s:e, None] + ar[None, :]]
        if err > 0:
            hit = torch.rand((e - s, read_len), generator=g, device=dev) < err
            delta = torch.randint(1, 4, (e - s, read_len), generator=g, device=dev, dtype=torch.uint8)
            r = torch.where(hit, (r + delta) & 3, r)
        rcm = rc[s:e, None]
        r = torch.where(rcm, (3 - r).flip(1), r)
        if n_frac > 0:
            has_n = torch.rand(e - s, generator=g, device=dev) < n_frac
            npos = torch.randint(0, read_len, (e - s,), generator=g, device=dev)
            mask = has_n[:, None] & (ar[None, :] == npos[:, None])
            r = torch.where(mask, torch.full_like(r, 4), r)
        reads[s:e] = r
    return reads


def to_batch(reads: "torch.Tensor | np.ndarray | list") -> tuple[np.ndarray, np.ndarray]:
    """Flat host batch (seq uint8, offs int64[n+1]) from a [n, len] matrix or a list of arrays."""
    if isinstance(reads, torch.Tensor):
        reads = reads.cpu().numpy()
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, l = reads.shape
        return np.ascontiguousarray(reads, dtype=np.uint8).reshape(-1), np.arange(n + 1, dtype=np.int64) * l
    lens = np.array([len(r) for r in reads], dtype=np.int64)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    seq = np.concatenate([np.asarray(r, dtype=np.uint8) for r in reads]) if len(reads) else np.zeros(0, np.uint8)
    return seq, offs


_NT = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _NT[ord(_c)] = _i
    _NT[ord(_c.lower())] = _i


def encode(s: str) -> np.ndarray:
    """ASCII -> 0..4 as nst_nt4_table does (bntseq.c:44)."""
    return _NT[np.frombuffer(s.encode(), dtype=np.uint8)]


def write_fasta(path: str, contigs: "list[tuple[str, np.ndarray]]") -> None:
    with open(path, "w") as f:
        for name, sym in contigs:
            f.write(f">{name}\n")
            txt = np.frombuffer(b"ACGT", dtype=np.uint8)[sym].tobytes().decode()
            for i in range(0, len(txt), 80):
                f.write(txt[i:i + 80] + "\n")


def write_fastq(path: str, reads: np.ndarray, prefix: str = "r") -> None:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(f"@{prefix}{i}\n{lut[np.minimum(r, 4)].tobytes().decode()}\n+\n{'I' * len(r)}\n")
