"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Every expected array below is produced by the reference's own code:
  * bwt words / primary / L2  <- `bwa_ref index`  (bwtindex.c:239-280)
  * collect_*                 <- smem_next2 loop  (bwamem.c:244-305 via oracle/ref_harness.c)
  * smem1_*                   <- bwt_smem1        (bwt.c:776-835)
"""
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
sy = importlib.import_module("bwa-mem-harp2_b200.synth")
from oracle.binding import Reference, SeedOpt, build_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
BWA = os.path.join(ROOT, "oracle", "_ref", "bwa_ref")

KAT_REF = ("ACGTTGCATGGACCTAGGATCCGATTACAGGCTTAACGGTCATGCAAGTCCGATAGCTTGACCATGGTTAACCGGTATCGGATCAAGTCTAGGCTAACGTAGCCTGATCG"
           "GAATTCCGTATTCAATTACAGGCTTAACGGTCATGCAAGTCTTG")
KAT_READS = ["AGGATCCGATTACAGGCTTAACGGTCATGCAAGTCCGATAG", "TACGGAATTCCGATCAGGCTNCGTTAGCCTAGACTTGATCC",
             "GATTACAGGCTTAACGGTCATGCAAGTTTGATAGCTTGACCATGGTTAACC"]

OPTS = {"default": (19, 1.5, 10, 1), "noexact": (19, 1.5, 10, 2), "short": (10, 1.2, 20, 1), "nosplit": (19, 0.0, 10, 1)}


def ref_index(contigs, workdir):
    fa = os.path.join(workdir, "g.fa")
    sy.write_fasta(fa, contigs)
    subprocess.run([BWA, "index", fa], check=True, capture_output=True, cwd=workdir)
    return fm.BwtIndex.load(fa + ".bwt")


def dump(name, fwd, contigs, reads, workdir, rng):
    ix = ref_index(contigs, workdir)
    ref = Reference(ix)
    seq, offs = sy.to_batch(reads)
    out = dict(fwd=fwd.astype(np.uint8), primary=np.uint64(ix.primary), L2=ix.L2, seq_len=np.uint64(ix.seq_len),
               bwt=ix.words_numpy(), seq=seq, offs=offs)
    for key, o in OPTS.items():
        r = ref.collect(seq, offs, SeedOpt(*o), nthreads=1)
        for k in ("intv", "read_off", "step", "n_steps", "last_start"):
            out[f"collect_{key}_{k}"] = r[k]
        out[f"opt_{key}"] = np.array(o, dtype=np.float64)
    # bwt_smem1 dereferences q[x] and curr->a[0] unconditionally (bwt.c:783,808): zero-length reads are
    # outside its domain, so the raw-call fixtures use the non-empty reads only.
    keep = [r for r in reads if len(r) > 0]
    seq1, offs1 = sy.to_batch(keep)
    out["smem1_seq"], out["smem1_offs"] = seq1, offs1
    lens = np.diff(offs1)
    for rep in range(3):
        x = np.array([rng.integers(0, max(1, l)) for l in lens], dtype=np.int32)
        mi = rng.integers(0, 6, len(lens)).astype(np.int32)
        r = ref.smem1(seq1, offs1, x, mi)
        out[f"smem1_{rep}_x"] = x
        out[f"smem1_{rep}_min_intv"] = mi
        for k in ("intv", "read_off", "ret"):
            out[f"smem1_{rep}_{k}"] = r[k]
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "reads", len(offs) - 1, "intervals(default)", len(out["collect_default_intv"]), os.path.getsize(path), "bytes")


def main():
    assert build_reference(), "reference not available"
    rng = np.random.default_rng(20260101)
    with tempfile.TemporaryDirectory() as wd:
        # 1. SURVEY.md Appendix D known-answer test
        fwd = sy.encode(KAT_REF)
        dump("kat154", fwd, [("k2", fwd)], [sy.encode(r) for r in KAT_READS], wd, rng)
        # 2. small random genome with planted repeats, ragged reads, ambiguous bases
        n = 24000
        fwd = rng.integers(0, 4, n).astype(np.uint8)
        fwd[3000:3600] = fwd[1000:1600]            # exact repeat
        fwd[9000:9400] = 3 - fwd[5000:5400][::-1]  # reverse-complement repeat
        fwd[15000:15200] = 0                       # poly-A
        fwd[15200:15400] = np.tile([0, 1], 100)    # dinucleotide repeat
        reads = []
        for i in range(260):
            L = int(rng.integers(20, 151))
            p = int(rng.integers(0, n - L))
            r = fwd[p:p + L].copy()
            err = rng.random(L) < rng.choice([0.0, 0.01, 0.05])
            r[err] = (r[err] + rng.integers(1, 4, int(err.sum()))) & 3
            if rng.random() < 0.5:
                r = (3 - r)[::-1]
            if rng.random() < 0.25:
                for _ in range(int(rng.integers(1, 4))):
                    r[int(rng.integers(0, L))] = 4
            reads.append(r)
        reads += [np.array([4, 4, 4, 4], np.uint8), np.array([2], np.uint8), np.array([4], np.uint8),
                  np.zeros(0, np.uint8), fwd[15000:15120].copy(), fwd[15150:15290].copy(),
                  np.concatenate([[4], fwd[100:140], [4]]).astype(np.uint8), fwd[1000:1101].copy(),
                  rng.integers(0, 4, 101).astype(np.uint8)]
        dump("small24k", fwd, [("c1", fwd[:10000]), ("c2", fwd[10000:])], reads, wd, rng)


if __name__ == "__main__":
    main()
