"""The C restatement (oracle/smem_oracle.c) against vectors dumped from the reference's own code."""
import numpy as np
import pytest

from conftest import OPT_KEYS
from oracle.binding import Oracle, SeedOpt


def _opt(z, key):
    o = z[f"opt_{key}"]
    return SeedOpt(int(o[0]), float(o[1]), int(o[2]), int(o[3]))


@pytest.mark.parametrize("key", OPT_KEYS)
def test_collect_matches_reference_dump(golden, key):
    name, z, ix = golden
    r = Oracle(ix).collect(z["seq"], z["offs"], _opt(z, key), nthreads=2)
    for k in ("intv", "read_off", "step", "n_steps", "last_start"):
        assert np.array_equal(r[k], z[f"collect_{key}_{k}"]), (name, key, k)


@pytest.mark.parametrize("rep", [0, 1, 2])
def test_smem1_matches_reference_dump(golden, rep):
    name, z, ix = golden
    r = Oracle(ix).smem1(z["smem1_seq"], z["smem1_offs"], z[f"smem1_{rep}_x"], z[f"smem1_{rep}_min_intv"])
    for k in ("intv", "read_off", "ret"):
        assert np.array_equal(r[k], z[f"smem1_{rep}_{k}"]), (name, rep, k)


def test_appendix_d_known_answers():
    """SURVEY.md Appendix D, written out literally (produced with the reference at survey time)."""
    import os
    from conftest import GOLDEN, GoldenIndex
    z = np.load(os.path.join(GOLDEN, "kat154.npz"))
    ix = GoldenIndex(z)
    assert (ix.primary, ix.seq_len, ix.bwt_size, list(ix.L2)) == (35, 308, 52, [0, 81, 154, 227, 308])
    r = Oracle(ix).collect(z["seq"], z["offs"], SeedOpt(19, 1.5, 10, 1))
    got = [(int(a), int(b), int(c), int(i >> 32), int(i & 0xFFFFFFFF), int(s)) for (a, b, c, i), s in zip(r["intv"], r["step"])]
    assert got == [(47, 144, 1, 0, 41, 0), (76, 163, 2, 8, 35, 0),
                   (241, 42, 1, 0, 20, 0), (136, 195, 1, 21, 41, 1),
                   (173, 37, 1, 0, 27, 0), (222, 1, 6, 25, 28, 1), (304, 59, 1, 27, 32, 1), (166, 207, 1, 29, 51, 2)]
    assert list(r["read_off"]) == [0, 2, 4, 8]
    assert list(r["last_start"]) == [41, 41, 51]
    s = Oracle(ix).smem1(z["seq"], z["offs"], [20, 20, 25], [1, 1, 1])
    assert list(s["ret"]) == [41, 21, 28] and list(s["read_off"]) == [0, 1, 1, 3]
    assert s["intv"][:, :3].tolist() == [[47, 144, 1], [173, 37, 1], [222, 1, 6]]


def test_occ4_and_extend_brute_force(golden):
    """occ4 against a direct count over the unpacked BWT string; extend against the occ definition."""
    name, z, ix = golden
    o = Oracle(ix)
    w = ix.bwt
    n = ix.seq_len
    syms = []
    nblk = (n + 127) // 128
    for b in range(nblk):
        words = w[16 * b + 8: 16 * b + 16] if b < nblk - 1 else w[16 * b + 8: 16 * b + 8 + ((n - 128 * b) + 15) // 16]
        for v in words:
            syms += [(int(v) >> (30 - 2 * j)) & 3 for j in range(16)]
    syms = np.array(syms[:n])
    cum = np.zeros((n + 1, 4), np.int64)
    for c in range(4):
        cum[1:, c] = np.cumsum(syms == c)
    rng = np.random.default_rng(3)
    ks = list(rng.integers(0, n + 1, 200)) + [0, n, ix.primary, max(ix.primary - 1, 0), min(ix.primary + 1, n), 127, 128, 129]
    for k in ks:
        kk = int(k) - (int(k) >= ix.primary)
        assert o.occ4(int(k)) == list(cum[kk + 1]), k
    assert o.occ4(-1) == [0, 0, 0, 0]
