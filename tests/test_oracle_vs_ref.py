"""Pins the restatement against the reference's own compiled objects (build container only)."""
import numpy as np
import pytest

from oracle.binding import Oracle, Reference, SeedOpt, build_reference

ref_so = build_reference()
pytestmark = pytest.mark.skipif(ref_so is None, reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def setup(fm, synth):
    ref = synth.make_reference(300000, 21)
    ix = fm.build_index(ref)
    return ref, ix, Oracle(ix), Reference(ix)


@pytest.mark.parametrize("n,L,err,nfrac,opt", [
    (1500, 101, 0.01, 0.06, (19, 1.5, 10, 1)),
    (600, 250, 0.02, 0.06, (19, 1.5, 10, 1)),
    (500, 101, 0.01, 0.0, (19, 1.5, 10, 2)),
    (500, 60, 0.05, 0.3, (10, 1.2, 20, 1)),
    (300, 150, 0.0, 0.0, (19, 1.5, 0, 1)),
])
def test_collect_equal(setup, synth, n, L, err, nfrac, opt):
    ref, ix, o, r = setup
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, L, err, seed=n + L, n_frac=nfrac))
    a = o.collect(seq, offs, SeedOpt(*opt), nthreads=3)
    b = r.collect(seq, offs, SeedOpt(*opt), nthreads=2)
    for k in ("intv", "read_off", "step", "n_steps", "last_start"):
        assert np.array_equal(a[k], b[k]), k
    ta, tb = o.time_collect(seq, offs, SeedOpt(*opt), 2), r.time_collect(seq, offs, SeedOpt(*opt), 2)
    assert ta["checksum"] == tb["checksum"] == o.checksum(a["intv"], a["read_off"])
    assert ta["n_intervals"] == tb["n_intervals"] == len(a["intv"])


def test_smem1_occ_extend_equal(setup, synth):
    ref, ix, o, r = setup
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 800, 101, 0.02, seed=9, n_frac=0.1))
    rng = np.random.default_rng(4)
    x = rng.integers(0, 101, 800).astype(np.int32)
    mi = rng.integers(0, 5, 800).astype(np.int32)
    a, b = o.smem1(seq, offs, x, mi), r.smem1(seq, offs, x, mi)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    for k in list(rng.integers(0, ix.seq_len + 1, 300)) + [0, ix.seq_len, ix.primary, ix.primary - 1]:
        assert o.occ4(int(k)) == r.occ4(int(k))
    for _ in range(200):
        c = int(rng.integers(0, 4))
        ik = [int(ix.L2[c]) + 1, int(ix.L2[3 - c]) + 1, int(ix.L2[c + 1] - ix.L2[c])]
        for back in (0, 1):
            assert np.array_equal(o.extend(ik, back), r.extend(ik, back))
