"""Every variant of the seed kernel -- lane pairs on the 64-byte blocks (lanes_per_read 2), lane pairs on the 32-byte sector form
with the backward sweep split over the two lanes (4, the default) or not (3), one lane per read on the sector form (1; smem_device.cuh
extend_single) -- must give the oracle's lists bit for bit: collect, raw bwt_smem1 calls, the trace, every shortcut, spills."""
import numpy as np
import pytest

from conftest import pkg, same_result
from oracle.binding import Oracle, SeedOpt as OSeedOpt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 2, 3, 4])
def world(request, fm, synth):
    sg = pkg("smem_gpu")
    ref = synth.make_reference(800_000, 17)
    ref[50_000:52_000] = ref[10_000:12_000]
    for k in range(12):                                        # a repeat family: long candidate lists
        ref[300_000 + 5_000 * k:300_000 + 5_000 * k + 400] = ref[20_000:20_400]
    ix = fm.build_index(ref, sa_intv=32)
    g = sg.SmemGpu(max_batch_reads=30_000, max_read_len=260)
    g.set_param("lanes_per_read", request.param)               # before the upload: it builds the sector form
    g.upload_index(ix); g.upload_sa(ix)
    g.build_repeat_filter(ref)
    g.build_text_index(ref)
    yield sg, ref, ix, Oracle(ix), g
    g.close()


@pytest.mark.parametrize("n,L,err,nfrac,opt", [
    (12000, 101, 0.01, 0.05, (19, 1.5, 10, 1)),
    (4000, 250, 0.02, 0.05, (19, 1.5, 10, 1)),
    (4000, 101, 0.03, 0.0, (19, 1.5, 10, 2)),
    (3000, 40, 0.08, 0.2, (10, 1.2, 20, 1)),
])
def test_collect(world, synth, n, L, err, nfrac, opt):
    sg, ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, L, err, seed=n + L, n_frac=nfrac))
    a = o.collect(seq, offs, OSeedOpt(*opt), nthreads=8)
    b = g.collect(seq, offs, sg.SeedOpt(int(opt[0]), float(opt[1]), int(opt[2]), int(opt[3])))
    same_result(a, b, ("read_off", "intv", "step"))


def test_ragged_reads_spills_and_knobs(world, synth):
    sg, ref, ix, o, g = world
    rng = np.random.default_rng(4)
    refn = ref.numpy()
    reads = [np.zeros(0, np.uint8), np.array([4], np.uint8), np.array([2], np.uint8), np.full(70, 4, np.uint8), refn[20_000:20_260].copy()]
    for _ in range(1500):
        L = int(rng.integers(1, 261))
        p = int(rng.integers(0, len(refn) - L))
        r = refn[p:p + L].copy()
        if rng.random() < 0.3:
            r[rng.integers(0, L, 2)] = 4
        if rng.random() < 0.5 and r.max(initial=0) < 4:
            r = (3 - r)[::-1].copy()
        if L > 30 and rng.random() < 0.7:
            r[int(rng.integers(0, L))] ^= 1
        reads.append(r)
    seq, offs = synth.to_batch(reads)
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    try:
        for name, val in [("b_cap", 17), ("b_cap", 3), ("slot_cap", 2), ("slot_cap", 224), ("unique_walk", 0), ("unique_walk", 1), ("spec_walk", 0),
                          ("repeat_filter", 0), ("unique_walk_min_run", 3), ("unique_walk_min_left", 1)]:
            g.set_param(name, val)
            same_result(a, g.collect(seq, offs), ("read_off", "intv", "step"))
    finally:
        g.set_param("b_cap", 17); g.set_param("slot_cap", 224); g.set_param("spec_walk", 1); g.set_param("repeat_filter", 1)
        g.set_param("unique_walk_min_left", 8)
    pr = sg.PackedReads(g.lib, seq, offs)
    got = g.collect_packed12(pr)
    assert np.array_equal(got["intv"], a["intv"]) and np.array_equal(got["read_off"], a["read_off"])


def test_smem1_and_trace(world, synth):
    sg, ref, ix, o, g = world
    n = 6000
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, 101, 0.02, seed=9, n_frac=0.08))
    rng = np.random.default_rng(3)
    x = rng.integers(0, 101, n).astype(np.int32)
    mi = rng.integers(0, 5, n).astype(np.int32)
    same_result(o.smem1(seq, offs, x, mi), g.smem1(seq, offs, x, mi), ("read_off", "intv", "ret"))
    g2 = sg.SmemGpu(max_batch_reads=30_000, max_read_len=260)      # the lane-pair kernel on the 64-byte blocks
    g2.set_param("lanes_per_read", 2)
    g2.upload_index(ix)
    ta, tb = g2.trace(seq, offs), g.trace(seq, offs)
    for k in ("read_off", "intv", "tag", "ret"):
        assert np.array_equal(ta[k], tb[k]), k
    g2.close()
