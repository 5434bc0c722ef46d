"""GPU: the k-mer count pyramid fast path of smem_gpu_collect (csrc/smem_fast.cuh) through the C ABI.

* the tables the library builds on the device equal their numpy definition (bwa-mem-harp2_b200/kmer_tables.py);
* smem_gpu_collect with the tables is bit-exact with the oracle on seeded sets (repeats, ambiguous bases, ragged and
  empty reads, re-seeding, long reads), with and without them it returns the same bytes, and reads the tables cannot
  answer for are seeded by the FM kernel.
"""
import numpy as np
import pytest

from conftest import pkg, same_result
from oracle.binding import Oracle, SeedOpt as OSeedOpt

pytestmark = pytest.mark.gpu


def gopt(o):
    sg = pkg("smem_gpu")
    return sg.SeedOpt(int(o[0]), float(o[1]), int(o[2]), int(o[3]))


@pytest.fixture(scope="module")
def sg():
    m = pkg("smem_gpu")
    m.load_library()
    return m


@pytest.fixture(scope="module")
def world(fm, synth, sg):
    ref = synth.make_reference(1_000_000, 78)
    ref[50_000:52_000] = ref[10_000:12_000]          # a 2 kbp repeat
    ref[400_000:400_400] = 1                         # a homopolymer run: saturated counts -> FM re-run
    ix = fm.build_index(ref)
    g = sg.SmemGpu(max_batch_reads=40_000, max_read_len=260)
    g.upload_index(ix)
    g.build_kmer_tables(ref, 6)                      # tables up to 11-mers on a 2 Mbp text (4^11 = 4.2 M)
    g.set_param("fast", 1)                           # the table-driven path is opt-in
    return ref, ix, Oracle(ix), g


@pytest.mark.parametrize("DL", [3, 6])
def test_device_tables_equal_definition(fm, synth, sg, DL):
    kt = pkg("kmer_tables")
    ref = synth.make_reference(150_001, 5)           # odd length: the .pac tail byte is partial
    T = fm.text_from_forward(ref).numpy()
    tb = kt.build_kmer_tables(T, DL, True)
    g = sg.SmemGpu(max_batch_reads=64, max_read_len=128)
    g.upload_index(fm.build_index(ref))
    g.build_kmer_tables(ref, DL)
    assert g.get_param("has_kmer_tables") == 1
    for L in range(1, DL + 1):
        assert np.array_equal(g.kmer_table(0, L), tb.cnt[L]), ("cnt", L)
    for L in range(1, DL + 2):
        assert np.array_equal(g.kmer_table(1, L), tb.cum[L]), ("cum", L)
    assert np.array_equal(g.kmer_table(2), tb.pyr)
    assert np.array_equal(g.kmer_table(3), tb.top)
    g.close()


@pytest.mark.parametrize("n,L,err,nfrac,opt", [
    (20000, 101, 0.01, 0.06, (19, 1.5, 10, 1)),
    (6000, 250, 0.02, 0.06, (19, 1.5, 10, 1)),
    (5000, 101, 0.01, 0.0, (19, 1.5, 10, 2)),
    (5000, 60, 0.05, 0.3, (10, 1.2, 20, 1)),
    (3000, 150, 0.0, 0.0, (19, 1.5, 0, 1)),
    (3000, 36, 0.1, 0.0, (19, 1.5, 10, 1)),
    (3000, 9, 0.0, 0.1, (19, 1.5, 10, 1)),           # reads shorter than the deepest table level
])
def test_fast_collect_vs_oracle(world, synth, n, L, err, nfrac, opt):
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, L, err, seed=n + L, n_frac=nfrac))
    a = o.collect(seq, offs, OSeedOpt(*opt), nthreads=8)
    g.set_param("fast", 1)
    b = g.collect(seq, offs, gopt(opt))
    esc = g.get_param("escaped_reads")
    same_result(a, b, ("read_off", "intv", "step"))
    assert esc < n // 2, "the tables should answer for most reads"
    g.set_param("fast", 0)
    c = g.collect(seq, offs, gopt(opt))
    g.set_param("fast", 1)
    same_result(b, c, ("read_off", "intv", "step"))


def test_fast_repeats_and_ragged(world, synth):
    """Reads from the repeat and the homopolymer (saturated table entries -> FM re-run), empty / all-N / 1-base reads."""
    ref, ix, o, g = world
    rng = np.random.default_rng(3)
    reads = []
    for _ in range(400):
        p = int(rng.integers(10_000, 11_800)); reads.append(ref[p:p + 120].numpy().copy())
    for _ in range(400):
        p = int(rng.integers(399_900, 400_300)); reads.append(ref[p:p + 100].numpy().copy())
    reads += [np.zeros(0, np.uint8), np.full(30, 4, np.uint8), np.array([2], np.uint8), np.array([4, 1, 4], np.uint8)]
    for _ in range(300):
        ln = int(rng.integers(1, 200)); p = int(rng.integers(0, 900_000)); r = ref[p:p + ln].numpy().copy()
        if ln > 4 and rng.random() < 0.5:
            r[rng.integers(0, ln)] = 4
        reads.append(r)
    seq, offs = synth.to_batch(reads)
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    b = g.collect(seq, offs)
    same_result(a, b, ("read_off", "intv", "step"))
    assert g.get_param("escaped_reads") > 0          # the homopolymer reads went through the FM kernel


def test_fast_small_slots_and_knobs(world, synth):
    """Overflowing result slots, tiny real-entry capacity and other launch geometries do not change the result."""
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 4000, 150, 0.03, seed=12, n_frac=0.02))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    old = g.get_param("slot_cap")
    for name, val, back in (("slot_cap", 8, old), ("fast_b_cap", 1, 4), ("fast_blocks_per_sm", 2, 4), ("fast_slots", 40, 128)):
        g.set_param(name, val)
        b = g.collect(seq, offs)
        g.set_param(name, back)
        same_result(a, b, ("read_off", "intv", "step"))
