"""Compact wire formats (second half): two-bit reads in, 16-byte interval records out.
The interval lists must be the ones smem_gpu_collect / the oracle produce, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from conftest import pkg
from oracle.binding import Oracle, SeedOpt as OSeedOpt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(fm, synth):
    sg = pkg("smem_gpu")
    ref = synth.make_reference(600_000, 91)
    ref[30_000:31_500] = ref[200_000:201_500]
    ix = fm.build_index(ref)
    g = sg.SmemGpu(max_batch_reads=20_000, max_read_len=260)
    g.upload_index(ix)
    g.build_repeat_filter(ref)
    return sg, ref, ix, Oracle(ix), g


def check(sg, g, o, seq, offs, opt=None, **kw):
    want = o.collect(seq, offs, opt or OSeedOpt(), nthreads=4)
    pr = sg.PackedReads(g.lib, seq, offs, **kw)
    gopt = sg.SeedOpt(*(getattr(opt, f) for f in ("min_seed_len", "split_factor", "split_width", "start_width"))) if opt else None
    got = g.collect_packed(pr, gopt)
    assert np.array_equal(got["read_off"], want["read_off"])
    assert np.array_equal(got["intv"], want["intv"])
    got = g.collect_packed12(pr, gopt)                                  # 12-byte records + exception list: the same lists again
    assert np.array_equal(got["read_off"], want["read_off"])
    assert np.array_equal(got["intv"], want["intv"])
    got = g.collect_packed11(pr, gopt)                                  # ... byte-packed 11-byte records (the handle's 260-base limit leaves 4 bits of size)
    assert np.array_equal(got["read_off"], want["read_off"])
    assert np.array_equal(got["intv"], want["intv"])
    got16 = g.collect_packed(pr, gopt)                                  # (kept last: the d2h byte counts asserted by callers are the 16-byte form's)
    assert np.array_equal(got16["intv"], want["intv"])
    return pr, want


def test_uniform_reads_with_ambiguous_bases(world, synth):
    sg, ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 6000, 101, 0.02, seed=5, n_frac=0.08, paired=True))
    pr, want = check(sg, g, o, seq, offs)
    assert pr.lens is None and pr.n_amb > 0 and pr.stride == 26
    assert g.timing()["h2d_bytes"] == 6000 * 26 + 8 * pr.n_amb          # 26 bytes per read instead of 101 + 8
    assert g.timing()["d2h_bytes"] == 6000 * 4 + 16 * len(want["intv"])   # 16 bytes per interval instead of 32, 4 per offset instead of 8


def test_ragged_reads_lens_and_edge_lengths(world, synth):
    sg, ref, ix, o, g = world
    rng = np.random.default_rng(3)
    refn = ref.numpy()
    reads = [np.zeros(0, np.uint8), np.full(5, 4, np.uint8), refn[100:101].copy()]
    for ln in (1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 255, 256, 257, 260):
        for _ in range(6):
            p = int(rng.integers(0, len(refn) - ln))
            q = refn[p:p + ln].copy()
            if rng.random() < 0.5:
                q = (3 - q)[::-1].copy()
            if ln > 8 and rng.random() < 0.6:
                q[rng.integers(0, ln, 2)] = 4                            # adjacent / repeated N positions share a packed byte
            if ln > 40:
                q[int(rng.integers(0, ln))] ^= 1
            reads.append(q)
    seq, offs = synth.to_batch(reads)
    pr, _ = check(sg, g, o, seq, offs)
    assert pr.lens is not None
    check(sg, g, o, seq, offs, stride=80)                               # a wider stride than needed
    check(sg, g, o, seq, offs, OSeedOpt(19, 1.5, 10, 2))                # NO_EXACT
    # empty batch
    e = sg.PackedReads(g.lib, np.zeros(0, np.uint8), np.zeros(1, np.int64))
    r = g.collect_packed(e)
    assert len(r["intv"]) == 0 and list(r["read_off"]) == [0]


def test_lanes_split_forms_and_conversions(world, synth):
    sg, ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 5000, 151, 0.03, seed=8, n_frac=0.1))
    want = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    pr = sg.PackedReads(g.lib, seq, offs)
    # compact reads in, 32-byte results out through the split form; then the same resident results as 16-byte records
    g.stage_packed(pr)
    tot = g.run_collect()
    a = g.fetch(tot, want_step=True)
    assert np.array_equal(a["intv"], want["intv"]) and np.array_equal(a["read_off"], want["read_off"]) and np.array_equal(a["step"], want["step"])
    b = g.fetch_packed(tot)
    assert np.array_equal(b["intv"], want["intv"]) and np.array_equal(b["read_off"], want["read_off"])
    # byte reads in, 16-byte records out
    g.stage(seq, offs); tot = g.run_collect()
    b = g.fetch_packed(tot)
    assert np.array_equal(b["intv"], want["intv"])
    # a handle with two pipeline lanes on the GPU: the shards' ambiguous-base ranges and offsets are stitched on the host
    g2 = sg.SmemGpu(max_batch_reads=8000, max_read_len=160, devices=[0, 0, 0])
    g2.share_index_from(g)
    got = g2.collect_packed(pr)
    assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"])
    # 32-byte fetch of 16-byte resident results is refused, not converted silently
    g2._n = pr.n
    with pytest.raises(sg.SmemGpuError):
        g2.fetch(len(want["intv"]))
    g2.close()


def test_packed12_exception_list(world, synth):
    """12-byte records (smem_intv12_t): sizes that do not fit the record's field travel in the exception list -- also when that list
    outgrows its first buffer, when reads outgrow their result slots (overflow re-run), through the split form and on a lane handle."""
    sg, ref, ix, o, g = world
    refn = ref.numpy()
    reads = []
    for k in range(400):                                                # single bases between Ns: every interval is an exception (x[2] ~ 300 k)
        q = np.full(260, 4, np.uint8)
        q[::2] = refn[1000 * k:1000 * k + 130]
        reads.append(q)
    reads += synth.simulate_reads(ref, 2000, 101, 0.02, seed=15, n_frac=0.05)
    for k in range(50):                                                 # short reads: few, large intervals
        reads.append(refn[5000 * k:5000 * k + 2 + k % 9].copy())
    seq, offs = synth.to_batch(reads)
    want = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    pr = sg.PackedReads(g.lib, seq, offs)
    got = g.collect_packed12(pr)
    assert got["pos_bits"] == 9 and len(got["exc"]) > 400 * 130        # (more than the list's first size: 1.5 entries per read of the handle)
    assert np.array_equal(got["read_off"], want["read_off"]) and np.array_equal(got["intv"], want["intv"])
    t = g.timing()
    assert t["d2h_bytes"] == len(offs) * 4 - 4 + 12 * len(want["intv"]) + 12 * len(got["exc"])
    got11 = g.collect_packed11(pr)
    assert np.array_equal(got11["read_off"], want["read_off"]) and np.array_equal(got11["intv"], want["intv"]) and len(got11["exc"]) >= len(got["exc"])
    assert g.timing()["d2h_bytes"] == len(offs) * 4 - 4 + 11 * len(want["intv"]) + 12 * len(got11["exc"])
    # reads that outgrow their slots: the first 64 entries are placed by the first compaction, the whole list again by the re-run
    g.set_param("slot_cap", 64)
    try:
        got = g.collect_packed12(pr)
        assert g.timing()["overflow_reads"] >= 400
        assert np.array_equal(got["read_off"], want["read_off"]) and np.array_equal(got["intv"], want["intv"])
        got = g.collect_packed11(pr)
        assert np.array_equal(got["read_off"], want["read_off"]) and np.array_equal(got["intv"], want["intv"])
    finally:
        g.set_param("slot_cap", 224)
    # split form: resident 32-byte results -> 12-byte records; and 16-byte resident results are not converted silently
    g.stage_packed(pr); tot = g.run_collect()
    b = g.fetch_packed12(tot)
    assert np.array_equal(b["intv"], want["intv"]) and np.array_equal(b["read_off"], want["read_off"])
    g.stage_packed(pr); tot = g.run_collect()
    b = g.fetch_packed11(tot)
    assert np.array_equal(b["intv"], want["intv"]) and np.array_equal(b["read_off"], want["read_off"])
    g.collect_packed(pr)
    g._n = pr.n
    with pytest.raises(sg.SmemGpuError):
        g.fetch_packed12(len(want["intv"]))
    # pipeline lanes: exception indices are rebased onto the whole batch
    g2 = sg.SmemGpu(max_batch_reads=8000, max_read_len=260, devices=[0, 0, 0])
    g2.share_index_from(g)
    got = g2.collect_packed12(pr)
    assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"])
    g2.close()
    # capacity protocol
    tot, ne, pb = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    roff = np.zeros(pr.n + 1, np.uint32)
    rec = np.empty((len(want["intv"]), 3), np.uint32)
    rc = g.lib.smem_gpu_collect_packed12(g.h, C.byref(pr.desc), C.byref(sg.SeedOpt()), C.c_void_p(rec.ctypes.data), C.c_int64(len(rec)),
                                         C.c_void_p(roff.ctypes.data), None, C.c_int64(0), C.byref(ne), C.byref(pb), C.byref(tot))
    assert rc == -5 and ne.value == len(got["exc"]) and tot.value == len(rec) and int(roff[-1]) == tot.value


def test_errors_are_reported_not_aborted(world, synth):
    sg, ref, ix, o, g = world
    seq, offs = synth.to_batch([ref.numpy()[:300].copy(), ref.numpy()[500:600].copy()])      # 300 > max_read_len 260
    with pytest.raises(sg.SmemGpuError) as e:
        g.collect(seq, offs)
    assert e.value.code == -5
    pr = sg.PackedReads(g.lib, seq, offs)
    with pytest.raises(sg.SmemGpuError) as e:
        g.collect_packed(pr)
    assert e.value.code == -5
    # the handle stays usable
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 500, 101, 0.01, seed=2))
    check(sg, g, o, seq, offs)
    # capacity protocol: read_off and the total come back, the caller re-sizes
    pr = sg.PackedReads(g.lib, seq, offs)
    tot = C.c_int64(0)
    roff = np.zeros(501, np.uint32)
    small = np.empty((4, 2), np.uint64)
    rc = g.lib.smem_gpu_collect_packed(g.h, C.byref(pr.desc), C.byref(sg.SeedOpt()), C.c_void_p(small.ctypes.data), C.c_int64(4),
                                       C.c_void_p(roff.ctypes.data), C.byref(tot))
    assert rc == -5 and tot.value > 4 and int(roff[-1]) == tot.value
