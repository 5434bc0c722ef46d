"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact intervals.

Everything here calls libsmem_gpu.so through ctypes (include/smem_gpu.h); the oracle is only the checker.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, OPT_KEYS, GoldenIndex, pkg, same_result
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
    """1 Mbp random reference with planted repeats, its index, the oracle and a GPU handle."""
    ref = synth.make_reference(1_000_000, 77)
    ref[50_000:52_000] = ref[10_000:12_000]
    ref[400_000:400_300] = 1
    ix = fm.build_index(ref)
    g = sg.SmemGpu(max_batch_reads=40_000, max_read_len=260)
    g.upload_index(ix)
    return ref, ix, Oracle(ix), g


@pytest.mark.parametrize("key", OPT_KEYS)
def test_golden_collect(golden, sg, key):
    name, z, ix = golden
    g = sg.SmemGpu(max_batch_reads=1024, max_read_len=160)
    g.upload_index(ix)
    r = g.collect(z["seq"], z["offs"], gopt(z[f"opt_{key}"]))
    for k in ("intv", "read_off", "step"):
        assert np.array_equal(r[k], z[f"collect_{key}_{k}"]), (name, key, k)
    g.close()


@pytest.mark.parametrize("rep", [0, 1, 2])
def test_golden_smem1(golden, sg, rep):
    name, z, ix = golden
    g = sg.SmemGpu(max_batch_reads=1024, max_read_len=160)
    g.upload_index(ix)
    r = g.smem1(z["smem1_seq"], z["smem1_offs"], z[f"smem1_{rep}_x"], z[f"smem1_{rep}_min_intv"])
    for k in ("intv", "read_off", "ret"):
        assert np.array_equal(r[k], z[f"smem1_{rep}_{k}"]), (name, rep, k)
    g.close()


@pytest.mark.parametrize("n,L,err,nfrac,opt", [
    (20000, 101, 0.01, 0.06, (19, 1.5, 10, 1)),
    (6000, 250, 0.02, 0.06, (19, 1.5, 10, 1)),       # BASELINE config 4 shape: re-seeding on
    (5000, 101, 0.01, 0.0, (19, 1.5, 10, 2)),        # MEM_F_NO_EXACT
    (5000, 60, 0.05, 0.3, (10, 1.2, 20, 1)),
    (3000, 150, 0.0, 0.0, (19, 1.5, 0, 1)),          # split_width 0: never re-seeds
    (3000, 36, 0.1, 0.0, (19, 1.5, 10, 1)),
])
def test_collect_vs_oracle(world, synth, n, L, err, nfrac, opt):
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, L, err, seed=n + L, n_frac=nfrac))
    a = o.collect(seq, offs, OSeedOpt(*opt), nthreads=8)
    b = g.collect(seq, offs, gopt(opt))
    same_result(a, b, ("read_off", "intv", "step"))
    t = g.timing()
    assert t["kernel_launches"] >= 3 and t["seed_kernel_ms"] > 0


def test_smem1_vs_oracle(world, synth):
    ref, ix, o, g = world
    n = 12000
    seq, offs = synth.to_batch(synth.simulate_reads(ref, n, 101, 0.02, seed=5, n_frac=0.1))
    rng = np.random.default_rng(8)
    x = rng.integers(0, 101, n).astype(np.int32)
    mi = rng.integers(0, 5, n).astype(np.int32)
    a, b = o.smem1(seq, offs, x, mi), g.smem1(seq, offs, x, mi)
    same_result(a, b, ("read_off", "intv", "ret"))


def test_ragged_and_degenerate_reads(world, synth):
    ref, ix, o, g = world
    rng = np.random.default_rng(2)
    refn = ref.numpy()
    reads = [np.zeros(0, np.uint8), np.array([4], np.uint8), np.array([3], np.uint8), np.full(50, 4, np.uint8),
             refn[400_000:400_200].copy(), refn[10_000:10_250].copy(), np.zeros(120, np.uint8)]
    for _ in range(500):
        L = int(rng.integers(1, 256))
        p = int(rng.integers(0, len(refn) - L))
        r = refn[p:p + L].copy()
        if rng.random() < 0.3:
            r[rng.integers(0, L, 3)] = 4
        if rng.random() < 0.5:
            r = (3 - np.minimum(r, 3))[::-1].copy() if r.max(initial=0) < 4 else r
        reads.append(r)
    seq, offs = synth.to_batch(reads)
    a = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    b = g.collect(seq, offs)
    same_result(a, b, ("read_off", "intv", "step"))
    # empty batch
    e = g.collect(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(e["intv"]) == 0 and list(e["read_off"]) == [0]


def test_slot_overflow_rerun_and_knobs(world, synth):
    """Results must not depend on launch geometry, slot capacity (forces the overflow re-run) or L2 hints."""
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 4000, 101, 0.03, seed=31, n_frac=0.05))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    try:
        for name, val in [("slot_cap", 2), ("slot_cap", 128), ("blocks_per_sm", 6), ("blocks_per_sm", 9), ("blocks_per_sm", 8),
                          ("b_cap", 2), ("b_cap", 7), ("force_wide", 1), ("b_cap", 19), ("force_wide", 0),
                          ("l2_hot_min_intv", 64), ("l2_hot_min_intv", 0)]:
            g.set_param(name, val)
            b = g.collect(seq, offs)
            same_result(a, b, ("read_off", "intv", "step"))
            if name == "slot_cap" and val == 2:
                assert g.timing()["overflow_reads"] > 0
        with pytest.raises(pkg("smem_gpu").SmemGpuError):      # only launch-bounds variants that were compiled in are accepted
            g.set_param("blocks_per_sm", 12)
    finally:
        g.set_param("slot_cap", 128); g.set_param("blocks_per_sm", 8); g.set_param("l2_hot_min_intv", 0)
        g.set_param("b_cap", 19); g.set_param("force_wide", 0)


def test_staged_run_is_idempotent_and_capacity_error(world, synth, sg):
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 3000, 101, 0.01, seed=1))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    g.stage(seq, offs)
    t1 = g.run_collect()
    t2 = g.run_collect()
    assert t1 == t2 == len(a["intv"])
    b = g.fetch(t2, want_step=True)
    same_result(a, b, ("read_off", "intv", "step"))
    assert o.checksum(b["intv"], b["read_off"]) == o.checksum(a["intv"], a["read_off"])
    # too-small output buffer -> SMEM_GPU_E_CAPACITY with counts still reported
    with pytest.raises(sg.SmemGpuError) as ei:
        g.fetch(t2, intv=np.empty((10, 4), np.uint64))
    assert ei.value.code == -5
    with pytest.raises(sg.SmemGpuError):
        g.stage(np.zeros(50_000 * 10, np.uint8), np.arange(50_001, dtype=np.int64) * 10)   # > max_batch_reads
    with pytest.raises(sg.SmemGpuError) as ei:                                             # one read longer than max_read_len
        g.collect(np.zeros(400, np.uint8), np.array([0, 100, 400], np.int64))
    assert ei.value.code == -5
    with pytest.raises(sg.SmemGpuError):                                                   # offsets not monotone
        g.collect(np.zeros(400, np.uint8), np.array([0, 200, 100], np.int64))


def test_exact_reads_give_full_length_smem(world, synth):
    """Size-independent property: an error-free read has one SMEM spanning it, with x[2] >= 1."""
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 5000, 101, 0.0, seed=99))
    b = g.collect(seq, offs)
    ro = b["read_off"]
    for i in range(0, 5000, 7):
        iv = b["intv"][ro[i]:ro[i + 1]]
        full = [(int(v[3]) >> 32, int(v[3]) & 0xFFFFFFFF) for v in iv]
        assert (0, 101) in full
        assert all(int(v[2]) >= 1 for v in iv)
        starts = [s for s, _ in full]
        # per step the list is sorted by start; across steps starts never decrease below the previous step's first
        assert all(0 <= s < e <= 101 for s, e in full)


def test_two_gpu_handle_matches_single(world, synth, sg):
    """One handle spanning 2 GPUs (index replicated, reads sharded, host gather) == single GPU == oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 9001, 101, 0.02, seed=77, n_frac=0.05))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    g2 = sg.SmemGpu(max_batch_reads=10_000, max_read_len=128, devices=[0, 1])
    g2.upload_index(ix)
    b = g2.collect(seq, offs)
    same_result(a, b, ("read_off", "intv", "step"))
    rng = np.random.default_rng(5)
    x = rng.integers(0, 101, 9001).astype(np.int32)
    mi = rng.integers(0, 4, 9001).astype(np.int32)
    same_result(o.smem1(seq, offs, x, mi), g2.smem1(seq, offs, x, mi), ("read_off", "intv", "ret"))
    g2.close()


def test_pipeline_lanes_on_one_gpu(world, synth, sg):
    """A device listed several times = several pipeline lanes sharing one index copy; same results."""
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 7003, 101, 0.02, seed=123, n_frac=0.05))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=8)
    g4 = sg.SmemGpu(max_batch_reads=8_000, max_read_len=128, devices=[0, 0, 0, 0])
    g4.upload_index(ix)
    for _ in range(2):
        b = g4.collect(seq, offs)
        same_result(a, b, ("read_off", "intv", "step"))
    g4.upload_index(ix)            # re-upload must not leak or double-free the shared copy
    same_result(a, g4.collect(seq, offs), ("read_off", "intv", "step"))
    g4.close()


def test_trace_replays_to_collect_and_matches_raw_calls(world, synth):
    """smem_gpu_trace: (a) every recorded call equals a raw bwt_smem1 of the oracle at the position / min_intv the
    reference's smem_next2 would use, (b) re-doing smem_next2's merge on the host from the trace gives collect."""
    ref, ix, o, g = world
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 1500, 101, 0.02, seed=17, n_frac=0.08))
    opt = (19, 1.5, 10, 1)
    tr = g.trace(seq, offs, gopt(opt))
    want = o.collect(seq, offs, OSeedOpt(*opt), nthreads=8)
    split_len0 = int(opt[0] * opt[1] + .499)
    calls, merged_all, merged_off = [], [], [0]          # calls: (read, x, min_intv, expected list, expected ret)
    for i in range(len(offs) - 1):
        q = seq[offs[i]:offs[i + 1]]
        L = len(q)
        sl = tr["intv"][tr["read_off"][i]:tr["read_off"][i + 1]]
        tg = tr["tag"][tr["read_off"][i]:tr["read_off"][i + 1]]
        rt = tr["ret"][tr["read_off"][i]:tr["read_off"][i + 1]]
        start, s = 0, 0
        split_len = min(split_len0, L)
        while True:
            while start < L and q[start] > 3:
                start += 1
            if start >= L:
                break
            p1, r1 = sl[tg == 2 * s], rt[tg == 2 * s]
            assert len(p1) >= 1 and (r1 == r1[0]).all()
            calls.append((i, start, opt[3], p1, int(r1[0])))
            lens = (p1[:, 3] & 0xFFFFFFFF).astype(np.int64) - (p1[:, 3] >> 32).astype(np.int64)
            mi = int(np.argmax(lens))               # first maximum
            mx = int(lens[mi])
            out = [tuple(int(v) for v in e) for e in p1]
            if split_len > 0 and mx >= split_len and int(p1[mi, 2]) <= opt[2]:
                p2 = sl[tg == 2 * s + 1]
                mid = (int(p1[mi, 3] & 0xFFFFFFFF) + int(p1[mi, 3] >> 32)) >> 1
                calls.append((i, mid, int(p1[mi, 2]) + 1, p2, None))
                key = lambda e: ((e[3] >> 32) << 32) | ((L - (e[3] & 0xFFFFFFFF)) & 0xFFFFFFFF)
                keep = lambda e: ((e[3] & 0xFFFFFFFF) - (e[3] >> 32)) >= (mx >> 1) and (e[3] & 0xFFFFFFFF) > start
                a, b, out, ia, ib = out, [tuple(int(v) for v in e) for e in p2], [], 0, 0
                while ia < len(a) and ib < len(b):
                    if key(a[ia]) < key(b[ib]):
                        out.append(a[ia]); ia += 1
                    else:
                        if keep(b[ib]):
                            out.append(b[ib])
                        ib += 1
                out += a[ia:] + [e for e in b[ib:] if keep(e)]
            else:
                assert not (tg == 2 * s + 1).any()
            merged_all += out
            start = int(r1[0])
            s += 1
        assert not (tg >= 2 * s).any()
        merged_off.append(len(merged_all))
    assert np.array_equal(np.array(merged_off), want["read_off"])
    assert np.array_equal(np.array(merged_all, dtype=np.uint64).reshape(-1, 4), want["intv"])
    # raw calls against the oracle's bwt_smem1
    rseq, roffs = synth.to_batch([seq[offs[c[0]]:offs[c[0] + 1]] for c in calls])
    raw = o.smem1(rseq, roffs, [c[1] for c in calls], [c[2] for c in calls])
    for k, c in enumerate(calls):
        assert np.array_equal(raw["intv"][raw["read_off"][k]:raw["read_off"][k + 1]], c[3]), k
        if c[4] is not None:
            assert int(raw["ret"][k]) == c[4]


def test_latency_path_of_small_batches(world, synth, sg):
    """Batches of up to 2048 reads through the one-call forms take the latency path (one H2D copy, three launches, results stored
    into a pinned arena by the last kernel): same lists as the ordinary path and the oracle; slot overflow and an arena that is
    too small fall back to the ordinary path / the ordinary fetch."""
    ref, ix, o, g = world
    refn = ref.numpy()
    for n, L in ((1, 101), (64, 101), (700, 150), (2048, 101)):
        seq, offs = synth.to_batch(synth.simulate_reads(ref, n, L, 0.02, seed=n, n_frac=0.05))
        a = o.collect(seq, offs, OSeedOpt(), nthreads=4)
        same_result(a, g.collect(seq, offs), ("read_off", "intv", "step"))
        assert g.timing()["kernel_launches"] == 3                    # pack, seed, tail
        rng = np.random.default_rng(n)
        x = rng.integers(0, L, n).astype(np.int32); mi = rng.integers(0, 4, n).astype(np.int32)
        same_result(o.smem1(seq, offs, x, mi), g.smem1(seq, offs, x, mi), ("read_off", "intv", "ret"))
        t1 = g.trace(seq, offs)
        g.set_param("tiny_path", 0)
        try:
            t0 = g.trace(seq, offs)
            same_result(a, g.collect(seq, offs), ("read_off", "intv", "step"))
        finally:
            g.set_param("tiny_path", 1)
        for k in ("read_off", "intv", "tag", "ret"):
            assert np.array_equal(t0[k], t1[k]), k
    # reads that outgrow their slots: the latency path hands over to the ordinary one
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 500, 101, 0.03, seed=3))
    a = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    g.set_param("slot_cap", 2)
    try:
        same_result(a, g.collect(seq, offs), ("read_off", "intv", "step"))
        assert g.timing()["overflow_reads"] > 0
    finally:
        g.set_param("slot_cap", 128)
    # more intervals than the arena of a small handle holds (64 x 48 + 1024): ordinary fetch of the resident results
    g2 = sg.SmemGpu(max_batch_reads=64, max_read_len=260)
    g2.share_index_from(g)
    reads = []
    for k in range(64):
        q = np.full(260, 4, np.uint8); q[::2] = refn[1000 * k:1000 * k + 130]; reads.append(q)
    seq, offs = synth.to_batch(reads)
    a = o.collect(seq, offs, OSeedOpt(), nthreads=4)
    assert len(a["intv"]) > 64 * 48 + 1024
    same_result(a, g2.collect(seq, offs), ("read_off", "intv", "step"))
    # errors are reported, the handle stays usable
    bad, boffs = synth.to_batch([refn[:300].copy()])
    with pytest.raises(sg.SmemGpuError):
        g2.collect(bad, boffs)
    same_result(a, g2.collect(seq, offs), ("read_off", "intv", "step"))
    g2.close()
