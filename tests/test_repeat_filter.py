"""The two exact shortcuts of the seeding kernel: the repeat filter of the re-seeding pass
(csrc/smem_repeat.cuh) and the speculative longest-only backward walk (PH_SPEC, csrc/smem_kernels.cuh).
CPU: the rule itself -- "if max>>1 >= K and every K-mer window through the re-seeding position occurs at most once, the
second pass of smem_next2 contributes nothing" -- modelled in the oracle with exact counts and compared with the unmodified
algorithm.  GPU: the device table against brute force, and seeding with the filter against the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import pkg
from oracle.binding import Oracle, SeedOpt
from test_chains import repeat_rich_reference

SETS = [(101, 0.01, SeedOpt()), (250, 0.02, SeedOpt()), (101, 0.03, SeedOpt(min_seed_len=15, split_factor=1.2, split_width=4)),
        (151, 0.005, SeedOpt(start_width=2))]


@pytest.fixture(scope="module")
def worlds(fm, synth):
    out = []
    for name, ref in (("random", synth.make_reference(500_000, 21)), ("repeats", repeat_rich_reference(300_000, 77))):
        ix = fm.build_index(ref)
        out.append((name, ref, ix, Oracle(ix)))
    return out


def test_skip_rule_is_exact_on_cpu(worlds, synth):
    for name, ref, ix, o in worlds:
        o.lib.orc_get_pass2_skipped.restype = C.c_uint64
        total_skipped = 0
        for rl, err, opt in SETS:
            seq, offs = synth.to_batch(synth.simulate_reads(ref, 1200, rl, err, seed=9, paired=True, n_frac=0.05))
            o.lib.orc_set_skip_kmer(0)
            want = o.collect(seq, offs, opt, nthreads=2)
            for K in (12, 15, 19):
                o.lib.orc_set_skip_kmer(K)
                try:
                    got = o.collect(seq, offs, opt, nthreads=2)
                    total_skipped += o.lib.orc_get_pass2_skipped()
                finally:
                    o.lib.orc_set_skip_kmer(0)
                for k in ("intv", "read_off", "step"):
                    assert np.array_equal(got[k], want[k]), (name, rl, K, k)
        assert total_skipped > 1000          # the rule fires (and is exercised) on these sets


def test_spec_walk_rule_is_exact_on_cpu(worlds, synth):
    """Model of PH_SPEC in the oracle (smem_oracle.c, orc_set_spec_walk) against the unmodified algorithm, alone and
    together with the pass-2 rule."""
    for name, ref, ix, o in worlds:
        o.lib.orc_get_spec_hits.restype = C.c_uint64
        hits = 0
        for rl, err, opt in SETS:
            seq, offs = synth.to_batch(synth.simulate_reads(ref, 1200, rl, err, seed=29, paired=True, n_frac=0.05))
            o.lib.orc_set_skip_kmer(0); o.lib.orc_set_spec_walk(0)
            want = o.collect(seq, offs, opt, nthreads=2)
            for K in (0, 15):
                o.lib.orc_set_skip_kmer(K); o.lib.orc_set_spec_walk(1)
                try:
                    got = o.collect(seq, offs, opt, nthreads=2)
                    hits += o.lib.orc_get_spec_hits()
                finally:
                    o.lib.orc_set_skip_kmer(0); o.lib.orc_set_spec_walk(0)
                for k in ("intv", "read_off", "step", "n_steps", "last_start"):
                    assert np.array_equal(got[k], want[k]), (name, rl, K, k)
        assert hits > 1000


def test_split_backward_sweep_rule_is_exact_on_cpu(worlds, synth):
    """Model of the device kernel's split backward sweep (smem_oracle.c, orc_set_split_pairs): the entries of a
    backward round extended two at a time, the push / emit decisions replayed in order from the two new sizes -- against the
    sequential loop of bwt.c:812-825, alone and together with the other shortcuts' models, on random and repeat-rich texts."""
    for name, ref, ix, o in worlds:
        o.lib.orc_get_split_trips.restype = C.c_uint64
        trips = 0
        for rl, err, opt in SETS:
            seq, offs = synth.to_batch(synth.simulate_reads(ref, 1200, rl, err, seed=31, paired=True, n_frac=0.05))
            o.lib.orc_set_skip_kmer(0); o.lib.orc_set_spec_walk(0); o.lib.orc_set_split_pairs(0)
            want = o.collect(seq, offs, opt, nthreads=2)
            for K, spec in ((0, 0), (15, 1)):
                o.lib.orc_set_skip_kmer(K); o.lib.orc_set_spec_walk(spec); o.lib.orc_set_split_pairs(1)
                try:
                    got = o.collect(seq, offs, opt, nthreads=2)
                    trips += o.lib.orc_get_split_trips()
                finally:
                    o.lib.orc_set_skip_kmer(0); o.lib.orc_set_spec_walk(0); o.lib.orc_set_split_pairs(0)
                for k in ("intv", "read_off", "step", "n_steps", "last_start"):
                    assert np.array_equal(got[k], want[k]), (name, rl, K, k)
            # raw calls too (every x, several min_intv)
            n = len(offs) - 1
            rng = np.random.default_rng(rl)
            x = rng.integers(0, rl, n).astype(np.int32); mi = rng.integers(0, 5, n).astype(np.int32)
            w1 = o.smem1(seq, offs, x, mi)
            o.lib.orc_set_split_pairs(1)
            try:
                g1 = o.smem1(seq, offs, x, mi)
            finally:
                o.lib.orc_set_split_pairs(0)
            for k in ("intv", "read_off", "ret"):
                assert np.array_equal(g1[k], w1[k]), (name, rl, k)
        assert trips > 100_000


def test_unique_walk_counting_model_leaves_results_alone(worlds, synth):
    """orc_set_unique_walk only changes the oracle's extend / block counts (bench.py's `executed` figure), never a result."""
    for name, ref, ix, o in worlds:
        seq, offs = synth.to_batch(synth.simulate_reads(ref, 1500, 101, 0.01, seed=31, paired=True, n_frac=0.05))
        want = o.collect(seq, offs, SeedOpt(), nthreads=2, stats=True)
        o.lib.orc_set_unique_walk(1)
        try:
            got = o.collect(seq, offs, SeedOpt(), nthreads=2, stats=True)
        finally:
            o.lib.orc_set_unique_walk(0)
        for k in ("intv", "read_off", "step"):
            assert np.array_equal(got[k], want[k]), (name, k)
        assert got["stats"]["extends"] <= want["stats"]["extends"] and got["stats"]["intervals"] == want["stats"]["intervals"]
        if name == "random":
            assert got["stats"]["extends"] < 0.9 * want["stats"]["extends"]      # long unique matches: the walks are taken


def _text_tables(o, ix, ref):
    """T = forward + reverse complement (one base per byte), its suffix array from the oracle's bwt_sa, and the inverse."""
    n = int(ix.seq_len)
    fsa = np.empty(n + 1, np.uint64)
    fsa[0] = n                                                            # the '$' row
    fsa[1:] = o.sa(ix, np.arange(1, n + 1, dtype=np.uint64))
    isa = np.empty(n + 1, np.uint64)
    isa[fsa.astype(np.int64)] = np.arange(n + 1, dtype=np.uint64)
    r = ref.numpy() if hasattr(ref, "numpy") else np.asarray(ref)
    T = np.ascontiguousarray(np.concatenate([r, (3 - r)[::-1]]), np.uint8)
    return T, fsa, isa


def _edge_reads(ref, T, rng):
    """Ragged lengths, both strands, one substitution / one N, the forward / reverse-complement junction, the text ends."""
    r = ref.numpy() if hasattr(ref, "numpy") else np.asarray(ref)
    reads = [np.zeros(0, np.uint8)]
    for ln in (7, 20, 33, 60, 101, 150, 255, 256):
        for _ in range(8):
            p = int(rng.integers(0, len(r) - ln)); q = r[p:p + ln].copy()
            if rng.random() < 0.5:
                q = (3 - q)[::-1].copy()
            if rng.random() < 0.5 and ln > 30:
                k = int(rng.integers(0, ln)); q[k] = (q[k] + 1) % 4
            if rng.random() < 0.2 and ln > 30:
                q[int(rng.integers(0, ln))] = 4
            reads.append(q)
    for off in (-90, -40, -1, 0, 30):
        reads.append(T[len(r) + off - 100: len(r) + off + 100].copy())
    reads.append(T[:180].copy()); reads.append(T[-180:].copy())
    return reads


def test_unique_walk_rule_is_exact_on_cpu(fm, synth):
    """Executable model of PH_UW_* in the oracle (smem_oracle.c, orc_set_unique_walk_tables: suffix array lookup, text
    comparison, inverse lookup instead of bwt_extend) against the unmodified algorithm: random and repeat-rich texts,
    every option set, edge reads, both start thresholds."""
    for name, ref in (("random", synth.make_reference(300_000, 5)), ("repeats", repeat_rich_reference(200_000, 8))):
        ix = fm.build_index(ref, sa_intv=32)
        o = Oracle(ix)
        o.lib.orc_get_unique_walks.restype = C.c_uint64
        T, fsa, isa = _text_tables(o, ix, ref)
        n = int(ix.seq_len)
        assert np.array_equal(np.sort(fsa), np.arange(n + 1, dtype=np.uint64))       # a permutation of the text positions
        sets = [synth.to_batch(_edge_reads(ref, T, np.random.default_rng(4))) + (SeedOpt(),)]
        for rl, err, opt in SETS:
            sets.append(synth.to_batch(synth.simulate_reads(ref, 1200, rl, err, seed=37, paired=True, n_frac=0.05)) + (opt,))
        walks = 0
        for seq, offs, opt in sets:
            want = o.collect(seq, offs, opt, nthreads=2)
            for run, left in ((3, 8), (1, 1)):
                o.lib.orc_set_unique_walk_tables(C.c_void_p(T.ctypes.data), C.c_void_p(fsa.ctypes.data), C.c_void_p(isa.ctypes.data),
                                                 C.c_uint64(n), C.c_int(run), C.c_int(left))
                try:
                    got = o.collect(seq, offs, opt, nthreads=2)
                    walks += o.lib.orc_get_unique_walks()
                finally:
                    o.lib.orc_set_unique_walk_tables(None, None, None, C.c_uint64(0), C.c_int(0), C.c_int(0))
                for k in ("intv", "read_off", "step", "n_steps", "last_start"):
                    assert np.array_equal(got[k], want[k]), (name, k, run, left)
        assert walks > 1000                  # the rule fires on these sets


def _brute_bits(T, K, log2_bits):
    n = len(T)
    codes = np.zeros(n - K + 1, np.uint64)
    for k in range(K):
        codes = (codes << np.uint64(2)) | T[k:n - K + 1 + k].astype(np.uint64)
    u, c = np.unique(codes, return_counts=True)
    rep = u[c > 1]
    idx = (rep * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(64 - log2_bits)
    bits = np.zeros(1 << (log2_bits - 5), np.uint32)
    np.bitwise_or.at(bits, (idx >> np.uint64(5)).astype(np.int64), (np.uint32(1) << (idx & np.uint64(31)).astype(np.uint32)))
    return bits, len(rep)


@pytest.mark.gpu
def test_gpu_repeat_filter_table_vs_brute_force(fm):
    sg = pkg("smem_gpu")
    for seed, K, lb in ((1, 12, 20), (2, 9, 16), (3, 16, 22)):
        ref = repeat_rich_reference(120_000, seed)
        T = fm.text_from_forward(ref).numpy()
        g = sg.SmemGpu(max_batch_reads=64, max_read_len=128)
        g.build_repeat_filter(ref, K, lb)
        assert g.get_param("rf_kmer") == K and g.get_param("rf_log2_bits") == lb
        want, n_rep = _brute_bits(T, K, lb)
        assert n_rep > 100
        assert np.array_equal(g.repeat_filter_bits(), want)
        # auto-sized: the table folded down from 8 bits per text position == the brute-force table of the size it chose,
        # and at most 1/256 of its bits are set (or it is as large as it gets)
        g.build_repeat_filter(ref, K, 0)
        lb2 = g.get_param("rf_log2_bits")
        got = g.repeat_filter_bits()
        assert np.array_equal(got, _brute_bits(T, K, lb2)[0])
        fill = sum(bin(int(w)).count("1") for w in got) / float(1 << lb2)
        assert fill <= 1 / 256 or lb2 == 21, (lb2, fill)
        g.close()


@pytest.mark.gpu
def test_gpu_collect_with_repeat_filter_vs_oracle(worlds, synth):
    sg = pkg("smem_gpu")
    for name, ref, ix, o in worlds:
        for devices in ([0], [0, 0]):
            g = sg.SmemGpu(max_batch_reads=4096, max_read_len=256, devices=devices)
            g.upload_index(ix)
            g.build_repeat_filter(ref, 14, 0 if name == "random" else 24)      # auto-sized / explicit table
            g.set_param("count_skips", 1)
            skipped = 0
            for rl, err, opt in SETS:
                seq, offs = synth.to_batch(synth.simulate_reads(ref, 3000, rl, err, seed=19, paired=True, n_frac=0.05))
                want = o.collect(seq, offs, opt, nthreads=4)
                gopt = sg.SeedOpt(opt.min_seed_len, opt.split_factor, opt.split_width, opt.start_width)
                got = g.collect(seq, offs, gopt)
                skipped += g.get_param("pass2_skipped")
                for k in ("intv", "read_off", "step"):
                    assert np.array_equal(got[k], want[k]), (name, rl, k)
                g.set_param("repeat_filter", 0)               # and the same answer with the filter switched off
                off = g.collect(seq, offs, gopt)
                assert g.get_param("pass2_skipped") == 0
                assert np.array_equal(off["intv"], want["intv"])
                g.set_param("spec_walk", 0)                   # ... and with both shortcuts off
                off2 = g.collect(seq, offs, gopt)
                assert np.array_equal(off2["intv"], want["intv"]) and np.array_equal(off2["step"], want["step"])
                # raw bwt_smem1 lists (TRACE) are never shortened by the filter and are the same with and without the walk
                tr0 = g.trace(seq[: offs[200]], offs[:201], gopt)
                g.set_param("repeat_filter", 1); g.set_param("spec_walk", 1)
                tr = g.trace(seq[: offs[200]], offs[:201], gopt)
                assert all(np.array_equal(tr[k], tr0[k]) for k in tr0)
            assert skipped > 2000
            g.close()


@pytest.mark.gpu
def test_gpu_repeat_filter_ragged_and_edge_reads(worlds, synth):
    """Empty / shorter than the k-mer / exactly k / 2k-1 / 2k / long error-free reads, ambiguous bases next to the re-seeding
    position, reads from both strands and across the forward/reverse junction -- filter and walk on, against the oracle."""
    sg = pkg("smem_gpu")
    name, ref, ix, o = worlds[0]
    r = ref.numpy()
    K = 13
    rng = np.random.default_rng(3)
    reads = [np.zeros(0, np.uint8), np.full(40, 4, np.uint8)]
    for ln in (1, 5, K - 1, K, K + 1, 2 * K - 1, 2 * K, 2 * K + 1, 60, 101, 255, 256):
        for _ in range(6):
            p = int(rng.integers(0, len(r) - ln))
            q = r[p:p + ln].copy()
            if rng.random() < 0.5:
                q = (3 - q)[::-1].copy()
            reads.append(q)
    for _ in range(40):                                   # one N close to the middle of a long exact match
        p = int(rng.integers(0, len(r) - 120)); q = r[p:p + 120].copy(); q[60 + int(rng.integers(-K, K))] = 4; reads.append(q)
    for _ in range(20):                                   # two substitutions K-ish apart
        p = int(rng.integers(0, len(r) - 150)); q = r[p:p + 150].copy()
        a = int(rng.integers(20, 100)); q[a] = (q[a] + 1) % 4; q[a + int(rng.integers(1, 2 * K))] ^= 1; reads.append(q)
    T = np.concatenate([r, (3 - r)[::-1]])
    for off in (-60, -30, -1, 0):                         # across the junction of forward text and reverse complement
        reads.append(T[len(r) + off - 50: len(r) + off + 51].copy())
    seq, offs = synth.to_batch(reads)
    want = o.collect(seq, offs, SeedOpt(), nthreads=2)
    import torch
    two = [[0, 1]] if torch.cuda.device_count() >= 2 else []          # one handle over two GPUs: the filter is built on both
    for devices in [[0], [0, 0, 0]] + two:
        g = sg.SmemGpu(max_batch_reads=1024, max_read_len=256, devices=devices)
        g.upload_index(ix)
        g.build_repeat_filter(ref, K, 0)
        g.set_param("count_skips", 1)
        got = g.collect(seq, offs)
        for k in ("intv", "read_off", "step"):
            assert np.array_equal(got[k], want[k]), k
        assert g.get_param("pass2_skipped") > 20
        g.close()


@pytest.mark.gpu
def test_gpu_repeat_filter_of_another_text_is_refused_or_ignored(worlds, synth, fm):
    sg = pkg("smem_gpu")
    name, ref, ix, o = worlds[0]
    other = synth.make_reference(300_000, 99)
    g = sg.SmemGpu(max_batch_reads=2048, max_read_len=128)
    g.upload_index(ix)
    with pytest.raises(sg.SmemGpuError):
        g.build_repeat_filter(other, 13, 0)                   # text length differs from the uploaded index
    g.build_repeat_filter(ref, 13, 0)
    ix2 = fm.build_index(other)
    g.upload_index(ix2)                                       # new index, old filter: must not be used
    seq, offs = synth.to_batch(synth.simulate_reads(other, 1500, 101, 0.01, seed=5, paired=True))
    g.set_param("count_skips", 1)
    got = g.collect(seq, offs)
    assert g.get_param("pass2_skipped") == 0
    want = Oracle(ix2).collect(seq, offs, SeedOpt(), nthreads=2)
    assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"])
    g.close()


@pytest.mark.gpu
def test_gpu_text_index_and_unique_walk(fm, synth):
    """Unique-walk tables (full SA / inverse from the samples on the device) against the oracle's bwt_sa, and seeding with
    the walk on against the oracle: random + repeat-rich text, both strands, junction, N, ragged lengths, TRACE."""
    sg = pkg("smem_gpu")
    for name, ref, shift in (("random", synth.make_reference(300_000, 5), 2), ("repeats", repeat_rich_reference(200_000, 8), 3),
                             ("random-full-isa", synth.make_reference(120_001, 6), 0)):
        ix = fm.build_index(ref, sa_intv=32)
        o = Oracle(ix)
        g = sg.SmemGpu(max_batch_reads=4096, max_read_len=256, devices=[0, 0])
        g.upload_index(ix); g.upload_sa(ix)
        g.set_param("unique_walk_isa_shift", shift)                       # sampling distance of the inverse suffix array (PH_UW_LF steps)
        g.build_text_index(ref)
        n = ix.seq_len
        fsa, isa_s = g.text_index(0), g.text_index(1)                     # 33-bit tables, unpacked by the test hook
        rows = np.arange(1, n + 1, dtype=np.uint64)
        assert np.array_equal(fsa[1:], o.sa(ix, rows))                    # every row but the '$' row
        assert int(fsa[0]) == n
        isa = np.empty(n + 1, np.uint64); isa[fsa.astype(np.int64)] = np.arange(n + 1, dtype=np.uint64)   # inverse of the full table
        k = 1 << g.get_param("unique_walk_isa_shift")                     # samples: position n - e * k for e = 0, 1, ...
        assert len(isa_s) == n // k + 1 and np.array_equal(isa_s, isa[n - k * np.arange(n // k + 1)])
        assert g.get_param("uw_table_bytes") < (0.5 + 4.58 * (1 + 1 / k)) * n + 8192   # text nibbles + 33-bit SA + sampled 33-bit inverse (6.2 bytes per position at k = 4)
        g.build_repeat_filter(ref, 13, 0)
        r = ref.numpy()
        T = np.concatenate([r, (3 - r)[::-1]])
        rng = np.random.default_rng(4)
        reads = [np.zeros(0, np.uint8)]
        for ln in (7, 20, 33, 60, 101, 150, 255, 256):
            for _ in range(8):
                p = int(rng.integers(0, len(r) - ln)); q = r[p:p + ln].copy()
                if rng.random() < 0.5:
                    q = (3 - q)[::-1].copy()
                if rng.random() < 0.5 and ln > 30:
                    q[int(rng.integers(0, ln))] = (q[int(rng.integers(0, ln))] + 1) % 4
                if rng.random() < 0.2 and ln > 30:
                    q[int(rng.integers(0, ln))] = 4
                reads.append(q)
        for off in (-90, -40, -1, 0, 30):                                 # across the forward / reverse-complement junction
            reads.append(T[len(r) + off - 100: len(r) + off + 100].copy())
        reads.append(T[:180].copy()); reads.append(T[-180:].copy())       # the two ends of the text
        seq1, offs1 = synth.to_batch(reads)
        sets = [(seq1, offs1, SeedOpt())]
        for rl, err, opt in SETS:
            s_, o_ = synth.to_batch(synth.simulate_reads(ref, 2500, rl, err, seed=23, paired=True, n_frac=0.05))
            sets.append((s_, o_, opt))
        for seq, offs, opt in sets:
            want = o.collect(seq, offs, opt, nthreads=4)
            gopt = sg.SeedOpt(opt.min_seed_len, opt.split_factor, opt.split_width, opt.start_width)
            for run, left in ((3, 8), (1, 1), (2, 40)):                    # when a walk starts: extends of a unique interval, bases left
                g.set_param("unique_walk_min_run", run); g.set_param("unique_walk_min_left", left)
                got = g.collect(seq, offs, gopt)
                for k in ("intv", "read_off", "step"):
                    assert np.array_equal(got[k], want[k]), (name, k, run, left)
            g.set_param("unique_walk_min_run", 3); g.set_param("unique_walk_min_left", 8)
            tr = g.trace(seq, offs, gopt)
            g.set_param("unique_walk", 0)
            tr0 = g.trace(seq, offs, gopt)
            g.set_param("unique_walk", 1)
            assert all(np.array_equal(tr[k], tr0[k]) for k in tr0)
        g.close()
