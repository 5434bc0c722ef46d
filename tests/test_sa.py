"""Seed -> reference position: bwt_sa.  CPU: restatement vs the reference and the
builder's SA samples vs `bwa index`; GPU: the device walk vs the oracle, plus an end-to-end property."""
import os

import numpy as np
import pytest

from conftest import pkg
from oracle.binding import Oracle, Reference, SeedOpt, build_reference


@pytest.fixture(scope="module")
def world_sa(fm, synth):
    ref = synth.make_reference(400_000, 31)
    ix = fm.build_index(ref, sa_intv=32)
    return ref, ix, Oracle(ix)


def _rows(ix, n, seed):
    rng = np.random.default_rng(seed)
    edge = [0, 1, 31, 32, 33, ix.primary, max(ix.primary - 1, 0), ix.primary + 1, ix.seq_len, ix.seq_len - 1]
    return np.concatenate([rng.integers(0, ix.seq_len + 1, n), edge]).astype(np.uint64)


@pytest.mark.skipif(build_reference() is None, reason="reference objects absent")
def test_oracle_sa_matches_reference(world_sa):
    ref, ix, o = world_sa
    k = _rows(ix, 20000, 1)
    assert np.array_equal(o.sa(ix, k), Reference(ix).sa(ix, k))


def test_sa_samples_spell_the_text(world_sa, fm):
    """SA[r] from the samples + walk must be a permutation consistent with the text: T[SA[r]:] sorted."""
    ref, ix, o = world_sa
    rows = np.arange(1, 4000, dtype=np.uint64)
    sa = o.sa(ix, rows).astype(np.int64)
    T = fm.text_from_forward(ref).numpy()
    pre = [bytes(T[p:p + 24]) for p in sa]
    assert all(pre[i] <= pre[i + 1] for i in range(len(pre) - 1))


@pytest.mark.gpu
def test_gpu_sa_vs_oracle_and_seed_positions(world_sa, synth):
    sg = pkg("smem_gpu")
    ref, ix, o = world_sa
    g = sg.SmemGpu(max_batch_reads=8192, max_read_len=128, devices=[0, 0])
    g.upload_index(ix)
    g.upload_sa(ix)
    k = _rows(ix, 50000, 2)
    assert np.array_equal(g.sa(k), o.sa(ix, k))
    assert len(g.sa(np.zeros(0, np.uint64))) == 0
    # end to end: error-free forward reads -> full-length SMEM -> SA of its rows contains the sampling position
    rng = np.random.default_rng(5)
    pos = rng.integers(0, len(ref) - 101, 500)
    reads = np.stack([ref.numpy()[p:p + 101] for p in pos])
    seq, offs = synth.to_batch(reads)
    res = g.collect(seq, offs)
    rows, owner = [], []
    for i in range(len(pos)):
        for iv in res["intv"][res["read_off"][i]:res["read_off"][i + 1]]:
            if int(iv[3]) >> 32 == 0 and int(iv[3]) & 0xFFFFFFFF == 101:
                for t in range(int(iv[2])):
                    rows.append(int(iv[0]) + t); owner.append(i)
    sa = g.sa(np.array(rows, np.uint64))
    hit = np.zeros(len(pos), bool)
    for r, i in zip(sa, owner):
        hit[i] |= int(r) == int(pos[i])
    assert hit.all()
    g.close()


@pytest.mark.gpu
def test_gpu_seeds_vs_oracle(world_sa, synth):
    """smem_gpu_seeds == the seed loop of mem_insert_seed (bwamem.c:462-476) rebuilt from oracle intervals + oracle bwt_sa."""
    sg = pkg("smem_gpu")
    ref, ix, o = world_sa
    refn = ref.numpy().copy()
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 6001, 101, 0.02, seed=11, n_frac=0.05))
    for devices, (min_seed_len, max_occ) in (([0], (19, 10000)), ([0, 0, 0], (25, 3))):
        g = sg.SmemGpu(max_batch_reads=8192, max_read_len=128, devices=devices)
        g.upload_index(ix); g.upload_sa(ix)
        res = g.collect(seq, offs)
        got = g.seeds(len(offs) - 1, min_seed_len, max_occ)
        want = o.collect(seq, offs, SeedOpt(), nthreads=4)
        rows, qb, ln, cnt = [], [], [], np.zeros(len(offs) - 1, np.int64)
        for i in range(len(offs) - 1):
            for iv in want["intv"][want["read_off"][i]:want["read_off"][i + 1]]:
                s, e = int(iv[3]) >> 32, int(iv[3]) & 0xFFFFFFFF
                if e - s < min_seed_len or int(iv[2]) > max_occ:
                    continue
                for t in range(int(iv[2])):
                    rows.append(int(iv[0]) + t); qb.append(s); ln.append(e - s); cnt[i] += 1
        rbeg = o.sa(ix, np.array(rows, np.uint64)).astype(np.int64)
        assert np.array_equal(got["seed_off"], np.concatenate([[0], np.cumsum(cnt)]))
        assert np.array_equal(got["seeds"]["rbeg"], rbeg)
        assert np.array_equal(got["seeds"]["qbeg"], np.array(qb, np.int32)) and np.array_equal(got["seeds"]["len"], np.array(ln, np.int32))
        g.close()


@pytest.mark.gpu
def test_gpu_sa_and_seeds_from_the_full_suffix_array(world_sa, synth):
    """With the unique-walk tables resident bwt_sa is one gather from the full suffix array (seed_expand_fsa_kernel,
    sa_from_fsa_kernel) instead of a walk of bwt_invPsi steps: same positions, row 0 included (sa[0] = -1, bwt.c:97)."""
    sg = pkg("smem_gpu")
    ref, ix, o = world_sa
    g = sg.SmemGpu(max_batch_reads=8192, max_read_len=128)
    g.upload_index(ix); g.upload_sa(ix)
    g.build_text_index(ref)
    k = np.concatenate([_rows(ix, 40000, 3), np.array([0, 1, ix.seq_len], np.uint64)])
    want = o.sa(ix, k)
    assert np.array_equal(g.sa(k), want)
    g.set_param("sa_from_tables", 0)
    assert np.array_equal(g.sa(k), want)                       # the walk, for comparison
    seq, offs = synth.to_batch(synth.simulate_reads(ref, 5000, 101, 0.02, seed=12, n_frac=0.05))
    g.collect(seq, offs)
    a = g.seeds(len(offs) - 1, 19, 10000)
    g.set_param("sa_from_tables", 1)
    g.collect(seq, offs)
    b = g.seeds(len(offs) - 1, 19, 10000)
    assert np.array_equal(a["seed_off"], b["seed_off"])
    for f in ("rbeg", "qbeg", "len"):
        assert np.array_equal(a["seeds"][f], b["seeds"][f]), f
    g.close()
