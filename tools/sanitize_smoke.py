"""Small end-to-end run for compute-sanitizer: collect, smem1, overflow re-run, spill path, sa, seeds, chains, repeat filter, trace."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
sy = importlib.import_module("bwa-mem-harp2_b200.synth")
from oracle.binding import Oracle, SeedOpt
ref = sy.make_reference(60_000, 3)
ref[2000:2400] = ref[100:500]
ix = fm.build_index(ref, sa_intv=32)
seq, offs = sy.to_batch(sy.simulate_reads(ref, 600, 101, 0.02, seed=4, n_frac=0.1))
o = Oracle(ix)
want = o.collect(seq, offs, SeedOpt(), nthreads=4)
for devices in ([0], [0, 0]):
    g = sg.SmemGpu(max_batch_reads=1024, max_read_len=128, devices=devices)
    g.upload_index(ix); g.upload_sa(ix)
    g.build_repeat_filter(ref, 12, 0)                  # filter + window flags + speculative walk are on from here
    for name, val in (("slot_cap", 128), ("slot_cap", 3), ("b_cap", 2), ("force_wide", 1), ("blocks_per_sm", 6)):
        g.set_param(name, val)
        got = g.collect(seq, offs)
        assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"]), (devices, name, val)
    x = np.random.default_rng(1).integers(0, 101, 600).astype(np.int32)
    s1, s2 = g.smem1(seq, offs, x, np.ones(600, np.int32)), o.smem1(seq, offs, x, np.ones(600, np.int32))
    assert np.array_equal(s1["intv"], s2["intv"]) and np.array_equal(s1["ret"], s2["ret"])
    g.collect(seq, offs)
    sd = g.seeds(600)
    ch = g.chains(600, ix.seq_len // 2)
    assert ch["chain_off"][-1] == len(ch["chains"]) and ch["chains"]["n_seeds"].sum() == len(ch["seeds"])
    tr = g.trace(seq, offs)
    assert tr["read_off"][-1] == len(tr["intv"])
    g.build_repeat_filter(ref, 16, 20)
    assert np.array_equal(g.collect(seq, offs)["intv"], want["intv"])
    k = np.random.default_rng(2).integers(0, ix.seq_len + 1, 2000).astype(np.uint64)
    assert np.array_equal(g.sa(k), o.sa(ix, k))
    g.close()
print("sanitize smoke ok", len(want["intv"]), "intervals", len(sd["seeds"]), "seeds")
