"""Extended variant smoke (was written for compute-sanitizer, which is closed on this pool; plain runs only): every kernel variant
(lanes_per_read 4, 2, 3, 1), the ordinary and the latency path, collect / smem1 / trace, overflow re-run, spill path, wide entries, the
three shortcuts with their tables, the compact wire formats, sa / seeds (walk and full suffix array) / chains -- all against the oracle."""
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
o = Oracle(ix)
batches = []
for n in (600, 2600):                                  # latency path (<= 2048 reads) and ordinary path
    seq, offs = sy.to_batch(sy.simulate_reads(ref, n, 101, 0.02, seed=n, n_frac=0.1))
    batches.append((seq, offs, o.collect(seq, offs, SeedOpt(), nthreads=4)))
for lpr, devices in ((4, [0]), (4, [0, 0]), (2, [0]), (3, [0]), (1, [0])):
    g = sg.SmemGpu(max_batch_reads=4096, max_read_len=128, devices=devices)
    g.set_param("lanes_per_read", lpr)
    g.upload_index(ix); g.upload_sa(ix)
    g.build_repeat_filter(ref, 12, 0)                  # filter + window flags + speculative walk are on from here
    g.build_text_index(ref)                            # unique walks + seed positions from the full suffix array
    for seq, offs, want in batches:
        n = len(offs) - 1
        for name, val in (("slot_cap", 128), ("slot_cap", 3), ("b_cap", 2), ("b_cap", 17), ("force_wide", 1), ("force_wide", 0)):
            g.set_param(name, val)
            got = g.collect(seq, offs)
            assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"]), (lpr, devices, n, name, val)
        g.set_param("slot_cap", 128)
        pr = sg.PackedReads(g.lib, seq, offs)
        for got in (g.collect_packed(pr), g.collect_packed12(pr)):
            assert np.array_equal(got["intv"], want["intv"]) and np.array_equal(got["read_off"], want["read_off"]), (lpr, n, "packed")
        x = np.random.default_rng(1).integers(0, 101, n).astype(np.int32)
        s1, s2 = g.smem1(seq, offs, x, np.ones(n, np.int32)), o.smem1(seq, offs, x, np.ones(n, np.int32))
        assert np.array_equal(s1["intv"], s2["intv"]) and np.array_equal(s1["ret"], s2["ret"])
        tr = g.trace(seq, offs)
        assert tr["read_off"][-1] == len(tr["intv"])
        g.collect(seq, offs)
        sd = g.seeds(n)
        ch = g.chains(n, ix.seq_len // 2)
        assert ch["chain_off"][-1] == len(ch["chains"]) and ch["chains"]["n_seeds"].sum() == len(ch["seeds"])
        g.set_param("sa_from_tables", 0)
        g.collect(seq, offs)
        sd2 = g.seeds(n)
        g.set_param("sa_from_tables", 1)
        assert np.array_equal(sd["seeds"]["rbeg"], sd2["seeds"]["rbeg"])
    k = np.random.default_rng(2).integers(0, ix.seq_len + 1, 2000).astype(np.uint64)
    assert np.array_equal(g.sa(k), o.sa(ix, k))
    g.close()
print("variant smoke ok")
