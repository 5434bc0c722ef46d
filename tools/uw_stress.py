"""Randomized CPU stress of the unique-walk rule: the oracle's executable model (orc_set_unique_walk_tables)
against the unmodified algorithm over random / repeat-rich references, read lengths 36-400, 0-5 % errors, three option sets and
three start thresholds.  Last run: all exact over 582 579 walks.  Usage: python tools/uw_stress.py"""
import os, sys, importlib, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle.binding import Oracle, SeedOpt
fm = importlib.import_module("bwa-mem-harp2_b200.fmindex"); sy = importlib.import_module("bwa-mem-harp2_b200.synth")
from test_chains import repeat_rich_reference
from test_repeat_filter import _text_tables, _edge_reads
tot_walks = 0
for trial in range(6):
    rng = np.random.default_rng(100 + trial)
    n_ref = int(rng.integers(20_000, 400_000))
    ref = sy.make_reference(n_ref, 50 + trial) if trial % 2 == 0 else repeat_rich_reference(n_ref, 60 + trial)
    ix = fm.build_index(ref, sa_intv=32)
    o = Oracle(ix); o.lib.orc_get_unique_walks.restype = C.c_uint64
    T, fsa, isa = _text_tables(o, ix, ref); n = int(ix.seq_len)
    for rl, err in ((36, 0.0), (101, 0.0), (101, 0.02), (151, 0.01), (250, 0.05), (400, 0.01)):
        for opt in (SeedOpt(), SeedOpt(min_seed_len=10, split_factor=1.0, split_width=20), SeedOpt(start_width=2)):
            seq, offs = sy.to_batch(list(sy.simulate_reads(ref, 600, rl, err, seed=int(rng.integers(1 << 30)), paired=True, n_frac=0.03).cpu().numpy()) + _edge_reads(ref, T, rng))
            want = o.collect(seq, offs, opt, nthreads=4)
            for run, left in ((3, 8), (1, 1), (2, 30)):
                o.lib.orc_set_unique_walk_tables(C.c_void_p(T.ctypes.data), C.c_void_p(fsa.ctypes.data), C.c_void_p(isa.ctypes.data), C.c_uint64(n), C.c_int(run), C.c_int(left))
                got = o.collect(seq, offs, opt, nthreads=4)
                tot_walks += o.lib.orc_get_unique_walks()
                o.lib.orc_set_unique_walk_tables(None, None, None, C.c_uint64(0), C.c_int(0), C.c_int(0))
                for k in ("intv", "read_off", "step", "n_steps", "last_start"):
                    assert np.array_equal(got[k], want[k]), (trial, rl, err, run, left, k)
    print("trial", trial, "ok", n_ref, tot_walks, flush=True)
print("all exact; walks", tot_walks)
