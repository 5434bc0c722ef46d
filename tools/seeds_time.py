import sys, time, importlib, numpy as np, ctypes as C, torch
sys.path.insert(0, "/root/repo")
sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu"); fm = importlib.import_module("bwa-mem-harp2_b200.fmindex"); sy = importlib.import_module("bwa-mem-harp2_b200.synth")
dev = torch.device("cuda", 0)
fwd = sy.make_reference(3_100_000_000, 13, dev); ix = fm.build_index(fwd, sa_intv=32)
reads = sy.simulate_reads(fwd, 2_000_000, 101, 0.01, seed=1000, paired=True); seq, offs = sy.to_batch(reads)
pac = sg.pack_pac(fwd); del fwd, reads; torch.cuda.empty_cache()
g = sg.SmemGpu(max_batch_reads=2_000_000, max_read_len=101, devices=[0]); g.upload_index(ix); g.upload_sa(ix); g.build_repeat_filter((pac, 3_100_000_000))
g.stage(seq, offs); g.run_collect()
lib = g.lib; n = 2_000_000
seed_off = np.zeros(n + 1, np.int64); tot = C.c_int64(0)
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    rc = lib.smem_gpu_seeds(g.h, C.c_int(19), C.c_int64(10000), None, C.c_int64(0), seed_off.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(tot))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ch = g.chains(n, ix.seq_len // 2, fetch=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("seeds (resident, no copy of seeds)", round((t1 - t) * 1e3, 2), "ms rc", rc, "n", tot.value, "| chains (no fetch)", round((t2 - t1) * 1e3, 2), "ms", ch["n_chains"])
