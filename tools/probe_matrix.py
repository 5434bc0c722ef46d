"""Random-access roofline exploration: load flavour x block size x L2 fetch granularity (GB/s)."""
import importlib, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
nbytes = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
words = torch.randint(-2**31, 2**31 - 1, (nbytes // 4,), dtype=torch.int32, device="cuda")
class Ix: pass
ix = Ix(); ix.primary = 1; ix.L2 = [0, 1, 2, 3, 4]; ix.seq_len = 2 * nbytes; ix.bwt_size = nbytes // 4; ix.bwt = words
g = sg.SmemGpu(1024, 101)
g.upload_index(ix)
print("default l2_fetch_granularity", g.get_param("l2_fetch_granularity"))
res = {}
for gran in (0,):
    if gran:
        g.set_param("l2_fetch_granularity", gran)
    for var in range(5):
        g.set_param("probe_variant", var)
        for bb in (32, 64, 128):
            v = g.gather_roofline(bb, 0, 1024, 600)
            res[f"gran{gran}_var{var}_{bb}B"] = round(v, 1)
    print(gran, {k: v for k, v in res.items() if k.startswith(f"gran{gran}_")}, flush=True)
names = {10: "2x32B=64B", 11: "4x16B=64B", 12: "4x32B=128B", 13: "8x16B=128B", 14: "2x16B=32B", 15: "8x32B=256B"}
unit = {10: 64, 11: 64, 12: 128, 13: 128, 14: 32, 15: 256}
for var in (10, 11, 12, 13, 14, 15):
    g.set_param("probe_variant", var)
    for threads in (512, 1024, 2048):
        v = g.gather_roofline(64, 0, threads, 600)
        res[f"coop_{names[var]}_thr{threads}"] = round(v, 1)
        print("coop", names[var], "threads/SM", threads, "GB/s", round(v, 1), "G units/s", round(v / unit[var], 2), flush=True)
g.set_param("probe_variant", 0)
for chains in (256, 512, 1024, 2048):
    print("chains", chains, round(g.gather_roofline(64, 0, chains, 600), 1))
json.dump(res, open("gpurun_out/probe_matrix.json", "w"))
