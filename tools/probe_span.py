"""Random-access request rate vs span (is the ceiling TLB reach, L2 capacity or DRAM?)."""
import importlib, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
nbytes = 6_400_000_000
words = torch.randint(-2**31, 2**31 - 1, (nbytes // 4,), dtype=torch.int32, device="cuda")
class Ix: pass
ix = Ix(); ix.primary = 1; ix.L2 = [0, 1, 2, 3, 4]; ix.seq_len = 2 * nbytes; ix.bwt_size = nbytes // 4; ix.bwt = words
g = sg.SmemGpu(1024, 101)
g.upload_index(ix)
res = {}
for var, name, unit in ((0, "1x(2x32B) per-lane 64B", 64), (10, "coop 2x32B=64B", 64), (12, "coop 4x32B=128B", 128)):
    g.set_param("probe_variant", var)
    for mb in (32, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3100, 4096, 6100):
        v = g.gather_roofline(64, mb * 1000 * 1000, 1024, 600)
        res[f"{name}@{mb}MB"] = v
        print(f"{name:26s} span {mb:5d} MB: {v:8.1f} GB/s  {v / unit:7.2f} G units/s", flush=True)
json.dump(res, open("gpurun_out/probe_span.json", "w"))
