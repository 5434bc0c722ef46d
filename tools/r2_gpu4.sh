set -x
# one traced drop-in run: 16 threads x 65536-read batches, 8 handles (the pathological row of the sweep)
python - <<'PY'
import importlib, os, subprocess, sys, time, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dropin_bench as db
import torch
fm = importlib.import_module("bwa-mem-harp2_b200.fmindex"); sy = importlib.import_module("bwa-mem-harp2_b200.synth")
tmp = tempfile.mkdtemp(dir="/dev/shm")
ref = sy.make_reference(100_000_000, 11, "cuda"); refn = ref.cpu().numpy()
fa = os.path.join(tmp, "g.fa"); db.write_fasta(fa, [("chr1", refn)])
subprocess.run([db.REF, "fa2pac", "-f", fa, fa], check=True, capture_output=True)
ix = fm.build_index(ref, sa_intv=32); ix.save(fa + ".bwt"); ix.save_sa(fa + ".sa")
reads = sy.simulate_reads(ref, 1_000_000, 101, 0.01, seed=21).cpu().numpy()
db.write_fastq(os.path.join(tmp, "r.fq"), reads)
for h, b in ((8, 65536), (4, 16384)):
    env = dict(os.environ, SMEM_GPU_ADAPTER_STATS="1", SMEM_GPU_ADAPTER_HANDLES=str(h), SMEM_GPU_ADAPTER_TRACE="1", SMEM_GPU_TRACE="1")
    t0 = time.time()
    p = subprocess.run([db.GPU, "mem", "-t", "16", "-b", str(b), fa, os.path.join(tmp, "r.fq")], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, text=True)
    print(f"==== handles {h} batch {b}: wall {time.time() - t0:.2f}s")
    print("\n".join(l for l in p.stderr.splitlines() if "trace" in l or "lists_from_cache" in l)[:12000])
PY
