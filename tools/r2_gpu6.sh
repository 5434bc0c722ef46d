set -x
python - <<'PY'
import importlib, os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dropin_bench as db
sy = importlib.import_module("bwa-mem-harp2_b200.synth")
tmp = tempfile.mkdtemp(dir="/dev/shm")
ref = sy.make_reference(300_000, 42); refn = ref.numpy()
fa = os.path.join(tmp, "g.fa"); sy.write_fasta(fa, [("chrA", refn)])
subprocess.run([db.REF, "index", fa], check=True, capture_output=True)
reads = sy.simulate_reads(ref, 3000, 101, 0.02, seed=7, n_frac=0.05).numpy()
sy.write_fastq(os.path.join(tmp, "r.fq"), reads)
cpu = subprocess.run([db.REF, "mem", "-t", "2", "-b", "1", fa, os.path.join(tmp, "r.fq")], capture_output=True, text=True)
for k, extra in enumerate([{"SMEM_GPU_ADAPTER_PRELOAD": "0"}, {"SMEM_GPU_ADAPTER_PRELOAD": "0"}, {}, {}, {"CUDA_MODULE_LOADING": "LAZY"}, {"CUDA_MODULE_LOADING": "LAZY"}]):
    env = dict(os.environ, SMEM_GPU_ADAPTER_STATS="1", SMEM_GPU_ADAPTER_TRACE="1", SMEM_GPU_TRACE="1", **extra)
    p = subprocess.run([db.GPU, "mem", "-t", "2", "-b", "64", fa, os.path.join(tmp, "r.fq")], capture_output=True, text=True, env=env)
    same = [l for l in cpu.stdout.splitlines() if not l.startswith("@PG")] == [l for l in p.stdout.splitlines() if not l.startswith("@PG")]
    print(f"==== run {k} {extra}: rc {p.returncode} same {same}")
    print("\n".join([l for l in p.stderr.splitlines() if "M::" not in l][:14])[:4000])
PY
