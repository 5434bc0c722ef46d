# bench.py under torchrun on N GPUs of one box (the driver's launch line); usage: bash tools/r2_scale.sh N [extra bench flags]
N=$1; shift
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo bench rc=$?
grep -E "^\[bench\] (e2e|single|pcie)|Error|error|Traceback" gpurun_out/r2_bench_n$N.err | cut -c1-1500 | head -30
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n$N.json"))
print("N", d["n_gpus"], "value %.1f M" % (d["value"] / 1e6), "e2e %.1f M (%.2f ms/step)" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]))
for k in ("e2e_bwtintv", "e2e_seeds", "e2e_chains"):
    if k in d: print(k, "%.1f M" % (d[k]["value"] / 1e6))
for r in d["e2e"]["stages_per_rank"]: print(r)
print(d.get("pcie_probe")); print(d.get("single_handle_all_devices"))
PY
