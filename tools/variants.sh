# Build kernel variants next to the shipped library (bwa-mem-harp2_b200/variants/libsmem_<name>.so, git-ignored) and, on a GPU box,
# time each through bench.py (SMEM_GPU_LIB selects the library).  usage: bash tools/variants.sh build|run [bench flags]
set -e
cd "$(dirname "$0")/.."
B=-DSEED_KEEP_SP,-DSEED_KEEP_GP,-DSEED_NARROW        # the shipped switches (csrc/Makefile SEED_FLAGS)
V="${VARIANTS:-shipped:$B nosplitcopy:$B,-DSEED_NO_SPLIT_COPY}"
if [ "$1" = build ]; then
  mkdir -p bwa-mem-harp2_b200/variants
  for v in $V; do
    name=${v%%:*}; fl=$(echo "${v#*:}" | tr ',' ' ')
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr $fl -shared \
      -o bwa-mem-harp2_b200/variants/libsmem_$name.so bwa-mem-harp2_b200/csrc/smem_gpu.cu -lcudart
  done
else
  shift || true
  for v in $V; do
    name=${v%%:*}
    SMEM_GPU_LIB=$PWD/bwa-mem-harp2_b200/variants/libsmem_$name.so python bench.py --skip-cpu --no-extras --steps 5 --warmup 3 "$@" 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $name: seed kernel %.3f ms, step %.3f ms, %.2f M reads/s, e2e %.2f M, parity %s' % (d['roofline']['kernel_ms'], d['device_ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['parity']['bit_exact']))"
  done
fi
