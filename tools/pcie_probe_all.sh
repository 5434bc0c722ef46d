# aggregate host<->device copy bandwidth with all GPUs of the box copying at once (one process per GPU)
n=$(nvidia-smi -L | wc -l)
for i in $(seq 0 $((n-1))); do CUDA_VISIBLE_DEVICES=$i python tools/pcie_probe.py > /tmp/pcie_$i.log 2>&1 & done
wait
for i in $(seq 0 $((n-1))); do echo "gpu $i: $(tr '\n' ' ' < /tmp/pcie_$i.log)"; done
