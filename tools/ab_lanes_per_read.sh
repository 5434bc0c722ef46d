# A/B of the kernel variants of DESIGN.md section 11 on one box: bash tools/ab_lanes_per_read.sh  (-> gpurun_out/ab_lpr{1,2,3}.json)
for v in 2 3 1; do
  timeout 600 python bench.py --no-extras --skip-cpu --steps 10 --set lanes_per_read=$v > gpurun_out/ab_lpr$v.json 2> gpurun_out/ab_lpr$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_lpr$v.json')); print('lanes_per_read=$v: %.1f M reads/s, seed kernel %.2f ms' % (d['value']/1e6, d['roofline']['kernel_ms']), d['parity']['bit_exact'])"
done
