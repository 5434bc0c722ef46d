set -x
timeout 1500 python -m pytest tests/test_gpu_sector_kernels.py -m gpu -x -q > gpurun_out/r2_pytest_kt.log 2>&1
echo pytest rc=$?
tail -15 gpurun_out/r2_pytest_kt.log
for lv in 14 12 0; do
timeout 600 python bench.py --no-extras --skip-cpu --steps 10 --kmer-levels $lv > gpurun_out/kt_$lv.json 2> gpurun_out/kt_$lv.err
echo bench rc=$?
grep "prefix table" gpurun_out/kt_$lv.err
python -c "
import json; d=json.load(open('gpurun_out/kt_$lv.json')); print('levels $lv: %.1f M reads/s, seed kernel %.2f ms e2e %.1f' % (d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6), d['parity']['bit_exact'], d['prefix_table'])"
done
timeout 600 python bench.py --no-extras --skip-cpu --steps 10 --no-kmer-table > gpurun_out/kt_none.json 2> gpurun_out/kt_none.err
python -c "
import json; d=json.load(open('gpurun_out/kt_none.json')); print('no table: %.1f M reads/s, seed kernel %.2f ms' % (d['value']/1e6, d['roofline']['kernel_ms']))"
