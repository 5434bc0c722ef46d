L=$PWD/bwa-mem-harp2_b200/variants
run() { SMEM_GPU_LIB=$1 python bench.py --skip-cpu --no-extras --steps 5 --warmup 3 $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2: seed kernel %.3f ms, %.2f M reads/s, parity %s' % (d['roofline']['kernel_ms'], d['value']/1e6, d['parity']['bit_exact']), d.get('prefix_table'))"; }
run $L/libsmem_nokt.so "no KT code, no table" "--no-kmer-table"
run $L/libsmem_shipped.so "KT code, no table" "--no-kmer-table"
run $L/libsmem_shipped.so "KT code, table 14, unused" "--set kmer_table=0"
run $L/libsmem_shipped.so "KT code, table 14, levels used 4" "--set kmer_table_levels_used=4"
run $L/libsmem_shipped.so "KT code, table 14, levels used 8" "--set kmer_table_levels_used=8"
run $L/libsmem_shipped.so "KT code, table 14, levels used 12" "--set kmer_table_levels_used=12"
run $L/libsmem_shipped.so "KT code, table 14, all levels" ""
