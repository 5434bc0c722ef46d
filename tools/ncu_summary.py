"""Print the metrics we track from an `ncu --page raw --csv` dump (see B200_PROFILING.md)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_lookup_miss.sum',
        'lts__t_sectors_srcunit_tex_lookup_hit.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__sectors_read.sum', 'sm__cycles_elapsed.max',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {u.get(k, ''):16s} {d[k]}")
    print("-- warp stall reasons (cycles per issued instruction)")
    st = [(float(d[h]), h) for h in hdr if 'issue_stalled' in h and h.endswith('per_warp_active.pct') and d[h]]
    if not st:
        st = [(float(d[h]), h) for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h and d[h]]
    for v, h in sorted(st, reverse=True)[:10]:
        print(f"   {h:90s} {v:.3f}")
    print()
