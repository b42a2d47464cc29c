"""Print the handful of ncu metrics we read from a .ncu-rep (ncu -i <rep> --page raw --csv).  Usage: python scripts/ncu_summary.py rep [rep...]"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'smsp__inst_executed.sum']
STALL = 'smsp__average_warps_issue_stalled_'

for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', rep)
        d = dict(zip(hdr, zip(vals, units)))
        for w in WANT:
            if w in d:
                print('  %-78s %s %s' % (w, d[w][0], d[w][1]))
        st = sorted(((float(v[0]), h) for h, v in d.items() if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and v[0]), reverse=True)
        for v, h in st[:6]:
            print('  stall %-72s %.2f' % (h[len(STALL):-len('_per_issue_active.ratio')], v))
