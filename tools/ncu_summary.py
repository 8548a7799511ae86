"""Summarise an `ncu --page raw --csv` export: the handful of metrics DESIGN.md / profiles/ quote."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__occupancy_limit_blocks', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_global_ld.sum']
STALL = 'smsp__warp_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')][:60], r[hdr.index('Block Size')], r[hdr.index('Grid Size')])
        for w in WANT:
            if w in hdr:
                print('  %-72s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith('_per_warp_active.pct') and r[i]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print('  stall %-66s %.1f %%' % (h[len(STALL):-len('_per_warp_active.pct')], v))


if __name__ == '__main__':
    main(sys.argv[1])
