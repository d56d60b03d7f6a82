"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
STALLS = ['long_scoreboard', 'wait', 'short_scoreboard', 'branch_resolving', 'not_selected', 'math_pipe_throttle', 'barrier',
          'mio_throttle', 'lg_throttle', 'dispatch_stall', 'no_instruction']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        print('## ' + name)
        for k in KEYS:
            if k in hdr: print('- %s = %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for s in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if k in hdr: st.append('%s %.2f' % (s, float(r[hdr.index(k)])))
        print('- stall cycles per issued instruction: ' + ', '.join(st))
        print()
if __name__ == '__main__':
    main(sys.argv[1])
