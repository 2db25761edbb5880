"""Extract the metrics we cite from an .ncu-rep into a small CSV.  usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
with open(sys.argv[2], 'w') as f:
    f.write('metric,unit,value\n')
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or h.startswith(k) for k in want) and not any(x in h for x in ('.max', '.min', '.sum.pct', 'dram__bytes_read.sum.p', 'dram__bytes_write.sum.p')):
            f.write('%s,%s,%s\n' % (h, u, v.replace(',', '')))
