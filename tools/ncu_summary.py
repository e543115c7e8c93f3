"""Summarise an .ncu-rep (read here, on the CPU box): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
st = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
for d in data:
    print('====', d[idx['Kernel Name']][:110])
    for k in keys:
        if k in idx:
            print('  %-70s %s %s' % (k, d[idx[k]], units[idx[k]]))
    vals = sorted(((float(d[idx[h]] or 0), h) for h in st), reverse=True)[:7]
    print('  stalls per issue: ' + ', '.join('%s %.2f' % (h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v) for v, h in vals))
