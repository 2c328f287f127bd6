"""Summarise an .ncu-rep (read with ncu -i ... --page raw --csv) into the few numbers DESIGN/profiles quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print('===', name[:90])
    for k in keys:
        if k in hdr:
            print('  %-72s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print('  stall %-40s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    break
