"""Print the key metrics of an .ncu-rep (ncu -i ... --page raw --csv) per kernel."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum']
stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')])
    for k in want:
        if k in hdr:
            print('   %-62s %s %s' % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
    st = sorted(((float(r[hdr.index(h)]), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print('   stall %6.1f%%  %s' % (v, h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')))
