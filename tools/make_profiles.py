"""Turn the outputs of tools/measure_round.sh (gpurun_out/) into the committed evidence under profiles/:
    rNN_bench_n1.json            the bench line
    rNN_launches.csv / _summary  ncu launch list of the same command and the per-kernel shares
    rNN_ncu_step_summary.txt     ncu --set full of the step's kernels: headline metrics plus the L1 / crossbar breakdown
    rNN_ncu_{fwd,bwd}_lines.txt  per-source-line instruction and stall tables
    rNN_ncu_splat_summary.txt    the S kernels
    traffic.json                 dram bytes per launch, read by bench.py for roofline.traffic
usage: python tools/make_profiles.py r02 [pixels_of_the_profiled_launch]"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else 64 * 1080 * 1920
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')

# ---- bench line
line = open(os.path.join(G, 'bench_final.json')).read().strip().splitlines()[-1]
bench = json.loads(line)
with open(os.path.join(P, tag + '_bench_n1.json'), 'w') as fh:
    json.dump(bench, fh, indent=1)

# ---- launch list
shutil.copy(os.path.join(G, 'launches_final.csv'), os.path.join(P, tag + '_launches.csv'))
rows = [r for r in csv.reader(open(os.path.join(G, 'launches_final.csv'))) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
t, c = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    t[r[ki]] += v
    c[r[ki]] += 1
tot = sum(t.values())
with open(os.path.join(P, tag + '_launches_summary.txt'), 'w') as fh:
    fh.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 5 --warmup 1 '
             '--no-cpu-baseline --no-e2e\n(launch list in %s_launches.csv; per-launch times are cold-cache and '
             'serialised -- the SHARE is what counts; the first 400 launches cover the headline step and the start of '
             'the other workloads)\n' % tag)
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        fh.write('%10.3f ms %5.1f%% %4dx %9.1f us  %s\n' % (v / 1000, 100 * v / tot, c[k], v / c[k], k[:110]))

# ---- ncu --set full summaries
HEAD = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        # the L1 data pipe and the L1 -> crossbar port: what bounds the gather / scatter kernels
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__inst_executed_op_global_red.sum', 'l1tex__cycles_elapsed.avg']


def raw_rows(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(raw.splitlines()))


def summary(rep, out, title):
    rows = raw_rows(rep)
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
    traffic = {}
    with open(out, 'w') as fh:
        fh.write(title + '\n')
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            fh.write('----- %s\n' % name)
            for k in HEAD:
                if k in hdr:
                    fh.write('   %-78s %s %s\n' % (k, r[hdr.index(k)], units[hdr.index(k)]))
            st = sorted(((float(r[hdr.index(h)]), h) for h in stall), reverse=True)[:6]
            for v, h in st:
                fh.write('   stall %6.1f%%  %s\n' % (v, h.replace('smsp__warp_issue_stalled_', '').replace(
                    '_per_warp_active.pct', '')))

            def gb(key):
                v, u = float(r[hdr.index(key)].replace(',', '')), units[hdr.index(key)]
                return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
            traffic[name] = int(gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'))
    return traffic


step_rep = os.path.join(G, 'prof_final_step.ncu-rep')
tr = summary(step_rep, os.path.join(P, tag + '_ncu_step_summary.txt'),
             'ncu --set full --clock-control none --import-source on, the kernels of the headline step at the bench size '
             '(64x3x1080x1920), python bench.py --steps 1 --warmup 1')
summary(os.path.join(G, 'prof_final_splat.ncu-rep'), os.path.join(P, tag + '_ncu_splat_summary.txt'),
        'ncu --set full --clock-control none --import-source on -k regex:splat, python tools/splat_probe.py 16 '
        '(16x3x1080x1920: Flow.apply s with masks fwd+bwd, then switch_ref)')
for pat, name in (('warp_t_fwd', 'fwd'), ('warp_t_bwd', 'bwd')):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), step_rep, str(npx), pat, '40'],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, '%s_ncu_%s_lines.txt' % (tag, name)), 'w') as fh:
        fh.write('ncu --set full --import-source on, bench workload 64x3x1080x1920, kernels matching %s '
                 '(instructions per 32 px, share, stall-sample share, source line)\n' % pat)
        fh.write(out)

tj = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, bench.py default workload '
                  '(64x3x1080x1920), %s; see %s_ncu_step_summary.txt.  ofb_warp_t_bwd is the kernel alone; the '
                  'cudaMemsetAsync of the payload gradient the call also owns adds 4*C bytes per pixel.' % (tag, tag)}
for k, v in tr.items():
    if 'warp_t_fwd' in k:
        tj['ofb_warp_t_fwd'] = v
    elif 'warp_t_bwd' in k:
        tj['ofb_warp_t_bwd'] = v
    elif 'flow_stats' in k:
        tj['ofb_flow_stats'] = tj['ofb_flow_stats_batch'] = v
with open(os.path.join(P, 'traffic.json'), 'w') as fh:
    json.dump(tj, fh, indent=1)
print('wrote profiles/%s_*; traffic:' % tag, {k: v for k, v in tj.items() if k != '_comment'})
