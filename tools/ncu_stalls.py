"""Rank the CUDA source lines of an .ncu-rep by warp-stall samples and list the stall reasons.
usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = cur = None
stall, inst, text = defaultdict(float), defaultdict(float), {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and cur:
        try:
            ln = int(r[0]); s = float(r[hdr.index('# Samples')]); v = float(r[hdr.index('Instructions Executed')])
        except Exception:
            continue
        stall[(cur, ln)] += s; inst[(cur, ln)] += v; text[(cur, ln)] = r[1]
ts, ti = sum(stall.values()) or 1, sum(inst.values()) or 1
print('%d stall samples, %.0f warp-instructions' % (ts, ti))
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:top]:
    print('%5.1f%% inst %4.1f%%  %s:%d  %s' % (v / ts * 100, inst[k] / ti * 100, k[0], k[1], text[k].strip()[:100]))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, r = rows[0], rows[2]
d = [(float(v.replace(',', '')), k) for k, v in zip(h, r)
     if 'pcsamp_warps_issue_stalled' in k and not k.endswith('not_issued')]
print('  '.join('%s %.0f' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), v) for v, k in sorted(d, reverse=True)[:8]))
