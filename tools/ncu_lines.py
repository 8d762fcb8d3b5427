"""Aggregate an .ncu-rep's source page by CUDA source line: instructions per 32 pixels and stall samples."""
import csv, subprocess, sys
from collections import defaultdict
rep, npx = sys.argv[1], float(sys.argv[2])          # npx = pixels processed by the profiled launch
pat = sys.argv[3] if len(sys.argv) > 3 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = fn = hdr = None
agg, stall, text, ops = defaultdict(float), defaultdict(float), {}, defaultdict(float)
for r in rows:
    if len(r) == 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and cur and fn and pat in fn:
        try:
            ln = int(r[0]); v = float(r[hdr.index('Instructions Executed')]); s = float(r[hdr.index('# Samples')])
        except Exception:
            continue
        agg[(cur, ln)] += v; stall[(cur, ln)] += s; text[(cur, ln)] = r[1]
tot, ts = sum(agg.values()), sum(stall.values()) or 1
per = npx / 32
print('total warp-instructions %.0f = %.1f per 32 px; %d stall samples' % (tot, tot / per, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print('%6.1f %5.1f%% stall %5.1f%%  %s:%d  %s' % (v / per, v / tot * 100, stall[k] / ts * 100, k[0], k[1], text[k].strip()[:95]))
