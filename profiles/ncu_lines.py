#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples of the first kernel in an ncu report.
Usage: python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [inst|samples]"""
import csv, sys, subprocess
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
inst = defaultdict(float); samp = defaultdict(float); text = {}
fname, hdr, ix = '?', None, {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; hdr = None
        continue
    if 'Instructions Executed' in r:
        hdr = r; ix = {}
        for i, h in enumerate(hdr): ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    key = (fname, int(r[0]))
    try:
        inst[key] += float(r[ix['Instructions Executed']] or 0); samp[key] += float(r[ix['# Samples']] or 0)
    except ValueError:
        continue
    text.setdefault(key, r[1])
ti, ts = sum(inst.values()), sum(samp.values())
print('total warp instructions %.0f, samples %.0f' % (ti, ts))
order = 'samples' if len(sys.argv) > 3 and sys.argv[3] == 'samples' else 'inst'
for k in sorted(inst, key=lambda k: -(samp[k] if order == 'samples' else inst[k]))[:top]:
    print('%-22s %5d  inst %5.1f%%  samples %5.1f%%  %s' % (k[0][:22], k[1], 100 * inst[k] / ti, 100 * samp[k] / ts, text[k][:100]))
