"""Stall samples of line_kernel by phase (split at BAR.SYNC):  python ncu_phase.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {k: i for i, k in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in body)
cuts = [0] + [i + 1 for i, r in enumerate(body) if 'BAR.SYNC' in r[idx['Source']]] + [len(body)]
nwarp = max(int(r[idx['Instructions Executed']]) for r in body[:5]) or 1
cols = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
for lo, hi in zip(cuts[:-1], cuts[1:]):
    s = sum(int(r[idx['# Samples']]) for r in body[lo:hi])
    ex = sum(int(r[idx['Instructions Executed']]) for r in body[lo:hi])
    agg = sorted(((sum(int(r[idx[c]] or 0) for r in body[lo:hi]), c[6:]) for c in cols), reverse=True)[:5]
    print(f"[{lo:5d},{hi:5d}) samples {100*s/tot:5.1f}%  warp-instr {ex:12d}  {agg}")
