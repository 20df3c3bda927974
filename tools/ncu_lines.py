"""Aggregate an ncu source page (cuda,sass) by source line / by opcode.  usage: ncu_lines.py rep.ncu-rep [units]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
import os
out = subprocess.run(["ncu", "-i", rep] + os.environ.get("NCU_ARGS", "").split() + ["--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}; byop = {}; stall_tot = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) <= ie: continue
    try: s = int(r[si]); n = int(r[ie])
    except ValueError: continue
    if r[2] == '-':
        agg[(cur, int(r[0]))] = (s, n, r[1])
    else:
        t = r[3].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
        byop.setdefault(op, [0, 0]); byop[op][0] += s; byop[op][1] += n
        for i, h in stall: stall_tot[h] = stall_tot.get(h, 0) + int(r[i] or 0)
tot = sum(v[0] for v in agg.values()); totn = sum(v[1] for v in agg.values())
print('total samples', tot, 'warp-instr', totn, 'per unit', totn / units)
T = sum(stall_tot.values()) or 1
print({k: round(100 * v / T, 1) for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]) if v * 100 > T})
S = sum(v[0] for v in byop.values()) or 1; N = sum(v[1] for v in byop.values()) or 1
for op, (s, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"  {op:10s} samples {100*s/S:5.1f}%  instr {100*n/N:5.1f}%  ({n/units:9.1f}/unit)")
for (f, l), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{f}:{l:4d} {100*s/tot:5.1f}% samp {100*n/totn:5.1f}% instr  {src.strip()[:100]}")
import os
if os.environ.get("BUCKETS"):
    print("--- buckets")
    for spec in os.environ["BUCKETS"].split(";"):
        f, a, b, name = spec.split(",")
        s = sum(v[0] for (ff, l), v in agg.items() if ff.startswith(f) and int(a) <= l <= int(b))
        n = sum(v[1] for (ff, l), v in agg.items() if ff.startswith(f) and int(a) <= l <= int(b))
        print(f"{name:28s} {100*s/tot:5.1f}% samples {100*n/totn:5.1f}% instr  {n/units:9.1f} instr/unit")
