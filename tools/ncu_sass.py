"""Developer tool: summarise an `ncu --page source --csv --print-source sass` dump: per instruction samples, executed
count, avg threads, and the top stall reason.  usage: ncu_sass.py src.csv [min_samples]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hdr = rows[1]
rows = rows[:2] + [r for r in rows[2:] if len(r) == len(hdr)][:0] + [r for r in rows[2:] if len(r) >= 40 and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in rows[2:])
minimum = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print("total samples", tot)
for n, r in enumerate(rows[2:]):
    s = int(r[ix["# Samples"]] or 0)
    if s < minimum:
        continue
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{n:4d} {s:6d} {100*s/tot:5.1f}% ex={int(r[ix['Instructions Executed']]):9d} thr={r[ix['Avg. Threads Executed']]:>5s} "
          f"{st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {r[ix['Source']].strip()}")
