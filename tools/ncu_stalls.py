"""Developer tool: total warp-stall samples by reason from an `ncu --page source --csv` dump of ONE kernel."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {c: 0 for c in cols}
for r in rows[2:]:
    if len(r) >= 40 and r[0].startswith("0x"):
        for c in cols:
            tot[c] += int(r[ix[c]] or 0)
s = sum(tot.values())
print("samples", s, " ".join(f"{c[6:]}={100*v/s:.0f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > s))
