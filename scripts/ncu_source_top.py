"""Top stall sites from `ncu --page source --csv` (SASS view). usage: ncu_source_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
key = "Warp Stall Sampling (All Samples)"
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[hi + 1:]):
    try: v = float(r[idx[key]])
    except Exception: continue
    data.append((v, n, r))
tot = sum(v for v, _, _ in data)
print(f"total samples {tot:.0f}")
for v, n, r in sorted(data, key=lambda x: -x[0])[:N]:
    top = sorted(((float(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{v:7.0f} {100 * v / tot:5.1f}%  #{n:5d} {r[idx['Source']].strip()[:90]:90s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
