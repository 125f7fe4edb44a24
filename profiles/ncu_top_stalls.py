"""Top stall sites of an `ncu --set full --import-source on` capture, read on the CPU box:
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv ; python profiles/ncu_top_stalls.py /tmp/src.csv [N] [ctx]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
smp = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(smp(r) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print(rows[0][1][:100])
print("total samples", tot, "| by reason:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -smp(data[i]))[:n]
for i in order:
    r = data[i]
    st = sorted(((h, int(r[ix[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(f"[{i:5d}] {smp(r):6d} {100 * smp(r) / tot:5.1f}%  x{r[ix['Instructions Executed']]:>9}  {r[ix['Source']][:80]:80s} {st}")
    if ctx:
        for k in range(max(0, i - ctx), min(len(data), i + 3)):
            print(f"          {smp(data[k]):6d}  {data[k][ix['Source']][:100]}")
