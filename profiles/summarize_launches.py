"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`) of
profiles/profile_stamp.py: keeps the LAST stamp (from its canvas pre-processing launch to its composite launch), prints a
per-kernel table (markdown) and, when the DRAM metrics are present, writes the mean DRAM bytes per contraction launch that
bench.py reports as `roofline.traffic`.
    python profiles/summarize_launches.py gpurun_out/launches_r1_final.csv --md profiles/launches_r1_summary.md \
        --traffic profiles/traffic_r1.json --resolution 512 --denoise-steps 20 --batch 1"""
import argparse
import collections
import csv
import json
import re
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--md")
ap.add_argument("--traffic")
ap.add_argument("--resolution", type=int, default=512)
ap.add_argument("--denoise-steps", type=int, default=20)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--title", default="ncu launch list, one stamp")
ap.add_argument("--command", default="")
a = ap.parse_args()

rows = {}
order = []
with open(a.csv, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in rows:
        rows[i] = {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    rows[i][m] = v


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^dtp::", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    if m:
        base = m.group(1).split("::")[-1]
        return base + (m.group(2) or "")
    return name[:60]


start = max(i for i in order if "dilate_rows_kernel" in rows[i]["name"] or "canvas" in rows[i]["name"].lower())
# the canvas pre-processing is a handful of launches; back up to the first of its run
while start - 1 in rows and ("dilate" in rows[start - 1]["name"] or "canvas" in rows[start - 1]["name"].lower()):
    start -= 1
end = max(i for i in order if "composite_kernel" in rows[i]["name"])
sel = [rows[i] for i in order if start <= i <= end]
agg = collections.OrderedDict()
for r in sel:
    k = short(r["name"])
    d = agg.setdefault(k, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    d["n"] += 1
    d["us"] += r.get("gpu__time_duration.sum", 0.0)
    d["rd"] += r.get("dram__bytes_read.sum", 0.0)
    d["wr"] += r.get("dram__bytes_write.sum", 0.0)
tot = sum(d["us"] for d in agg.values())
has_dram = any(d["rd"] > 0 for d in agg.values())
out = []
out.append("# %s\n" % a.title)
if a.command:
    out.append("Command (on the B200 box): `%s`\n" % a.command)
out.append("Per-launch times under ncu are cold-cache and serialised (short kernels are inflated by several microseconds); "
           "compare SHARES.")
out.append("Last (warm) stamp only: %d launches, %.1f ms summed kernel time.\n" % (len(sel), tot / 1e3))
hdr = "| kernel | launches | total ms | avg us | share |" + (" DRAM read MB | DRAM write MB | avg DRAM B/launch |" if has_dram else "")
out.append(hdr)
out.append("|---|---:|---:|---:|---:|" + ("---:|---:|---:|" if has_dram else ""))
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    line = "| `%s` | %d | %.2f | %.1f | %.1f%% |" % (k, d["n"], d["us"] / 1e3, d["us"] / d["n"], 100 * d["us"] / tot)
    if has_dram:
        line += " %.1f | %.1f | %.0f |" % (d["rd"] / 1e6, d["wr"] / 1e6, (d["rd"] + d["wr"]) / d["n"])
    out.append(line)
gem = [d for k, d in agg.items() if k.startswith("gemm_tc_kernel")]
gn, gus = sum(d["n"] for d in gem), sum(d["us"] for d in gem)
if gem:
    out.append("\nContraction kernel (`gemm_tc_kernel`, all instantiations): %d launches, %.1f ms, share %.1f%%"
               % (gn, gus / 1e3, 100 * gus / tot)
               + (", mean DRAM traffic %.0f bytes per launch (%.1f GB per stamp)."
                  % (sum(d["rd"] + d["wr"] for d in gem) / gn, sum(d["rd"] + d["wr"] for d in gem) / 1e9) if has_dram else "."))
text = "\n".join(out) + "\n"
if a.md:
    open(a.md, "w").write(text)
else:
    sys.stdout.write(text)
if a.traffic and has_dram and gem:
    json.dump({"resolution": a.resolution, "denoise_steps": a.denoise_steps, "batch": a.batch,
               "kernel": "gemm_tc_kernel", "launches": gn,
               "avg_dram_bytes_per_launch": sum(d["rd"] + d["wr"] for d in gem) / gn,
               "dram_bytes_per_stamp": sum(d["rd"] + d["wr"] for d in gem),
               "share_of_summed_kernel_time": gus / tot,
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over "
                         "profiles/profile_stamp.py --no-op-profile --no-graph (last stamp)"},
              open(a.traffic, "w"), indent=1)
