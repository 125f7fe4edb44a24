"""Opcode histogram per kernel of the in-tree library (cuobjdump -sass), written to profiles/sass_r2.md: the evidence that the
contraction and attention kernels are tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM /
STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and not mma.sync (HMMA).   python profiles/sass_histogram.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusiontexturepainting_b200", "libdtp_sm100.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "MUFU.EX2",
        "MUFU.RCP", "FFMA", "FMNMX", "F2FP", "STS", "LDS", "STG", "LDG", "BAR", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            hist[kern]["UTCHMMA.2CTA"] += 1
with open(os.path.join(ROOT, "profiles", "sass_r2.md"), "w") as f:
    f.write("# SASS opcode histogram per kernel (cuobjdump -sass libdtp_sm100.so, sm_100a)\n\n")
    f.write("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, "
            "LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none), "
            "STL / LDL = local-memory spills.\n\n")
    cols = [k for k in KEYS if any(h[k] for h in hist.values())]
    f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols) + "\n")
    for k, h in hist.items():
        if h["_total"] < 50:
            continue
        f.write(f"| `{k[:90]}` | {h['_total']} | " + " | ".join(str(h[c]) if h[c] else "" for c in cols) + " |\n")
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    f.write("\nTotals: " + ", ".join(f"{c} {tot[c]}" for c in cols) + "\n")
print(open(os.path.join(ROOT, "profiles", "sass_r2.md")).read()[:3000])
