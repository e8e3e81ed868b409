#!/usr/bin/env python3
"""Per-device-function breakdown of an ncu source page: executed instructions, stall samples and their reasons.
usage: ncu_funcs.py REPORT.ncu-rep KERNEL_SUBSTR [LIB.so]"""
import bisect
import csv
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "video-steganography-pcamv_b200/libpcamv_cuda.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first section only (SASS view of the first matching kernel)
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern in r[1])
hdr = rows[start + 1]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    data.append(r)
base = int(data[0][0], 16)
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
syms = []
for l in elf.splitlines():
    m = re.match(r"\s+0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+\S+\s+(\S+)", l)
    # symbols of device functions are "$<kernel>$<function>"; an exact kernel mangled name can be given as 4th argument so
    # that other template instantiations of the same kernel are not mixed in
    if m and (sys.argv[4] in m.group(3) if len(sys.argv) > 4 else kern in m.group(3)):
        short = re.sub(r"_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "", m.group(3).split("$")[-1])
        syms.append((int(m.group(1), 16), int(m.group(2), 16), short[:40]))
syms.sort()
offs = [s[0] for s in syms]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for d in data:
    a = int(d[0], 16) - base
    i = bisect.bisect_right(offs, a) - 1
    name = syms[i][2] if i >= 0 and a < syms[i][0] + syms[i][1] else "(kernel body)"
    x = agg.setdefault(name, {"inst": 0, "samp": 0, "static": 0, **{r: 0 for r in reasons}})
    x["inst"] += int(d[col["Instructions Executed"]]); x["samp"] += int(d[col["# Samples"]]); x["static"] += 1
    for r in reasons:
        x[r] += int(d[col[r]])
ti = sum(v["inst"] for v in agg.values()); ts = sum(v["samp"] for v in agg.values())
print("total warp-instructions %d, samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
    top = sorted(reasons, key=lambda r: -v[r])[:4]
    print("%-42s inst %5.1f%%  samples %5.1f%%  static %5d  | %s" % (
        k, 100.0 * v["inst"] / ti, 100.0 * v["samp"] / ts, v["static"],
        "  ".join("%s %.0f%%" % (r[6:], 100.0 * v[r] / max(v["samp"], 1)) for r in top)))
if len(sys.argv) > 4:
    # top SASS lines by samples inside a function
    want = sys.argv[4]
    lines = []
    for d in data:
        a = int(d[0], 16) - base
        i = bisect.bisect_right(offs, a) - 1
        name = syms[i][2] if i >= 0 and a < syms[i][0] + syms[i][1] else "(kernel body)"
        if want in name:
            lines.append((int(d[col["# Samples"]]), d[1].strip(), {r[6:]: int(d[col[r]]) for r in reasons if int(d[col[r]])}))
    for s, src, rs in sorted(lines, key=lambda x: -x[0])[:40]:
        print("%7d  %-60s %s" % (s, src[:60], rs))
