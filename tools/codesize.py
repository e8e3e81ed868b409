#!/usr/bin/env python3
"""Per-device-function SASS sizes of libpcamv_cuda.so (the instruction cache is the scarce resource of the per-MB code)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "video-steganography-pcamv_b200/libpcamv_cuda.so"
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for l in out.splitlines():
    m = re.match(r"\s+0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+\S+\s+(\S+)", l)
    if m:
        name = m.group(3)
        kern = "AP" if "k_analyse_p" in name else "CT" if "k_cost_table" in name else "SB" if "k_search_batch" in name else "--"
        short = name.split("$")[-1]
        short = re.sub(r"_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "", short)
        rows.append((kern, int(m.group(2), 16), short[:80]))
for kern in ("AP", "CT", "SB"):
    r = sorted([x for x in rows if x[0] == kern], key=lambda x: -x[1])
    print(kern, "callee total", sum(x[1] for x in r))
    for x in r:
        print("   %7d  %s" % (x[1], x[2]))
