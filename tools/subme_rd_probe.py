#!/usr/bin/env python3
"""What --subme 6 / 7 (RD mode decision, refused by the GPU host) would change, measured on the reference itself (CPU only):
identical output for 6 and 7 (the RD refinement of 7 is compiled out, encoder/analyse.c:3112 `if(0 && ...)`), encode time,
stream size, PSNR and carrier count against --subme 5 at the same QP.   python tools/subme_rd_probe.py > profiles/r02_subme_rd_measurement.txt"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader  # noqa: E402
import refrun  # noqa: E402


def main():
    pcamv = pcamv_loader.load()
    exe = os.path.join(ROOT, "oracle", "_ref", "x264_wide")
    print("# clip | options | seconds  bytes  PSNR-Y (mean)  MV carriers (thousands)  md5")
    for (w, h, frames, synth) in ((352, 288, 30, 1), (1920, 1080, 8, 2)):
        with tempfile.TemporaryDirectory() as wd:
            clip = refrun.synth_clip(pcamv, w, h, frames, config=synth, stream=1, workdir=wd)
            for cab in ("", "--no-cabac"):
                for sub, parts in ((5, ""), (6, ""), (7, ""), (5, "--partitions all")):
                    args = ("--qp 26 --ref 1 --keyint 250 --me umh --subme %d --emrate 0.2 %s %s" % (sub, cab, parts)).split()
                    out = os.path.join(wd, "o.264")
                    t0 = time.perf_counter()
                    p = subprocess.run([exe] + args + ["-o", out, clip, "%dx%d" % (w, h)], capture_output=True)
                    dt = time.perf_counter() - t0
                    err = p.stderr.decode("latin-1")
                    psnr = re.search(r"PSNR Mean Y:([0-9.]+) U:[0-9.]+ V:[0-9.]+ Avg:[0-9.]+ Global:[0-9.]+ kb/s", err)
                    mvs = re.search(r":([0-9.]+) K", err)          # "total MV carriers: N K" of the reference's closing lines (GB18030 text)
                    print("%dx%d x %d | %-60s | %6.2f s  %8d B  %s dB  %s  %s" % (w, h, frames, " ".join(args), dt, os.path.getsize(out),
                          psnr.group(1) if psnr else "?", mvs.group(1) if mvs else "?", hashlib.md5(open(out, "rb").read()).hexdigest()[:12]))
                    sys.stdout.flush()


if __name__ == "__main__":
    main()
