#!/usr/bin/env python3
"""What the conformance switch costs in rate and distortion (CPU only): the unmodified reference (oracle/_ref/x264_dump) against the
same sources with the switch compiled in (x264_dump_conformant) on the same synthetic clips at the same QP - stream size,
encoder-side PSNR, carrier count, encode time.    python tools/conformant_cost_probe.py > profiles/r02_conformant_vs_default_cost.txt"""
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
    print("# clip | binary | bytes  PSNR-Y (encoder side)  MV carriers (thousands)  seconds")
    for (w, h, frames, synth) in ((352, 288, 30, 1), (1280, 720, 10, 5), (1920, 1080, 8, 2)):
        with tempfile.TemporaryDirectory() as wd:
            clip = refrun.synth_clip(pcamv, w, h, frames, config=synth, stream=1, workdir=wd)
            for parts in ("", "--partitions all"):
                for b in ("x264_dump", "x264_dump_conformant"):
                    out = os.path.join(wd, "o.264")
                    t0 = time.perf_counter()
                    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", b)] + ("--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2 %s" % parts).split()
                                       + ["-o", out, clip, "%dx%d" % (w, h)], capture_output=True)
                    dt = time.perf_counter() - t0
                    err = p.stderr.decode("latin-1")
                    ps = re.search(r"PSNR Mean Y:([0-9.]+) U:[0-9.]+ V:[0-9.]+ Avg:[0-9.]+ Global:[0-9.]+ kb/s", err)
                    mv = re.search(r":([0-9.]+) K", err)
                    print("%dx%d x %d %s | %-20s | %8d B  %s dB  %s  %.2f s" % (w, h, frames, parts, b, os.path.getsize(out), ps.group(1) if ps else "?",
                                                                             mv.group(1) if mv else "?", dt), flush=True)


if __name__ == "__main__":
    main()
