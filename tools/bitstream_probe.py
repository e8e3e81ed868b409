#!/usr/bin/env python3
"""How much of the payload survives in the .264 itself: the unmodified reference against the conformance switch.

For a list of configurations: encode with oracle/_ref/x264_dump (the reference) and x264_dump_conformant (the three pass-2
statements corrected, tools/reftree.py::conformance_switch), read both streams back with `x264_pcamv --parse-mv` /
`--extract-264` (host/pcamv_bitstream.c) and compare with what the encoder recorded: vectors of the written pass ('MBAN'),
stego vector and message ('EMBD').  CPU only.  Output: one line per configuration and binary -> profiles/r02_bitstream_extraction.txt"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader  # noqa: E402
import test_bitstream as tb  # noqa: E402

CASES = [
    ((352, 288), 10, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", "0.2", 32),
    ((352, 288), 10, "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --no-cabac", "0.2", 32),
    ((352, 288), 10, "--qp 26 --ref 3 --keyint 250 --me umh --subme 5", "0.2", 32),
    ((352, 288), 10, "--qp 20 --ref 1 --keyint 250 --me hex --subme 5 --partitions all", "0.1", 32),
    ((352, 288), 10, "--qp 32 --ref 1 --keyint 250 --me hex --subme 5", "0.2", 8),
    ((352, 288), 10, "--qp 26 --ref 1 --keyint 250 --me hex --subme 5", "60", 32),
    ((1280, 720), 6, "--qp 26 --ref 1 --keyint 250 --me umh --subme 5", "0.2", 32),
    ((1920, 1080), 4, "--qp 26 --ref 1 --keyint 250 --me umh --subme 5", "0.2", 32),
]


def main():
    pcamv = pcamv_loader.load()
    print("# size frames options emrate | binary: P pictures, macroblocks whose parsed vectors differ from the encoder's / all, "
          "carriers, stego bits read wrong, frames whose payload is recovered / frames")
    for k, (size, frames, args, emrate, noise) in enumerate(CASES):
        for binary in ("x264_dump", "x264_dump_conformant"):
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as wd:
                try:
                    stream, dump = tb.encode(pcamv, binary, size, frames, args + " --emrate " + emrate, noise, 20 + k, wd)
                    pics = tb.parse_mv(stream, wd)
                    bad, total = tb.count_vector_mismatches(pics, dump)
                    messages, stegos = tb.extract_264(stream, emrate, wd)
                except AssertionError as e:
                    print("%dx%d %d %s %s | %-20s: NOT READABLE (%s)" % (size[0], size[1], frames, args, emrate, binary, str(e).strip().splitlines()[-1][:160]))
                    continue
                embeds = dump.embeds()
                wrong = carriers = ok = 0
                for e, (_, an, msg), (_, n, _, stego) in zip(embeds, messages, stegos):
                    carriers += e["length"]
                    wrong += int(np.count_nonzero(stego != e["stego"])) if n == e["length"] else e["length"]
                    ok += an == e["an"] and np.array_equal(msg, e["message"][:an])
                print("%dx%d %d %s %s | %-20s: %d P pictures, %d / %d macroblocks differ, %d carriers, %d stego bits wrong, %d / %d frames recovered"
                      % (size[0], size[1], frames, args, emrate, binary, len(pics), bad, total, carriers, wrong, ok, len(embeds)))
                sys.stdout.flush()


if __name__ == "__main__":
    main()
