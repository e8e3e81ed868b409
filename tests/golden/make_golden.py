#!/usr/bin/env python3
"""Regenerates the golden fixtures from the compiled reference (oracle/_ref/x264_dump).

Run in the build container (needs /root/reference to have been compiled by oracle/build_ref.py):
    python tests/golden/make_golden.py
Each fixture is the instrumented reference's dump (planes + every search call + per-MB decisions +
embed-stage vectors) of a short QCIF clip from synth/pcamv_synth.c, xz-compressed.
"""
import lzma
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader  # noqa: E402
import refrun  # noqa: E402

CASES = {
    # name: (w, h, frames, noise16, dump frame range, reference CLI args)
    "qcif_hex5": (176, 144, 4, 32, "1:4", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2"),
    "qcif_umh5_ref2": (176, 144, 4, 16, "2:4", "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.2"),
    "qcif_esa5": (176, 144, 3, 32, "1:2", "--qp 26 --ref 1 --keyint 250 --me esa --merange 16 --subme 5 --emrate 0.2"),
    "qcif_tesa5": (176, 144, 3, 32, "1:2", "--qp 26 --ref 1 --keyint 250 --me tesa --merange 16 --subme 5 --emrate 0.2"),
    "qcif_dia2_lownoise": (176, 144, 4, 4, "1:4", "--qp 30 --ref 1 --keyint 250 --me dia --subme 2 --emrate 0.2"),
}


def main():
    pcamv = pcamv_loader.load()
    work = tempfile.mkdtemp(prefix="pcamv_mkgold_")
    only = sys.argv[1:]
    for name, (w, h, n, noise, rng, args) in CASES.items():
        if only and name not in only:
            continue
        clip = refrun.synth_clip(pcamv, w, h, n, config=9, stream=len(name), noise16=noise, workdir=work)
        dump = os.path.join(work, name + ".bin")
        refrun.run_ref(clip, w, h, args.split(), dump=dump, frames=rng, count=True)
        raw = open(dump, "rb").read()
        with lzma.open(os.path.join(HERE, name + ".bin.xz"), "wb", preset=9) as f:
            f.write(raw)
        print(name, len(raw), "->", os.path.getsize(os.path.join(HERE, name + ".bin.xz")))


if __name__ == "__main__":
    main()
