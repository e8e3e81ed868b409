#!/usr/bin/env python3
"""CPU: the decoder side over a random grid of options, sizes and seeds (companion of tools/host_sweep.py).

Per case a random size / option set / seed; two encodes with the instrumented reference builds and three checks:
  A  oracle/_ref/x264_dump, no embedding       : `--parse-mv` returns every macroblock's type / partitioning / references / vectors
  B  x264_dump_conformant, embedding on        : same, both passes, forced decisions and flips in
  C  same stream                               : the stego vector read from the stream equals the embedder's ('EMBD'; all zero where
                                                 the embedder gave up), and the payload of every frame whose message is at least as
                                                 long as the code's constraint height equals the embedded message
    python tools/bitstream_sweep.py [seed=1] [cases=60]      one line per case; exit code 1 if any check fails"""
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader  # noqa: E402
import test_bitstream as tb  # noqa: E402


def main():
    rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    pcamv = pcamv_loader.load()
    bad = tot_mb = tot_frames = tot_bits = short = gave_up = trellis_gave_up = 0
    for case in range(n):
        w, h = rnd.choice([(176, 144), (352, 288), (64, 48), (96, 64), (48, 32), (320, 240), (16, 16), (128, 16), (16, 96), (640, 368), (720, 480)])
        me = rnd.choice(["dia", "hex", "umh"])
        subme = rnd.choice([1, 2, 3, 4, 5])
        ref = rnd.choice([1, 1, 2, 3, 4])
        parts = rnd.choice(["", "--partitions all", "--partitions p8x8,p4x4", "--partitions none"])
        extra = rnd.choice(["", "", "--no-cabac", "--no-cabac", "--no-fast-pskip", "--no-dct-decimate", "--no-chroma-me", "--mvrange 24", "--keyint 3 --min-keyint 3", "--nf"])
        qp = rnd.choice([4, 12, 18, 26, 32, 38, 44, 50])
        em = rnd.choice(["0.1", "0.2", "0.3", "0.7", "20", "0.04"])
        noise = rnd.choice([0, 2, 8, 32, 64])
        frames = rnd.choice([3, 4, 6])
        args = "--qp %d --ref %d --keyint 250 --me %s --subme %d %s %s" % (qp, ref, me, subme, parts, extra)
        msg = []
        before = trellis_gave_up
        with tempfile.TemporaryDirectory() as wd:
            try:
                stream, dump = tb.encode(pcamv, "x264_dump", (w, h), frames, args + " --emrate 0", noise, 300 + case, wd)
                pics = tb.parse_mv(stream, wd)
                b, t = tb.count_vector_mismatches(pics, dump)
                tot_mb += t
                if b or len(pics) == 0:
                    msg.append("A: %d / %d macroblocks differ" % (b, t))
                stream, dump = tb.encode(pcamv, "x264_dump_conformant", (w, h), frames, args + " --emrate " + em, noise, 300 + case, wd)
                pics = tb.parse_mv(stream, wd)
                b, t = tb.count_vector_mismatches(pics, dump)
                tot_mb += t
                if b:
                    msg.append("B: %d / %d macroblocks differ" % (b, t))
                messages, stegos = tb.extract_264(stream, em, wd)
                embeds = dump.embeds()
                if not (len(messages) == len(stegos) == len(embeds)):
                    msg.append("C: %d frames extracted, %d embedded" % (len(messages), len(embeds)))
                for e, (_, an, m), (_, ns, _, s) in zip(embeds, messages, stegos):
                    tot_frames += 1
                    # where the embedder gives up (message longer than the cover, empty message, syndrome out of range) its stego
                    # vector stays zeroed and pass 2 still flips every carrier whose cover bit is 1 (encoder/encoder.c:1848-1855)
                    failed = e["length"] > 0 and not e["stego"].any() and e["cover"].any()
                    if ns != e["length"] or not np.array_equal(s, e["stego"]):
                        msg.append("C: frame %d stego vector differs" % e["frame"])
                    elif failed or e["an"] < 1 or e["an"] > e["length"]:
                        gave_up += 1
                        # the decoder side recognises such a frame (every carrier even) and reports it as carrying nothing
                        if an != 0 and ns >= 16:
                            msg.append("C: frame %d: the embedder gave up, the extractor returned %d bits" % (e["frame"], an))
                        trellis_gave_up += failed and 1 <= e["an"] <= e["length"] and ns >= 16
                    elif e["an"] < 10:
                        short += 1
                    elif an != e["an"] or not np.array_equal(m, e["message"][:an]):
                        msg.append("C: frame %d payload differs" % e["frame"])
                    else:
                        tot_bits += an
            except AssertionError as ex:
                msg.append("ERROR " + str(ex).strip().splitlines()[-1][:200])
        bad += bool(msg)
        print("%s | %dx%d x %d noise %d seed %d | %s --emrate %s%s%s" % ("OK  " if not msg else "FAIL", w, h, frames, noise, 300 + case, " ".join(args.split()), em,
                                                                  " | trellis gave up on %d frame(s)" % (trellis_gave_up - before) if trellis_gave_up > before else "",
                                                                  "" if not msg else " | " + "; ".join(msg[:3])), flush=True)
    print("# %d cases, %d failed; %d macroblocks compared, %d embedded frames: %d payload bits recovered exactly, %d frames with messages shorter than the constraint height, %d frames the embedder gave up on (%d of them inside the trellis, all recognised as empty by the extractor)"
          % (n, bad, tot_mb, tot_frames, tot_bits, short, gave_up, trellis_gave_up))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
