#!/usr/bin/env python3
"""Hostile input for the decoder side: host/pcamv_bitstream.c built with AddressSanitizer + UBSan and fed mutated streams.

The parser reads untrusted bytes (a .264 from anywhere), so every table index, array size and loop bound it derives from the
stream has to hold for ANY input.  This builds the file stand-alone (against oracle/_ref/libx264_wide.a for the reference's
tables; headers from /root/reference at build time, nothing copied), takes valid streams from the reference encoder, mutates
them (byte flips, truncations, splices, zero / 0xff runs) and runs `--parse-mv` / `--extract-264` on each: any sanitizer
report or signal fails.  CPU only; needs /root/reference.    python tools/bitstream_fuzz.py [cases=400] [seed=1]"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import reftree  # noqa: E402

HARNESS = r'''
#include <stdint.h>
#include <string.h>
int pcamv_bitstream_main( int argc, char **argv );
/* the syndrome extractor lives in the scratch copy of encoder/encoder.c (it needs the embedder's static getMatrix); the fuzz target
 * is the parser, so a stand-in that touches its arguments the same way is enough */
int pcamv_stc_extract( const uint8_t *stego, int n, uint8_t *message, int an, int matrixheight )
{
    int i; unsigned acc = 0;
    if( an <= 0 || n < an ) return -1;
    for( i = 0; i < n; i++ ) acc += stego[i];
    memset( message, acc & 1, an );
    return n;
}
int main( int argc, char **argv ) { return pcamv_bitstream_main( argc, argv ) ? 1 : 0; }
'''


def build(workdir):
    open(os.path.join(workdir, "config.h"), "w").write(reftree.CONFIG_H)
    open(os.path.join(workdir, "harness.c"), "w").write(HARNESS)
    exe = os.path.join(workdir, "bitstream_fuzz")
    cmd = (["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
           + [f for f in reftree.CFLAGS if f.startswith("-D") or f == "-w"] + ["-I" + workdir, "-I" + reftree.REF,
           os.path.join(workdir, "harness.c"), os.path.join(ROOT, "host", "pcamv_bitstream.c"),
           os.path.join(ROOT, "oracle", "_ref", "libx264_wide.a"), "-lm", "-lpthread", "-o", exe])
    subprocess.check_call(cmd)
    return exe


def mutate(rng, data):
    b = bytearray(data)
    kind = rng.randrange(6)
    if kind == 0:                                   # a few byte flips anywhere
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    elif kind == 1:                                 # truncation
        b = b[:rng.randrange(1, len(b))]
    elif kind == 2:                                 # random bytes in the parameter sets / first slice header
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(min(64, len(b)))] = rng.randrange(256)
    elif kind == 3:                                 # a run of zeros or ones
        p, n = rng.randrange(len(b)), rng.randrange(1, 200)
        b[p:p + n] = bytes([rng.choice((0, 255))]) * min(n, len(b) - p)
    elif kind == 4:                                 # splice two places of the stream
        p, q, n = rng.randrange(len(b)), rng.randrange(len(b)), rng.randrange(1, 400)
        b[p:p + n] = b[q:q + n]
    else:                                           # random garbage behind valid headers
        p = rng.randrange(min(200, len(b)), len(b))
        b[p:] = bytes(rng.randrange(256) for _ in range(min(2000, len(b) - p)))
    return bytes(b)


def main():
    import pcamv_loader
    import refrun
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    pcamv = pcamv_loader.load()
    with tempfile.TemporaryDirectory() as wd:
        exe = build(wd)
        seeds = []
        for k, args in enumerate(["--qp 26 --ref 2 --keyint 250 --me hex --subme 4 --emrate 0.2",
                                  "--qp 22 --ref 3 --keyint 4 --me hex --subme 4 --emrate 0.2 --partitions all --no-cabac",
                                  "--qp 36 --ref 1 --keyint 250 --me dia --subme 2 --emrate 0.2 --partitions all"]):
            clip = refrun.synth_clip(pcamv, 176, 144, 5, config=1, stream=40 + k, noise16=16, workdir=wd)
            out, _ = refrun.run_ref(clip, 176, 144, args.split(), binary="x264_dump_conformant", out=os.path.join(wd, "seed%d.264" % k))
            seeds.append(open(out, "rb").read())
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
        clean = rejected = 0
        for i in range(cases):
            data = seeds[i % len(seeds)] if i < len(seeds) else mutate(rng, seeds[rng.randrange(len(seeds))])
            path = os.path.join(wd, "case.264")
            open(path, "wb").write(data)
            for mode in (["--parse-mv", path, "-o", os.path.join(wd, "o.bin")],
                         ["--extract-264", path, "--emrate", "0.2", "--stego", os.path.join(wd, "s.bin"), "-o", os.path.join(wd, "m.bin")]):
                p = subprocess.run([exe, "x"] [:0] + [exe] + mode, capture_output=True, env=env, timeout=120)
                if p.returncode not in (0, 1) or b"Sanitizer" in p.stderr or b"runtime error" in p.stderr:
                    keep = os.path.join(ROOT, "gpurun_out", "fuzz_case_%d.264" % i)
                    os.makedirs(os.path.dirname(keep), exist_ok=True)
                    open(keep, "wb").write(data)
                    print("FAIL case %d (%s): rc %d, kept as %s\n%s" % (i, mode[0], p.returncode, keep, p.stderr[-3000:].decode("latin-1")))
                    return 1
                clean += p.returncode == 0
                rejected += p.returncode == 1
        print("bitstream fuzz: %d inputs x 2 modes under ASan + UBSan: %d parsed, %d refused with a message, 0 sanitizer reports, 0 signals"
              % (cases, clean, rejected))
    return 0


if __name__ == "__main__":
    sys.exit(main())
