#!/usr/bin/env python3
"""CPU-only sweep: device logic under emulation (tests/emu/emu_frame_check) against live reference dumps over a grid of
options, sizes and seeds.  Prints one line per case; exit code 1 if any case differs.
    python tools/emu_sweep.py [seed] [cases] [conformant]
`conformant`: the device logic with pcamv_set_conformant's switch on against oracle/_ref/x264_dump_conformant."""
import itertools, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
chk = pcamv.build.build_tool("emu_frame_check", os.path.join(ROOT, "tests", "emu", "emu_frame_check.cpp"))
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
conformant = len(sys.argv) > 3 and sys.argv[3] == "conformant"
wd = tempfile.mkdtemp(prefix="pcamv_sweep_")
bad = 0
for case in range(n):
    w, h = rnd.choice([(176, 144), (352, 288), (64, 48), (96, 64), (48, 32), (320, 240), (16, 16), (128, 16), (16, 96)])
    me = rnd.choice(["dia", "hex", "umh", "esa", "tesa"])
    subme = rnd.choice([1, 2, 3, 4, 5])
    ref = rnd.choice([1, 1, 2, 3, 4])
    parts = rnd.choice(["", "--partitions all", "--partitions p8x8,p4x4", "--partitions none"])
    extra = rnd.choice(["", "", "--no-cabac", "--no-fast-pskip", "--no-dct-decimate", "--no-chroma-me", "--mvrange 24"])
    qp = rnd.choice([18, 26, 32, 38, 44, 50])
    merange = rnd.choice([4, 8, 12, 16]) if me in ("esa", "tesa") else rnd.choice([8, 16, 24])
    em = rnd.choice(["0.1", "0.3", "0.7", "20"])
    noise = rnd.choice([0, 2, 8, 32])
    frames = 5
    args = "--qp %d --ref %d --keyint 250 --me %s --merange %d --subme %d --emrate %s %s %s" % (qp, ref, me, merange, subme, em, parts, extra)
    clip = os.path.join(wd, "c%d.yuv" % case)
    subprocess.check_call([pcamv.build.build_synth(), str(w), str(h), str(frames), "1", str(100 + case), clip, str(noise)])
    dump = os.path.join(wd, "d%d.bin" % case)
    try:
        refrun.run_ref(clip, w, h, args.split(), dump=dump, frames="1:4", binary="x264_dump_conformant" if conformant else "x264_dump")
    except subprocess.CalledProcessError:
        print("REF-FAIL | %dx%d | %s" % (w, h, args)); continue
    p = subprocess.run([chk, dump], capture_output=True, text=True, env=dict(os.environ, PCAMV_EMU_CONFORMANT="1" if conformant else "0"))
    ok = p.returncode == 0
    bad += not ok
    print("%s | %dx%d noise %d | %s | %s" % ("OK  " if ok else "DIFF", w, h, noise, args, p.stdout.strip()[-90:] if ok else (p.stdout + p.stderr)[-400:]), flush=True)
    os.remove(clip); os.remove(dump)
sys.exit(1 if bad else 0)
