#!/usr/bin/env python3
"""GPU box: whole-encoder md5 parity (host/_build/x264_pcamv vs oracle/_ref/x264_wide) over a random grid of options, sizes
and seeds.  One line per case; exit code 1 if any case differs.
    python tools/host_sweep.py [seed] [cases] [conformant]
`conformant`: the same sweep for PCAMV_CONFORMANT=1 against oracle/_ref/x264_dump_conformant - embedding always on."""
import hashlib, json, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
conformant = len(sys.argv) > 3 and sys.argv[3] == "conformant"
wd = tempfile.mkdtemp(prefix="pcamv_hsweep_")
synth = os.path.join(ROOT, "build", "pcamv_synth")
bad = 0
for case in range(n):
    w, h = rnd.choice([(176, 144), (352, 288), (64, 48), (96, 64), (48, 32), (320, 240), (16, 16), (128, 16), (16, 96), (640, 368)])
    me = rnd.choice(["dia", "hex", "umh", "esa", "tesa"])
    subme = rnd.choice([1, 2, 3, 4, 5])
    ref = rnd.choice([1, 1, 2, 3, 4])
    parts = rnd.choice(["", "--partitions all", "--partitions p8x8,p4x4", "--partitions none"])
    extra = rnd.choice(["", "", "--no-cabac", "--no-fast-pskip", "--no-dct-decimate", "--no-chroma-me", "--mvrange 24", "--keyint 3 --min-keyint 3"])
    qp = rnd.choice([18, 26, 32, 38, 44, 50])
    merange = rnd.choice([4, 8, 12, 16]) if me in ("esa", "tesa") else rnd.choice([8, 16, 24])
    em = rnd.choice(["0.1", "0.3", "0.7", "20", "0"])
    if conformant and em == "0":
        em = "0.2"
    noise = rnd.choice([0, 2, 8, 32])
    args = ("--qp %d --ref %d --keyint 250 --me %s --merange %d --subme %d %s %s %s" % (
        qp, ref, me, merange, subme, ("--emrate " + em) if em != "0" else "", parts, extra)).split()
    clip = os.path.join(wd, "c.yuv")
    subprocess.check_call([synth, str(w), str(h), "6", "1", str(200 + case), clip, str(noise)])
    outs = []
    # odd cases run in check mode: every reference picture the GPU builds is compared with the host's own planes before use
    stats_file = os.path.join(wd, "stats.json")
    env = dict(os.environ, PCAMV_STATS=stats_file, **({"PCAMV_CHECK_RECON": "1"} if case & 1 else {}), **({"PCAMV_CONFORMANT": "1"} if conformant else {}))
    if os.path.exists(stats_file):
        os.remove(stats_file)
    for exe in (os.path.join(ROOT, "oracle", "_ref", "x264_dump_conformant" if conformant else "x264_wide"), os.path.join(ROOT, "host", "_build", "x264_pcamv")):
        o = os.path.join(wd, "o_%d.264" % len(outs))
        p = subprocess.run([exe] + args + ["-o", o, clip, "%dx%d" % (w, h)], capture_output=True, env=env)
        outs.append((p.returncode, hashlib.md5(open(o, "rb").read()).hexdigest() if os.path.exists(o) else None, p.stderr[-200:]))
    ok = outs[0][:2] == outs[1][:2]
    st = json.load(open(stats_file)) if os.path.exists(stats_file) else {}
    ok = ok and st.get("recon_mismatch", 0) == 0 and st.get("stale_mismatch", 0) == 0
    bad += not ok
    print("%s | %dx%d noise %d | recon %s%s patched %s direct %s | %s%s" % ("OK  " if ok else "DIFF", w, h, noise, st.get("recon_frames"), " (checked)" if case & 1 else "",
                                                                           st.get("recon_patched_mbs"), st.get("direct_pass1"), " ".join(args), "" if ok else " | rc %s/%s %s" % (outs[0][0], outs[1][0], outs[1][2].decode("latin-1")[-160:])), flush=True)
sys.exit(1 if bad else 0)
