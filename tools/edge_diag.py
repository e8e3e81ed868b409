#!/usr/bin/env python3
"""Which layer differs at a tiny frame size?  Reference dump of a 48x32 clip vs (1) GPU planes, (2) frame seam, (3) GPU trellis."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pcamv_loader, refrun, frame_parity, test_gpu_stc
pcamv = pcamv_loader.load()
w, h = int(sys.argv[1]), int(sys.argv[2])
ARGS = sys.argv[3] if len(sys.argv) > 3 else "--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"
SYNTH = sys.argv[4].split() if len(sys.argv) > 4 else ["6", "1", "0", "32"]
wd = tempfile.mkdtemp()
clip = os.path.join(wd, "e.yuv")
subprocess.check_call([os.path.join(ROOT, "build", "pcamv_synth"), str(w), str(h), SYNTH[0], SYNTH[1], SYNTH[2], clip, SYNTH[3]])
dumpf = os.path.join(wd, "d.bin")
refrun.run_ref(clip, w, h, ARGS.split(), dump=dumpf)
dump = pcamv.dumpfmt.Dump(dumpf)
for s in dump.slices():
    if not s.with_planes: continue
    ctx = frame_parity.open_ctx(pcamv, dump, s)
    for slot, r in enumerate(s.refs):
        H, W = s.lines_y, s.width
        ctx.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
        for k in range(4):
            got = ctx.get_ref_plane(slot, k)
            if not np.array_equal(got, r["luma"][k]):
                bad = np.argwhere(got != r["luma"][k])
                print("frame", s.frame, "pass", s.pass_, "slot", slot, "luma plane", k, "differs at", len(bad), "bytes, first (row,col)", bad[0], "stride", got.shape)
        for k, nm in ((4, "u"), (5, "v")):
            if not np.array_equal(ctx.get_ref_plane(slot, k), r[nm]): print("chroma plane", nm, "differs")
    ctx.close()
try:
    print("frame seam:", frame_parity.check_dump(pcamv, dump))
except AssertionError as e:
    print("frame seam MISMATCH:", str(e)[:300])
try:
    print("stc:", test_gpu_stc.check_embeds(pcamv, dump.embeds()))
except AssertionError as e:
    print("stc MISMATCH:", str(e)[:300])
for e in dump.embeds(): print("embed frame", e["frame"], "n", e["length"], "an", e["an"])
