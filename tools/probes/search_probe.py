#!/usr/bin/env python3
"""Times k_search_batch (stateless x264_me_search_ref / refine_qpel evaluations, no wavefront, no per-macroblock control
code) on every search the reference made for the bench frame, replicated REP times: what the search code alone costs when
its ~45 KB are the only hot code on the SM.  Compare with the wavefront kernel's time per search (tools/quick_time.py)."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import frame_parity  # noqa: E402
import pcamv_loader  # noqa: E402


def main():
    rep = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    pcamv = pcamv_loader.load()
    workdir = os.environ.get("PCAMV_QT_DIR") or tempfile.mkdtemp(prefix="pcamv_qt_")
    os.makedirs(workdir, exist_ok=True)
    clip, dumpf = bench.prepare_inputs(pcamv, 0, workdir)
    dump = pcamv.dumpfmt.Dump(dumpf)
    units = [u for u in dump.slice_units() if u["slice"].frame == bench.BATCH_FRAME and u["slice"].with_planes]
    s = units[0]["slice"]
    ctx = frame_parity.open_ctx(pcamv, dump, s)
    H, W = s.lines_y, s.width
    r = s.refs[0]
    ctx.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
    ctx.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
    for u in units:
        rc, rf = u["calls"], u["refine"]
        abi = pcamv.dumpfmt.calls_to_abi(rc, rf)
        res = ctx.me_search_batch(abi)
        ok = (res["mv"] == rc["mv"]).all(axis=1) & (res["cost"] == rc["cost"])
        assert ok.all(), "search seam parity"
        big = np.tile(abi, rep)
        ctx.me_batch_upload(big)
        ctx.me_batch_run(1)
        ms = ctx.me_batch_run(3)
        print(json.dumps({"pass": int(u["slice"].pass_), "searches": int(len(abi)), "refines": int(rf.sum()), "rep": rep, "ms_per_launch": round(ms, 3),
                          "ns_per_search": round(ms * 1e6 / len(big), 2)}), flush=True)


if __name__ == "__main__":
    main()
