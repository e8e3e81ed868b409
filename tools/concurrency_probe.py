#!/usr/bin/env python3
"""How does whole-GPU throughput of the frame seam scale with the number of independent encoder contexts
(GOP shards / streams) analysing concurrently, each on its own CUDA stream?  Prints one JSON line per S."""
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import frame_parity  # noqa: E402
import pcamv_loader  # noqa: E402


def main():
    pcamv = pcamv_loader.load()
    counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32]
    workdir = tempfile.mkdtemp(prefix="pcamv_conc_")
    clip, dumpf = bench.prepare_inputs(pcamv, 0, workdir)
    dump = pcamv.dumpfmt.Dump(dumpf)
    units = [u for u in dump.slice_units() if u["slice"].frame == bench.BATCH_FRAME and u["slice"].with_planes]
    s, x, e = units[0]["slice"], units[0]["ctx"], units[0]["embd"]
    H, W = s.lines_y, s.width
    r = s.refs[0]
    col = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
    pass1 = frame_parity.pass1_records(pcamv, e)
    refs, pocs, cur_poc = list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"]
    ctxs = []
    for i in range(max(counts)):
        c = frame_parity.open_ctx(pcamv, dump, s)
        c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        c.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
        c.frame_upload(1, refs, pocs, cur_poc, cost_table=True, **col)
        m1, _ = c.analyse_p(1, refs, pocs, cur_poc, cost_table=True, **col)
        c.frame_upload(2, refs, pocs, cur_poc, pass1=pass1, filp=e["filp"], stale_mv=m1["mv"][-1], **col)
        ctxs.append(c)
    iters = 3
    for S in counts:
        def work(c):
            c.frame_run(1, iters)
            c.frame_run(2, iters)
        for rep in range(2):
            th = [threading.Thread(target=work, args=(c,)) for c in ctxs[:S]]
            t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            dt = time.perf_counter() - t0
        print(json.dumps({"contexts": S, "frames": S * iters, "wall_s": dt, "frames_per_s": S * iters / dt,
                          "ms_per_frame_per_ctx": dt / iters * 1e3}), flush=True)


if __name__ == "__main__":
    main()
