#!/usr/bin/env python3
"""Per-macroblock timeline of one wavefront launch at 1080p: busy time per MB, waiting, critical path."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import frame_parity  # noqa: E402
import pcamv_loader  # noqa: E402


def main():
    pcamv = pcamv_loader.load()
    workdir = tempfile.mkdtemp(prefix="pcamv_trace_")
    clip, dumpf = bench.prepare_inputs(pcamv, 0, workdir)
    dump = pcamv.dumpfmt.Dump(dumpf)
    units = [u for u in dump.slice_units() if u["slice"].frame == bench.BATCH_FRAME and u["slice"].with_planes]
    s, x = units[0]["slice"], units[0]["ctx"]
    H, W = s.lines_y, s.width
    r = s.refs[0]
    col = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
    refs, pocs, cur_poc = list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"]
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1          # contexts running concurrently (context 0 is traced)
    rpc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    others = []
    for i in range(S - 1):
        o = frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rpc)
        o.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        o.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
        o.frame_upload(1, refs, pocs, cur_poc, cost_table=False, **col)
        others.append(o)
    import threading
    stop = []
    def spin(o):
        while not stop:
            o.frame_run(1, 1)
    th = [threading.Thread(target=spin, args=(o,)) for o in others]
    for x_ in th: x_.start()
    c = frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rpc)
    c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
    c.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
    c.frame_upload(1, refs, pocs, cur_poc, cost_table=False, **col)
    c.frame_run(1, 2)
    c.frame_trace(True)
    ms = c.frame_run(1, 1)
    tr = c.frame_trace(False, fetch=True).astype(np.int64)
    stop.append(1)
    for x_ in th: x_.join()
    mbs, log = c.frame_download()
    mb_w, mb_h = W // 16, H // 16
    t0 = tr[:, 0].min()
    st = (tr[:, 0] - t0).reshape(mb_h, mb_w) / 1e3
    en = (tr[:, 1] - t0).reshape(mb_h, mb_w) / 1e3
    busy = en - st
    gap = np.zeros_like(busy)
    gap[:, 1:] = st[:, 1:] - en[:, :-1]
    n_log = mbs["n_log"].reshape(mb_h, mb_w)
    typ = mbs["type"].reshape(mb_h, mb_w)
    out = {"contexts": S, "rows_per_cta": rpc, "kernel_ms": ms, "span_us": float(en.max()), "busy_us_mean": float(busy.mean()), "busy_us_median": float(np.median(busy)),
           "busy_us_p90": float(np.percentile(busy, 90)), "busy_us_max": float(busy.max()),
           "gap_us_mean": float(gap.mean()), "gap_us_median": float(np.median(gap)),
           "row0_total_us": float(en[0, -1]), "row0_busy_us": float(busy[0].sum()),
           "sum_busy_ms": float(busy.sum() / 1e3), "searches_per_mb": float(n_log.mean()),
           "busy_us_skip": float(busy[typ == 6].mean()) if (typ == 6).any() else None,
           "busy_us_nonskip": float(busy[typ != 6].mean()), "skip_frac": float((typ == 6).mean())}
    print(json.dumps(out))
    np.save(os.path.join(ROOT, "gpurun_out", "trace_busy.npy"), busy.astype(np.float32))


if __name__ == "__main__":
    main()
