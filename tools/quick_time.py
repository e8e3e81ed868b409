#!/usr/bin/env python3
"""A/B timing of the frame-seam kernels for an experimental library build.

    PCAMV_LIB=build/variants/lib_x.so python tools/quick_time.py [S] [rows_per_cta] [reps]

Runs the parity gate of bench.py (one 1080p P frame, both passes, bit-exact vs the reference dump) on the library
named by PCAMV_LIB, then times the multi-context launches (pass-1 wavefront, cost table, pass-2 wavefront) over S
contexts and prints one JSON line.  Not a bench: no L2 handling, no clocks sampling — only for comparing builds."""
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
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rpc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    pcamv = pcamv_loader.load()
    workdir = os.environ.get("PCAMV_QT_DIR") or tempfile.mkdtemp(prefix="pcamv_qt_")
    os.makedirs(workdir, exist_ok=True)
    clip, dumpf = bench.prepare_inputs(pcamv, 0, workdir)
    dump = pcamv.dumpfmt.Dump(dumpf)
    units = [u for u in dump.slice_units() if u["slice"].frame == bench.BATCH_FRAME and u["slice"].with_planes]
    s, x, e = units[0]["slice"], units[0]["ctx"], units[0]["embd"]
    ctxs = [frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rpc) for _ in range(S)]
    par = frame_parity.check_dump(pcamv, dump, units=units, ctx=ctxs[0], keep_ctx=True)
    H, W = s.lines_y, s.width
    r = s.refs[0]
    col = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
    pass1 = frame_parity.pass1_records(pcamv, e)
    refs, pocs, cur_poc = list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"]
    m_ref = None
    for c in ctxs:
        c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        c.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
        c.frame_upload(1, refs, pocs, cur_poc, cost_table=True, **col)
        m1, lg1 = c.analyse_p(1, refs, pocs, cur_poc, cost_table=True, **col)
        if m_ref is None:
            l1_ref = lg1.copy()
        c.frame_upload(2, refs, pocs, cur_poc, pass1=pass1, filp=e["filp"], stale_mv=m1["mv"][-1], **col)
        if m_ref is None:
            m_ref = m1.copy()
        assert (m1["mv"] == m_ref["mv"]).all()
    # PCAMV_QT_SWEEP="ctrl_sms:rows,ctrl_sms:rows,...": split wavefront settings (read by the library at every launch)
    sweep = [tuple(x.split(":")) for x in os.environ.get("PCAMV_QT_SWEEP", "").split(",") if x] or [None]
    for setting in sweep:
        if setting:
            os.environ["PCAMV_SPLIT_CTRL_SMS"], os.environ["PCAMV_SPLIT_ROWS"] = setting[:2]
            if len(setting) > 2:        # ctrl_sms:rows:phase_ns[:patience_ns]
                os.environ["PCAMV_SPLIT_PHASE_NS"] = setting[2]
                os.environ["PCAMV_SPLIT_PATIENCE_NS"] = setting[3] if len(setting) > 3 else "2000"
        time_one(pcamv, ctxs, S, rpc, reps, m_ref, l1_ref, par, setting)


def time_one(pcamv, ctxs, S, rpc, reps, m_ref, l1_ref, par, setting):
    acc = np.zeros(3)
    for k in range(reps + 1):
        _, w1, ct = pcamv.host.frame_run_batch(ctxs, 1)
        if k == 0:
            # the multi-context launch must reproduce the single launch, log included, in every context looked at
            l_ref = None
            for c in (ctxs[S // 2], ctxs[-1]):        # (same history: stale log entries beyond n_log are equal too)
                mb, lg = c.frame_download()
                if l_ref is None:
                    l_ref, b_ref = lg, mb
                assert (mb["mv"] == m_ref["mv"]).all() and (mb["type"] == m_ref["type"]).all() and (mb["partition"] == m_ref["partition"]).all(), "batch launch: decisions differ from the single launch"
                assert mb.tobytes() == b_ref.tobytes(), "batch launch: contexts differ"
                # every log entry the macroblocks own equals the single launch's (search results, refinements, cost table)
                valid = np.arange(lg.shape[1])[None, :] < mb["n_log"][:, None]
                a = lg.view(np.uint8).reshape(lg.shape[0], lg.shape[1], -1)[valid]
                b = l1_ref.view(np.uint8).reshape(lg.shape[0], lg.shape[1], -1)[valid]
                assert (mb["n_log"] == m_ref["n_log"]).all() and (a == b).all(), "batch launch: logs differ from the single launch"
        _, w2, _ = pcamv.host.frame_run_batch(ctxs, 2)
        if k:
            acc += (w1, ct, w2)
    acc /= reps
    print(json.dumps({"lib": os.path.basename(os.environ.get("PCAMV_LIB", "default")), "S": S, "rows_per_cta": rpc, "split": setting,
                      "pass1_ms": round(float(acc[0]), 2), "cost_table_ms": round(float(acc[1]), 2), "pass2_ms": round(float(acc[2]), 2),
                      "total_ms": round(float(acc.sum()), 2), "parity": "ok %d searches" % par["calls"]}), flush=True)


if __name__ == "__main__":
    main()
