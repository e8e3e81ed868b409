"""Comparison of pcamv_analyse_p results with the reference encoder's records (dumps of oracle/_ref/x264_dump).
Shared by tests/test_gpu_frame.py, bench.py's parity gate and __graft_entry__.smoke()."""
import numpy as np


def open_ctx(pcamv, dump, s, device=0, rows_per_cta=1, pass2_elide=0):
    c = dump.cfg
    ctx = pcamv.PcamvContext(s.width, s.lines_y, me_method=c["me_method"], me_range=c["me_range"],
                             subpel_refine=c["subme"], chroma_me=c["chroma_me"], max_refs=max(c["refs"], 1),
                             mv_range=c["mv_range"], b_cabac=c["b_cabac"], b_fast_pskip=c["fast_pskip"],
                             b_dct_decimate=c["dct_decimate"], analyse_inter=c["inter"], device=device, rows_per_cta=rows_per_cta, pass2_elide=pass2_elide)
    t = dump.cost_tables[s.qp]
    q = dump.quant_tables()[s.qp]
    ctx.set_qp_tables(s.qp, t["lambda"], t["cost_mv"], t["cost_ref"], lambda2_chroma=q["lambda2_chroma"],
                      chroma_qp=q["chroma_qp"], quant4_mf=q["quant4_mf"], quant4_bias=q["quant4_bias"],
                      dequant4_mf=q["dequant4_mf"])
    return ctx


def pass1_records(pcamv, embd):
    mbs = embd["mbs"]
    out = np.zeros(len(mbs), dtype=pcamv.host.PASS1_MB_DTYPE)
    for k in ("type", "partition", "used", "sub", "ref", "mv", "mv_stego"):
        out[k] = mbs[k]
    return out


def check_dump(pcamv, dump, units=None, ctx=None, keep_ctx=False):
    """Analyses every recorded P-slice pass on the GPU and compares; returns counters, raises on the first mismatch."""
    units = dump.slice_units() if units is None else units
    n = {"passes": 0, "calls": 0, "mbs": 0, "ih": 0}
    last_embd = {}
    last_mv = {}            # frame -> cache MVs of the last MB of pass 1 (what pass 2 finds in the MV cache)
    for u in units:
        s = u["slice"]
        if not s.with_planes or s.qp not in dump.cost_tables:
            continue
        if ctx is None:
            ctx = open_ctx(pcamv, dump, s)
        x = u["ctx"]
        H, W = s.lines_y, s.width
        ctx.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        for slot, r in enumerate(s.refs):
            ctx.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                        r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
        if s.pass_ == 2:
            e = last_embd[s.frame]
            kw.update(pass1=pass1_records(pcamv, e), filp=e["filp"], stale_mv=last_mv[s.frame])
        mbs, log = ctx.analyse_p(s.pass_, list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"], cost_table=True, **kw)
        n_mb = x["n_mb"]
        # (1) call log
        calls, refine = u["calls"], u["refine"]
        order = np.argsort(calls["mb_xy"], kind="stable")
        calls, refine = calls[order], refine[order]
        counts = np.bincount(calls["mb_xy"], minlength=n_mb)
        n_ih = np.where(mbs["type"] != 6, mbs["n_part"], 0) if (s.pass_ == 1 and u["embd"] is not None) else np.zeros(n_mb, int)
        assert (mbs["n_log"] == counts + n_ih).all(), "frame %d pass %d: log lengths differ at MBs %s" % (
            s.frame, s.pass_, np.nonzero(mbs["n_log"] != counts + n_ih)[0][:8])
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        slot = np.arange(len(calls)) - starts[calls["mb_xy"]]
        got = log[calls["mb_xy"], slot]
        ok = (got["kind"] == refine.astype(np.int8)) & (got["i_pixel"] == calls["i_pixel"]) & \
             (got["mv"] == calls["mv"]).all(axis=1) & (got["cost"] == calls["cost"])
        thr_exit = (~refine) & (calls["has_thresh"] != 0)        # cost_mv is left stale by the multi-ref early-out
        ok &= thr_exit | (got["cost_mv"] == calls["cost_mv"])
        assert ok.all(), "frame %d pass %d: %d of %d search results differ (first at MB %d)" % (
            s.frame, s.pass_, (~ok).sum(), len(ok), calls["mb_xy"][np.argmin(ok)])
        n["calls"] += len(calls)
        # (2) decisions
        m = u["mban"]
        g = mbs[m["mb_xy"]]
        okm = (g["type"] == m["type"]) & ((g["type"] == 6) | (g["partition"] == m["partition"]))
        okm &= (g["mv"] == m["mv"]).all(axis=(1, 2)) & (g["ref"] == m["ref"]).all(axis=1) & (g["pskip_mv"] == m["pskip_mv"]).all(axis=1)
        assert okm.all(), "frame %d pass %d: %d macroblock decisions differ (first MB %d)" % (
            s.frame, s.pass_, (~okm).sum(), m["mb_xy"][np.argmin(okm)])
        n["mbs"] += len(m)
        # (3) cost table
        if s.pass_ == 1 and u["embd"] is not None:
            e = u["embd"]
            last_embd[s.frame] = e
            last_mv[s.frame] = mbs["mv"][n_mb - 1]
            for mb in np.nonzero(mbs["type"] != 6)[0]:
                r = mbs[mb]
                if int(r["type"]) == 5:
                    # P_8x8: info.cache slots of the (sub-)blocks in the reference's order (encoder/analyse.c:3546-3606), from the
                    # reference's own sub-partition types; a block's original vector is the final cache MV of its first 4x4
                    slots = []
                    for i8, kind in enumerate(e["mbs"][mb]["sub"]):
                        slots += [4 * i8 + d for d in {3: [0], 1: [0, 2], 2: [0, 1], 0: [0, 1, 2, 3]}[int(kind)]]
                    assert int(r["n_part"]) == len(slots), "frame %d MB %d: %d MV-carrying blocks, reference has %d" % (s.frame, mb, r["n_part"], len(slots))
                    origs = [r["mv"][sl] for sl in slots]
                else:
                    slots = [{16: 0, 14: 8 * k, 15: 4 * k}[int(r["partition"])] for k in range(r["n_part"])]
                    origs = [r["part"][k]["mv"] for k in range(r["n_part"])]
                for k, (sl, orig) in enumerate(zip(slots, origs)):
                    le = log[mb, counts[mb] + k]
                    want = e["mbs"][mb]["mv_stego"][sl] - orig
                    assert le["kind"] == 2 and (le["mv"] == want).all() and le["cost"] == e["mbs"][mb]["inter_stego_cost"][sl], \
                        "frame %d MB %d part %d: cost table differs: got d%s cost %d, reference d%s cost %d" % (
                            s.frame, mb, k, le["mv"], le["cost"], want, e["mbs"][mb]["inter_stego_cost"][sl])
                    n["ih"] += 1
        n["passes"] += 1
    if ctx is not None:
        n["launches"] = ctx.launch_count()
        if not keep_ctx:
            ctx.close()
    return n
