"""GPU parity tests of the frame seam (pcamv_analyse_p): the macroblock wavefront and the PCAMV cost table.

For every P-slice pass recorded from the reference encoder (golden fixtures = dumps of oracle/_ref/x264_dump) the
GPU analyses the whole slice from the same pixels and must reproduce, per macroblock and bit-exactly:
  (1) the sequence of x264_me_search_ref / x264_me_refine_qpel results (mv, cost, cost_mv) in call order,
  (2) the decision x264_macroblock_analyse left (type, partition, cache MVs, refs, pskip MV) — both passes,
      pass 2 being driven by the reference's own info.cache[] + filp[] through the C-ABI,
  (3) pass 1: every x264_ih_get_mv_cost result (replacement MV and embedding cost) of info.cache[].
"""
import numpy as np
import pytest

import refrun

pytestmark = pytest.mark.gpu

GOLDEN = ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_tesa5", "qcif_dia2_lownoise"]


from frame_parity import check_dump


@pytest.mark.parametrize("name", GOLDEN)
def test_frame_analysis_matches_reference(pcamv, cuda_lib, name, tmp_path):
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path)))
    n = check_dump(pcamv, dump)
    assert n["passes"] >= 2 and n["calls"] > 1000 and n["mbs"] >= 198 and n["ih"] > 100, n


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,frames,size", [
    ("--me hex --subme 5 --ref 1", "1:4", (352, 288)),
    ("--me umh --subme 5 --ref 3", "3:5", (352, 288)),
    ("--me umh --subme 5 --ref 1", "1:2", (1280, 720)),
    ("--me esa --merange 32 --subme 5 --ref 4", "3:4", (352, 288)),      # BASELINE config 3's search at CIF
    ("--me tesa --merange 24 --subme 4 --ref 2", "2:3", (352, 288)),
    ("--me hex --subme 5 --ref 1 --partitions p8x8,p4x4", "1:3", (352, 288)),      # sub-8x8 partitions
    ("--me umh --subme 5 --ref 2 --partitions all", "2:4", (352, 288)),
    ("--me tesa --merange 16 --subme 5 --ref 1 --partitions p8x8,p4x4", "1:2", (352, 288)),      # 4x4 integral plane
])
def test_frame_analysis_live_reference(pcamv, cuda_lib, args, frames, size, tmp_path):
    w, h = size
    clip = refrun.synth_clip(pcamv, w, h, 5 if w < 1000 else 2, config=2, stream=2, workdir=str(tmp_path))
    dumpf = str(tmp_path / "d.bin")
    refrun.run_ref(clip, w, h, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), dump=dumpf, frames=frames)
    n = check_dump(pcamv, pcamv.dumpfmt.Dump(dumpf))
    assert n["passes"] >= 2 and n["calls"] > 5000, n


def test_frame_seam_rejects_unsupported(pcamv, cuda_lib):
    """subme >= 6 (RD mode decision) is refused loudly, never silently approximated."""
    ctx = pcamv.PcamvContext(176, 144, subpel_refine=7)
    cm = np.zeros(32769, dtype=np.int16)
    z16 = np.zeros(16, dtype=np.uint16); z96 = np.zeros(96, dtype=np.int32)
    ctx.set_qp_tables(26, 4, cm, np.zeros(99, np.uint16), quant4_mf=(z16, z16), quant4_bias=(z16, z16), dequant4_mf=(z96, z96))
    y = np.zeros((144, 176), np.uint8); c = np.zeros((72, 88), np.uint8)
    ctx.put_fenc(y, c, c); ctx.put_ref(0, 0, y, c, c)
    with pytest.raises(pcamv.PcamvError, match="subpel_refine"):
        ctx.analyse_p(0, [0], [0], 2)
    ctx.close()


@pytest.mark.parametrize("rows_per_cta", [4, 2, -1, -2], ids=["row-groups-4", "row-groups-2", "row-pool", "split"])
def test_batch_launch_equals_single(pcamv, cuda_lib, rows_per_cta, tmp_path):
    """pcamv_analyse_p_batch: several encoder contexts (different frames) analysed by ONE wavefront launch give exactly
    the records and logs each context gets on its own — for the row-group kernels, for the row pool (rows as resumable
    tasks claimed by whichever team finds them ready) and for the split wavefront (searches served by search SMs, the
    per-macroblock control code resumed on control SMs, pcamv_split.cu)."""
    import frame_parity
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path("qcif_hex5", str(tmp_path)))
    units = [u for u in dump.slice_units() if u["slice"].with_planes and u["slice"].pass_ == 1][:3]
    assert len(units) == 3
    ctxs, args, single = [], [], []
    for rpc, u in zip((1, 1, 1), units):
        s, x = u["slice"], u["ctx"]
        H, W = s.lines_y, s.width
        c = frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rows_per_cta)
        c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        for slot, r in enumerate(s.refs):
            c.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                      r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"],
                  cost_table=True)
        a = (1, list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"], kw)
        single.append(c.analyse_p(a[0], a[1], a[2], a[3], **kw))
        ctxs.append(c); args.append(a)
    outs = pcamv.host.analyse_p_batch(ctxs, args)
    for (m0, l0), (m1, l1) in zip(single, outs):
        assert m0.tobytes() == m1.tobytes()
        for mb in range(len(m0)):
            n = int(m0["n_log"][mb])
            assert l0[mb, :n].tobytes() == l1[mb, :n].tobytes()
    assert not (single[0][0]["mv"] == single[1][0]["mv"]).all()       # the frames really differ
    ms, w, ct = pcamv.host.frame_run_batch(ctxs, 1, iters=2)
    assert ms > 0 and w > 0 and ct > 0
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("rows_per_cta", [1, 4, -1], ids=["rows-1", "row-groups-4", "row-pool"])
def test_rows_behind_the_wavefront_equal_the_download(pcamv, cuda_lib, rows_per_cta, tmp_path):
    """pcamv_analyse_p_begin / _batch_begin + pcamv_analyse_p_rows: the records and log entries copied out row by row while
    the kernel is still running are the ones the blocking call downloads at its end — one context, and three contexts
    (different frames) in one launch.  Repeated, so that a poll of the row counters that saw the previous launch's final values
    would show as rows reported early with stale contents."""
    import frame_parity
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path("qcif_hex5", str(tmp_path)))
    units = [u for u in dump.slice_units() if u["slice"].with_planes and u["slice"].pass_ == 1][:3]
    ctxs, args, single, outs = [], [], [], []
    for u in units:
        s, x = u["slice"], u["ctx"]
        H, W = s.lines_y, s.width
        c = frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rows_per_cta)
        c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        for slot, r in enumerate(s.refs):
            c.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                      r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
        a = (0, list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]], x["cur_poc"], kw)          # pass 0: the single pass of a frame without embedding
        m, l = c.analyse_p(a[0], a[1], a[2], a[3], **kw)
        single.append((m.copy(), l.copy()))
        ctxs.append(c); args.append(a); outs.append(c.alloc_outputs(pinned=True))
    mb_h = units[0]["slice"].lines_y // 16
    mb_w = units[0]["slice"].width // 16

    def same(got, want, rows):
        m0, l0 = want
        m1, l1 = got
        n_mb = rows * mb_w
        assert m0[:n_mb].tobytes() == m1[:n_mb].tobytes()
        for mb in range(n_mb):
            n = int(m0["n_log"][mb])
            assert l0[mb, :n].tobytes() == l1[mb, :n].tobytes()
    for rep in range(3):
        # one context: ask for every row in turn; whatever is reported complete must already be final
        c, a, o = ctxs[rep % 3], args[rep % 3], outs[rep % 3]
        o[0][:] = 0; o[1][:] = 0
        c.analyse_p_begin(a[0], a[1], a[2], a[3], o, **a[4])
        for r in range(mb_h):
            n = c.analyse_p_rows(r)
            assert r + 1 <= n <= mb_h
            same(o, single[rep % 3], n)
        # three contexts, one launch
        for o in outs:
            o[0][:] = 0; o[1][:] = 0
        pcamv.host.analyse_p_batch_begin(ctxs, args, outs)
        for r in (0, mb_h // 2, mb_h - 1):
            for c, o, w in zip(ctxs, outs, single):
                n = c.analyse_p_rows(r)
                assert n >= r + 1
                same(o, w, n)
    with pytest.raises(pcamv.host.PcamvError):
        ctxs[0].frame_upload(a[0], a[1], a[2], a[3], **a[4]); ctxs[0].analyse_p_rows(0)       # nothing in flight
    for c in ctxs:
        c.close()


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("rows_per_cta", [4, -1, -2], ids=["row-groups-4", "row-pool", "split"])
def test_batch_pass2_equals_single(pcamv, cuda_lib, rows_per_cta, tmp_path):
    """Pass 2 through a multi-context launch on the P_SKIP-heavy clip, where the 'forced skip keeps the previous macroblock's
    MV cache' quirk (analyse.c:2668-2676) makes row starts depend on the whole previous row: the row-group kernels wait for
    it inside the macroblock, the row pool folds it into the readiness test.  Every context must reproduce the single-launch
    pass-2 records (which test_frame_analysis_matches_reference pins against the reference)."""
    import frame_parity
    clip = refrun.synth_clip(pcamv, 352, 288, 5, config=1, stream=5, noise16=0, workdir=str(tmp_path))
    dumpf = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, "--qp 36 --keyint 250 --emrate 0.2 --me hex --subme 3 --ref 1".split(), dump=dumpf, frames="1:4")
    dump = pcamv.dumpfmt.Dump(dumpf)
    units = [u for u in dump.slice_units() if u["slice"].with_planes]
    frames = sorted({u["slice"].frame for u in units})[:3]
    ctxs, args2, single = [], [], []
    for fr in frames:
        u1 = next(u for u in units if u["slice"].frame == fr and u["slice"].pass_ == 1)
        u2 = next(u for u in units if u["slice"].frame == fr and u["slice"].pass_ == 2)
        s, x, e = u1["slice"], u1["ctx"], u1["embd"]
        H, W = s.lines_y, s.width
        c = frame_parity.open_ctx(pcamv, dump, s, rows_per_cta=rows_per_cta)
        c.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        for slot, r in enumerate(s.refs):
            c.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                      r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
        refs, pocs = list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]]
        m1, _ = c.analyse_p(1, refs, pocs, x["cur_poc"], cost_table=True, **kw)
        kw2 = dict(pass1=frame_parity.pass1_records(pcamv, e), filp=e["filp"], stale_mv=m1["mv"][-1].copy(), **kw)
        m2, l2 = c.analyse_p(2, refs, pocs, x["cur_poc"], **kw2)
        single.append((m2.copy(), l2.copy()))
        ctxs.append(c); args2.append((2, refs, pocs, x["cur_poc"], kw2))
    assert all((m["type"] == 6).mean() > 0.2 for m, _ in single)         # really skip-heavy
    outs = pcamv.host.analyse_p_batch(ctxs, args2)
    for k, ((m0, l0), (m1, l1)) in enumerate(zip(single, outs)):
        for field in m0.dtype.names:
            bad = np.nonzero((m0[field] != m1[field]).reshape(len(m0), -1).any(axis=1))[0]
            assert len(bad) == 0, "context %d: field %s differs at %d macroblocks, first %d: single %s batch %s" % (
                k, field, len(bad), bad[0], m0[field][bad[0]], m1[field][bad[0]])
        for mb in range(len(m0)):
            n = int(m0["n_log"][mb])
            assert l0[mb, :n].tobytes() == l1[mb, :n].tobytes()
    for c in ctxs:
        c.close()
