"""GPU parity tests proper: every call goes through the C-ABI of libpcamv_cuda.so (ctypes)."""
import os

import numpy as np
import pytest

import refrun

pytestmark = pytest.mark.gpu

GOLDEN = ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_tesa5", "qcif_dia2_lownoise"]


def open_ctx(pcamv, dump, slice_):
    c = dump.cfg
    ctx = pcamv.PcamvContext(slice_.width, slice_.lines_y, me_method=c["me_method"], me_range=c["me_range"],
                             subpel_refine=c["subme"], chroma_me=c["chroma_me"], max_refs=max(c["refs"], 1),
                             mv_range=c["mv_range"], b_cabac=c["b_cabac"], b_fast_pskip=c["fast_pskip"],
                             b_dct_decimate=c["dct_decimate"], analyse_inter=c["inter"])
    assert ctx.plane_stride(0) == slice_.stride_y and ctx.plane_stride(4) == slice_.stride_c
    return ctx


@pytest.mark.parametrize("name", GOLDEN)
def test_frame_filter_matches_reference_planes(pcamv, cuda_lib, name, tmp_path):
    """A6/A7: borders + H/V/HV planes built on the GPU from the integer reconstruction are byte-identical,
    padding included, to the planes the reference encoder held (x264_frame_filter + expand_border*)."""
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path)))
    checked = 0
    for s in dump.slices():
        if not s.with_planes:
            continue
        ctx = open_ctx(pcamv, dump, s)
        for slot, r in enumerate(s.refs):
            H, W = s.lines_y, s.width
            y = r["luma"][0][32:32 + H, 32:32 + W]
            u = r["u"][16:16 + H // 2, 16:16 + W // 2]
            v = r["v"][16:16 + H // 2, 16:16 + W // 2]
            ctx.put_ref(slot, r["poc"], y, u, v)
            for k in range(4):
                got = ctx.get_ref_plane(slot, k)
                assert np.array_equal(got, r["luma"][k]), "luma plane %d of slot %d differs" % (k, slot)
            assert np.array_equal(ctx.get_ref_plane(slot, 4), r["u"])
            assert np.array_equal(ctx.get_ref_plane(slot, 5), r["v"])
            checked += 1
        ctx.close()
    assert checked >= 1


def run_calls(pcamv, dump, use_gpu_filter):
    calls, refine = dump.calls()
    tables = dump.cost_tables
    total = bad = 0
    for s in dump.slices():
        if not s.with_planes or s.pass_ == 2:
            continue
        ctx = open_ctx(pcamv, dump, s)
        t = tables[s.qp]
        ctx.set_qp_tables(s.qp, t["lambda"], t["cost_mv"], t["cost_ref"])
        ctx.put_fenc(s.fenc[0][:, :s.width], s.fenc[1][:, :s.width // 2], s.fenc[2][:, :s.width // 2])
        for slot, r in enumerate(s.refs):
            if use_gpu_filter:
                H, W = s.lines_y, s.width
                ctx.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                            r["v"][16:16 + H // 2, 16:16 + W // 2])
            else:
                ctx.put_ref_planes(slot, r["poc"], r["luma"], r["u"], r["v"])
        sel = calls["frame"] == s.frame          # both passes read the same planes
        rc, rf = calls[sel], refine[sel]
        res = ctx.me_search_batch(pcamv.dumpfmt.calls_to_abi(rc, rf))
        ok = (res["mv"] == rc["mv"]).all(axis=1) & (res["cost"] == rc["cost"])
        thr = (rc["has_thresh"] != 0) & ~rf
        ok &= np.where(thr, res["thresh_out"] == rc["thresh_out"], res["cost_mv"] == rc["cost_mv"])
        total += len(rc)
        bad += int((~ok).sum())
        ctx.close()
    return total, bad


@pytest.mark.parametrize("name", GOLDEN)
def test_search_calls_match_reference(pcamv, cuda_lib, name, tmp_path):
    """A1-A5, A8: every recorded x264_me_search_ref / x264_me_refine_qpel call reproduces mv, cost, cost_mv and
    the multi-ref half-pel threshold bit-exactly, reading planes the GPU filtered itself."""
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path)))
    total, bad = run_calls(pcamv, dump, use_gpu_filter=True)
    assert total > 1000 and bad == 0, (total, bad)


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,frames,size", [
    ("--me umh --subme 5 --ref 1", "1:3", (352, 288)),
    ("--me hex --subme 5 --ref 3 --partitions all --mixed-refs", "3:5", (352, 288)),
    ("--me umh --subme 5 --ref 1", "1:2", (1280, 720)),
    ("--me tesa --merange 16 --subme 5 --ref 1 --partitions all", "1:2", (352, 288)),     # tesa through the search seam, both integral planes
    ("--me esa --merange 24 --subme 4 --ref 2 --partitions all", "2:3", (352, 288)),
])
def test_search_calls_live_reference(pcamv, cuda_lib, args, frames, size, tmp_path):
    w, h = size
    clip = refrun.synth_clip(pcamv, w, h, 5 if w < 1000 else 2, config=2, stream=1, workdir=str(tmp_path))
    dumpf = str(tmp_path / "d.bin")
    refrun.run_ref(clip, w, h, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), dump=dumpf, frames=frames)
    dump = pcamv.dumpfmt.Dump(dumpf)
    total, bad = run_calls(pcamv, dump, use_gpu_filter=True)
    assert total > 5000 and bad == 0, (total, bad)


@pytest.mark.parametrize("w,h,mode", [(176, 144, "random"), (352, 288, "saturating"), (1280, 720, "random"), (1920, 1088, "edges")])
def test_put_ref_planes_equal_plain_c_oracle(pcamv, cuda_lib, w, h, mode):
    """A6/A7 on inputs no reference dump covers (random, saturating 0/255, hard edges): border expansion + the 6-tap
    half-pel planes built on the GPU equal the plain-C restatement (oracle/leaf_oracle.c, itself pinned against the
    reference's function tables and planes by tests/test_oracle_leaf.py)."""
    import ctypes as C
    from test_oracle_leaf import build_oracle_lib
    lib = build_oracle_lib()
    rng = np.random.default_rng(w * 7 + h)
    if mode == "random":
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    elif mode == "saturating":
        y = (rng.integers(0, 2, (h, w), dtype=np.uint8) * 255).astype(np.uint8)
    else:
        y = np.zeros((h, w), dtype=np.uint8); y[:, ::7] = 255; y[::5, :] = 255; y[h // 2:, w // 2:] = 128
    u = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    ctx = pcamv.PcamvContext(w, h)
    ctx.put_ref(0, 0, y, u, v)
    S, Sc = ctx.plane_stride(0), ctx.plane_stride(4)
    bufs = [np.zeros((h + 64, S), dtype=np.uint8) for _ in range(4)]
    bufs[0][32:32 + h, 32:32 + w] = y
    ptrs = (C.c_void_p * 4)(*[b.ctypes.data + 32 * S + 32 for b in bufs])
    lib.pcamv_oracle_frame_planes(ptrs, S, w, h)
    for k in range(4):
        got = ctx.get_ref_plane(0, k)
        assert np.array_equal(got[:, :w + 64], bufs[k][:, :w + 64]), "luma plane %d" % k
    for k, src in ((4, u), (5, v)):
        c = np.zeros((h // 2 + 32, Sc), dtype=np.uint8)
        c[16:16 + h // 2, 16:16 + w // 2] = src
        lib.pcamv_oracle_chroma_border(C.c_void_p(c.ctypes.data + 16 * Sc + 16), Sc, w // 2, h // 2)
        assert np.array_equal(ctx.get_ref_plane(0, k)[:, :w // 2 + 32], c[:, :w // 2 + 32]), "chroma plane %d" % k
    ctx.close()


def test_integral_plane_matches_oracle(pcamv, cuda_lib, tmp_path):
    """A11: the integral plane k_box_sum8 builds for the exhaustive searches equals the restatement of the reference's
    integral_init8h / 8v chain (oracle/leaf_oracle.c, pinned against the reference's own functions in test_oracle_leaf.py)
    wherever the reference defines it: every 8x8 box inside the padded plane."""
    import ctypes as C
    import test_oracle_leaf
    lib = test_oracle_leaf.build_oracle_lib()
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path("qcif_esa5", str(tmp_path)))
    s = next(x for x in dump.slices() if x.with_planes)
    ctx = open_ctx(pcamv, dump, s)
    r = s.refs[0]
    H, W = s.lines_y, s.width
    ctx.put_ref(0, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2], r["v"][16:16 + H // 2, 16:16 + W // 2])
    got = ctx.get_integral(0)
    plane = np.ascontiguousarray(r["luma"][0])
    rows, stride = plane.shape
    assert got.shape == (rows, stride)
    want = np.zeros((rows + 1, stride), dtype=np.uint16)
    lib.pcamv_oracle_integral8(C.c_void_p(plane.ctypes.data), C.c_void_p(want.ctypes.data), stride, rows)
    assert np.array_equal(got[:rows - 8, :stride - 8], want[:rows - 8, :stride - 8])
    assert got[:rows - 8, :stride - 8].max() > 1000
    ctx.close()
