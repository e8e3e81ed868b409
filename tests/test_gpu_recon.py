"""Reference frames built on the device (csrc/pcamv_recon.cuh / pcamv_recon.cu): reconstruction of a
P frame from the decisions its final analysis pass left in HBM, the in-loop deblocking filter, borders and half-pel planes —
(1) through the C ABI against the planes the reference encoder itself held for the NEXT frame (dumps of the instrumented twin),
(2) through the bound host: with PCAMV_DEVICE_RECON=1 no P frame is uploaded as a reference any more, the bitstream must still
    be the reference's, and in check mode every GPU-built picture is compared with the host's own planes before use."""
import os

import numpy as np
import pytest

import refrun
import test_gpu_host as th

pytestmark = pytest.mark.gpu


def check_dump(pcamv, dump):
    import frame_parity
    units = [u for u in dump.slice_units() if u["slice"].with_planes]
    frames = sorted({u["slice"].frame for u in units})
    n = {"compared": 0, "q1_frames": 0}
    ctx = None
    for fr in frames:
        if fr + 1 not in frames:
            continue
        us = [u for u in units if u["slice"].frame == fr]
        nxt = next(u for u in units if u["slice"].frame == fr + 1)["slice"]
        s, x = us[0]["slice"], us[0]["ctx"]
        if ctx is None:
            ctx = frame_parity.open_ctx(pcamv, dump, s)
        H, W = s.lines_y, s.width
        ctx.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        n_ref = x["n_ref"]
        for slot, r in enumerate(s.refs):
            ctx.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                        r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
        refs, pocs = list(range(n_ref)), x["ref_poc"][:n_ref]
        if len(us) == 2:
            e = us[0]["embd"]
            m1, _ = ctx.analyse_p(1, refs, pocs, x["cur_poc"], cost_table=True, **kw)
            m, _ = ctx.analyse_p(2, refs, pocs, x["cur_poc"], pass1=frame_parity.pass1_records(pcamv, e), filp=e["filp"],
                                 stale_mv=m1["mv"][-1].copy(), **kw)
            final = 2
        else:
            final = s.pass_
            m, _ = ctx.analyse_p(final, refs, pocs, x["cur_poc"], **kw)
        if (m["early_skip"] == 2).any():
            n["q1_frames"] += 1          # needs the host's patches: covered through the bound host below
            continue
        target = n_ref               # a slot that is not one of the frame's references
        ctx.reconstruct_ref(target, x["cur_poc"], final)
        want = nxt.refs[0]
        for k in range(4):
            got = ctx.get_ref_plane(target, k)
            assert np.array_equal(got, want["luma"][k]), "frame %d: luma plane %d of the GPU-built reference differs (%d bytes)" % (
                fr, k, int((got != want["luma"][k]).sum()))
        assert np.array_equal(ctx.get_ref_plane(target, 4), want["u"]) and np.array_equal(ctx.get_ref_plane(target, 5), want["v"])
        n["compared"] += 1
    if ctx is not None:
        ctx.close()
    return n


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_dia2_lownoise"])
def test_reconstructed_reference_matches_golden(pcamv, cuda_lib, name, tmp_path):
    n = check_dump(pcamv, pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path))))
    assert n["compared"] >= 1


LIVE = [
    ("--me hex --subme 5 --ref 1 --emrate 0.2", 32),
    ("--me umh --subme 5 --ref 3 --emrate 0.2", 32),
    ("--me hex --subme 5 --ref 1 --partitions p8x8,p4x4 --emrate 0.2", 32),       # sub-8x8 vectors: per-4x4 motion compensation, inner-edge bS
    ("--me hex --subme 4 --ref 2 --no-dct-decimate --qp 34 --emrate 0.2", 16),
    ("--me hex --subme 5 --ref 1", 32),                                           # no embedding: the single pass is the final one
    ("--me hex --subme 5 --ref 1 --deblock 2:-1 --emrate 0.2", 32),
]


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,noise", LIVE)
def test_reconstructed_reference_matches_live(pcamv, cuda_lib, args, noise, tmp_path):
    clip = refrun.synth_clip(pcamv, 352, 288, 5, config=1, stream=8, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--qp 26 --keyint 250 " + args).split(), dump=dump, frames="1:4")
    d = pcamv.dumpfmt.Dump(dump)
    if "--deblock" in args:
        pytest.skip("the dump does not record the deblocking offsets (covered through the bound host)")
    n = check_dump(pcamv, d)
    assert n["compared"] + n["q1_frames"] >= 2


HOST_CASES = [c for c in th.CASES if c[0] in ("cif_hex5", "cif_umh5_ref3", "cif_dia2_lownoise", "cif_hex4_idr", "cif_noembed", "cif_p4x4_umh_ref3",
                                              "cif_qp48_skips", "tiny_48x32", "cif_nocabac", "cif_nodecimate_qp36", "odd_size_umh", "1080p_umh5")]
HOST_CASES += [
    ("cif_deblock_offsets", 352, 288, 8, 1, 32, "x264_wide", "--qp 30 --ref 2 --keyint 250 --me hex --subme 5 --deblock 2:-1 --emrate 0.2"),
    ("cif_no_deblock", 352, 288, 6, 1, 32, "x264_wide", "--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --nf --emrate 0.2"),
    ("cif_qp12_filter_off", 352, 288, 5, 1, 32, "x264_wide", "--qp 12 --ref 1 --keyint 250 --me hex --subme 4 --emrate 0.2"),
]


@pytest.mark.parametrize("case", HOST_CASES, ids=[c[0] for c in HOST_CASES])
def test_host_with_device_built_references(pcamv, cuda_lib, case, tmp_path):
    """No P frame is uploaded as a reference: the GPU rebuilds each from its final pass (with the host's patches for the
    macroblocks of quirk q1), check mode compares every such picture with the host's planes before it is used, and the bitstream
    is still the reference's."""
    ref_out, out, stats = th.encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env={"PCAMV_DEVICE_RECON": "1", "PCAMV_CHECK_RECON": "1"})
    assert th.md5(out) == th.md5(ref_out)
    assert stats["recon_frames"] >= 2 and stats["recon_mismatch"] == 0, stats
    if case[0] == "cif_qp48_skips":
        assert stats["recon_patched_mbs"] > 0            # the q1 path was really exercised
    # and without the safety net: whatever the GPU built is what the next frames were searched in
    # — including the half-pel planes the host's own motion compensation reads: they come back from the GPU, x264_frame_filter is
    # skipped for these frames
    ref_out, out, stats = th.encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env={"PCAMV_DEVICE_RECON": "1"})
    assert th.md5(out) == th.md5(ref_out)
    assert stats["hpel_frames"] == stats["recon_frames"] >= 2, stats
    ref_out, out, stats = th.encode_pair(pcamv, *case, workdir=str(tmp_path), extra_env={"PCAMV_DEVICE_RECON": "1", "PCAMV_HOST_HPEL": "1"})
    assert th.md5(out) == th.md5(ref_out) and stats["hpel_frames"] == 0
