"""Embed stage on the device (csrc/pcamv_embed.cu): pcamv_embed_prepare / pcamv_embed_stc against the
reference's own embed stage (encoder/encoder.c:1561-1855) as recorded by the instrumented twin ('EMBD' records: info.cache[]
of every macroblock, cover, rho_final, message, stego, filp), then pass 2 from the device-resident forced decisions against
pass 2 fed with the reference's records."""
import os
import subprocess

import numpy as np
import pytest

import refrun

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "_build", "x264_pcamv")
_cols = {}


def columns(width, height=10):
    if (width, height) not in _cols:
        out = subprocess.run([HOST, "--stc-columns", str(width), str(height)], capture_output=True, check=True).stdout
        _cols[(width, height)] = np.array([int(x) for x in out.split()], dtype=np.uint32)
    return _cols[(width, height)]


def check_dump(pcamv, dump, max_frames=3):
    import frame_parity
    units = [u for u in dump.slice_units() if u["slice"].with_planes]
    frames = sorted({u["slice"].frame for u in units})[:max_frames]
    ctx = None
    n = {"frames": 0, "carriers": 0, "bits": 0, "penalised": 0}
    for fr in frames:
        u1 = next(u for u in units if u["slice"].frame == fr and u["slice"].pass_ == 1)
        u2 = next((u for u in units if u["slice"].frame == fr and u["slice"].pass_ == 2), None)
        s, x, e = u1["slice"], u1["ctx"], u1["embd"]
        if e is None or u2 is None:
            continue
        if ctx is None:
            ctx = frame_parity.open_ctx(pcamv, dump, s)
        H, W = s.lines_y, s.width
        ctx.put_fenc(s.fenc[0][:, :W], s.fenc[1][:, :W // 2], s.fenc[2][:, :W // 2])
        for slot, r in enumerate(s.refs):
            ctx.put_ref(slot, r["poc"], r["luma"][0][32:32 + H, 32:32 + W], r["u"][16:16 + H // 2, 16:16 + W // 2],
                        r["v"][16:16 + H // 2, 16:16 + W // 2])
        kw = dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"])
        refs, pocs = list(range(x["n_ref"])), x["ref_poc"][:x["n_ref"]]
        m1, _ = ctx.analyse_p(1, refs, pocs, x["cur_poc"], cost_table=True, **kw)
        # (1) cover / rho / info.cache records
        length = ctx.embed_prepare()
        assert length == e["length"], "frame %d: %d carriers, the reference has %d" % (fr, length, e["length"])
        cover, rho, _, _, p1 = ctx.embed_download(want_stego=False)
        assert np.array_equal(cover, e["cover"]), "frame %d: cover bits differ" % fr
        assert np.array_equal(rho.view(np.uint32), e["rho"].view(np.uint32)), "frame %d: rho_final differs (%d of %d)" % (
            fr, int((rho != e["rho"]).sum()), length)
        em = e["mbs"]
        used = em["used"] != 0
        assert np.array_equal(p1["type"], em["type"]) and np.array_equal(p1["used"] != 0, used)
        pl0 = em["type"] == 4
        assert np.array_equal(p1["partition"][pl0], em["partition"][pl0])
        p8 = em["type"] == 5
        assert np.array_equal(p1["sub"][p8], em["sub"][p8])
        assert np.array_equal(p1["ref"][used], em["ref"][used]) and np.array_equal(p1["mv"][used], em["mv"][used])
        # mv_stego: only the slots the reference writes this frame are comparable (the others keep older frames' values there)
        for mb in np.nonzero(used)[0]:
            if em["type"][mb] == 4:
                slots = {16: [0], 15: [0, 4], 14: [0, 8]}[int(em["partition"][mb])]
            else:
                slots = []
                for i8, kind in enumerate(em["sub"][mb]):
                    slots += [4 * i8 + d for d in {3: [0], 1: [0, 2], 2: [0, 1], 0: [0, 1, 2, 3]}[int(kind)]]
            assert np.array_equal(p1["mv_stego"][mb][slots], em["mv_stego"][mb][slots]), "frame %d MB %d: mv_stego differs" % (fr, mb)
        n["penalised"] += int((e["rho"] != np.round(e["rho"])).sum())
        # (2) trellis on the device-resident vectors, flips, forced decisions
        an = e["an"]
        if 1 <= an <= length and 2 <= length // an and -(-length // an) <= 20:
            rc, stego = ctx.embed_stc(e["message"][:an], columns(length // an), columns(-(-length // an)))
            n["bits"] += an if rc == 0 else 0
        else:
            rc, stego = ctx.embed_stc(None, None, None)
        assert np.array_equal(stego, e["stego"]), "frame %d: stego vector differs" % fr
        _, _, _, filp, _ = ctx.embed_download()
        assert np.array_equal(filp, e["filp"]), "frame %d: flips differ" % fr
        # (3) pass 2 from the forced decisions in HBM == pass 2 from the reference's records
        kw2 = dict(pass1=frame_parity.pass1_records(pcamv, e), filp=e["filp"], stale_mv=m1["mv"][-1].copy(), **kw)
        m2a, l2a = ctx.analyse_p(2, refs, pocs, x["cur_poc"], **kw2)
        m2a, l2a = m2a.copy(), l2a.copy()
        m2b, l2b = ctx.analyse_p(2, refs, pocs, x["cur_poc"], device_forced=True, stale_mv=m1["mv"][-1].copy(), **kw)
        assert m2a.tobytes() == m2b.tobytes(), "frame %d: pass 2 differs with the device-resident forced decisions" % fr
        valid = np.arange(l2a.shape[1])[None, :] < m2a["n_log"][:, None]
        assert (l2a.view(np.uint8).reshape(l2a.shape[0], l2a.shape[1], -1)[valid] == l2b.view(np.uint8).reshape(l2a.shape[0], l2a.shape[1], -1)[valid]).all()
        n["frames"] += 1; n["carriers"] += length
    if ctx is not None:
        ctx.close()
    return n


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_dia2_lownoise"])
def test_embed_stage_matches_reference_golden(pcamv, cuda_lib, name, tmp_path):
    n = check_dump(pcamv, pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path))))
    assert n["frames"] >= 1 and n["carriers"] > 50


LIVE = [
    ("--me hex --subme 5 --ref 1", (352, 288), 32),
    ("--me umh --subme 5 --ref 2 --partitions all", (352, 288), 24),          # P_8x8 with 4x4 / 8x4 / 4x8 splits: every MVC penalty class
    ("--me hex --subme 5 --ref 1 --partitions p8x8,p4x4 --emrate 0.4", (352, 288), 32),
    ("--me dia --subme 4 --ref 1 --qp 34", (352, 288), 2),                     # P_SKIP-heavy: few carriers per frame
    ("--me hex --subme 5 --ref 1", (48, 32), 32),                              # a handful of carriers: an < matrix height, nothing embeddable
]


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,size,noise", LIVE)
def test_embed_stage_matches_reference_live(pcamv, cuda_lib, args, size, noise, tmp_path):
    w, h = size
    clip = refrun.synth_clip(pcamv, w, h, 5, config=1, stream=6, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    extra = [] if "--emrate" in args else ["--emrate", "0.2"]
    refrun.run_ref(clip, w, h, ("--qp 26 --keyint 250 " + args).split() + extra, dump=dump, frames="1:4")
    n = check_dump(pcamv, pcamv.dumpfmt.Dump(dump))
    assert n["frames"] >= 2
    if "partitions all" in args:
        assert n["penalised"] > 0          # the x (0.7 n + 1) class really occurred
