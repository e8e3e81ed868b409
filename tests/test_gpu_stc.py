"""Embed stage on the GPU: pcamv_stc_embed against the reference's own stc_embed.

The reference encoder's dumps (EMBD records of oracle/_ref/x264_dump: golden fixtures and live runs) hold, per P frame, the
cover bits, the float embedding costs rho_final, the message and the stego vector its stc_embed produced
(embed.h:309-548, called at encoder/encoder.c:1843).  From the same cover / rho / message the GPU trellis must return the
same stego vector bit for bit; its syndrome (x264_pcamv --extract) must be the message."""
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
    """getMatrix(width, height) of the reference (embed.h:276-306), through the host binary; built-in tables only
    (widths 2..20): beyond that the reference draws from an LCG whose state depends on the encoder's history."""
    assert 2 <= width <= 20
    if (width, height) not in _cols:
        out = subprocess.run([HOST, "--stc-columns", str(width), str(height)], capture_output=True, check=True).stdout
        _cols[(width, height)] = np.array([int(x) for x in out.split()], dtype=np.uint32)
    return _cols[(width, height)]


def check_embeds(pcamv, embeds):
    ctx = pcamv.PcamvContext(176, 144)
    n_ok = bits = 0
    for e in embeds:
        n, an = e["length"], e["an"]
        if an < 1 or an > n:
            continue
        shorter, longer = n // an, -(-n // an)
        if not (2 <= shorter and longer <= 20):
            continue
        stego = ctx.stc_embed(e["cover"], e["message"][:an], e["rho"], columns(shorter), columns(longer))
        if stego is None:
            # "not in the range of the syndrome matrix": the reference's stc_embed then returns 0 and leaves the zeroed stego
            # buffer alone (encoder/encoder.c:1826,1843) — happens for messages shorter than the matrix height
            assert not e["stego"].any(), "frame %d: GPU trellis reports the message as not embeddable, the reference embedded it" % e["frame"]
            n_ok += 1
            continue
        assert np.array_equal(stego, e["stego"]), "frame %d: stego vector differs from the reference's (%d of %d bits)" % (
            e["frame"], int((stego != e["stego"]).sum()), n)
        n_ok += 1; bits += an
    ctx.close()
    return n_ok, bits


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_tesa5", "qcif_dia2_lownoise"])
def test_stc_embed_matches_reference_golden(pcamv, cuda_lib, name, tmp_path):
    n_ok, bits = check_embeds(pcamv, pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path))).embeds())
    assert n_ok >= 1 and bits > 10


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("size,emrate,extra", [((48, 32), "0.3", "--ref 2"),        # messages shorter than the matrix height: the reference's
                                               ((64, 48), "0.5", ""),               # forward and backward column masks differ (embed.h:415,519-524)
                                               ((352, 288), "0.2", ""), ((352, 288), "0.1", "--partitions all"), ((1280, 720), "0.3", ""),
                                               ((1920, 1080), "0.2", "")])
def test_stc_embed_matches_reference_live(pcamv, cuda_lib, size, emrate, extra, tmp_path):
    w, h = size
    frames = 6 if w < 1000 else 3
    tiny = w < 100
    clip = refrun.synth_clip(pcamv, w, h, frames, config=2, stream=3, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, w, h, ("--qp 26 --ref 1 --keyint 250 --me hex --subme 5 --emrate %s %s" % (emrate, extra)).split(), dump=dump,
                   planes=False, calls=False)
    n_ok, bits = check_embeds(pcamv, pcamv.dumpfmt.Dump(dump).embeds())
    assert n_ok == frames - 1 and bits > (5 if tiny else 100)


def test_stc_embed_unembeddable_and_bad_arguments(pcamv, cuda_lib):
    ctx = pcamv.PcamvContext(176, 144)
    c5 = columns(5)
    # infinite costs everywhere but a message that needs flips: not in the range of the (wet) matrix -> None, like stc_embed's 0
    cover = np.zeros(50, np.uint8); msg = np.ones(10, np.uint8); rho = np.full(50, np.inf, np.float32)
    assert ctx.stc_embed(cover, msg, rho, c5, c5) is None
    with pytest.raises(pcamv.PcamvError, match="widths"):
        pcamv.PcamvContext(176, 144).stc_embed(cover, msg, np.ones(50, np.float32), columns(4), c5)
    # an even column would make the per-state survivor rule differ from the reference's pairwise update once masked to zero
    even = np.array(c5, dtype=np.uint32).copy(); even[2] &= ~np.uint32(1)
    with pytest.raises(pcamv.PcamvError, match="odd"):
        pcamv.PcamvContext(176, 144).stc_embed(cover, msg, np.ones(50, np.float32), even, c5)
    ctx.close()
