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

GOLDEN = ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_dia2_lownoise"]


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
])
def test_frame_analysis_live_reference(pcamv, cuda_lib, args, frames, size, tmp_path):
    w, h = size
    clip = refrun.synth_clip(pcamv, w, h, 5 if w < 1000 else 2, config=2, stream=2, workdir=str(tmp_path))
    dumpf = str(tmp_path / "d.bin")
    refrun.run_ref(clip, w, h, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), dump=dumpf, frames=frames)
    n = check_dump(pcamv, pcamv.dumpfmt.Dump(dumpf))
    assert n["passes"] >= 2 and n["calls"] > 5000, n


def test_frame_seam_rejects_unsupported(pcamv, cuda_lib):
    """subme >= 6 (RD) and sub-8x8 partitions are refused loudly, never silently approximated."""
    ctx = pcamv.PcamvContext(176, 144, subpel_refine=7)
    cm = np.zeros(32769, dtype=np.int16)
    z16 = np.zeros(16, dtype=np.uint16); z96 = np.zeros(96, dtype=np.int32)
    ctx.set_qp_tables(26, 4, cm, np.zeros(99, np.uint16), quant4_mf=(z16, z16), quant4_bias=(z16, z16), dequant4_mf=(z96, z96))
    y = np.zeros((144, 176), np.uint8); c = np.zeros((72, 88), np.uint8)
    ctx.put_fenc(y, c, c); ctx.put_ref(0, 0, y, c, c)
    with pytest.raises(pcamv.PcamvError, match="subpel_refine"):
        ctx.analyse_p(0, [0], [0], 2)
    ctx.close()
