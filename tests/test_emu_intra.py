"""csrc/pcamv_intra.cuh — intra mode analysis of a macroblock (first pieces; not on the product path) — against the
reference's own analysis: every call of x264_mb_analyse_intra in I and P slices ('INTR' records of oracle/_ref/x264_dump_rd) must
give the same cost for every 16x16 luma mode and the same best mode, the same chroma cost and mode, and the same 4x4 result - the
summed cost or "gave up", and the mode of every block the analysis got to (which includes the intra encode between blocks)."""
import os
import subprocess

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    pytest.param("--qp 26 --ref 1 --me hex --subme 5", 32, id="qp26"),
    pytest.param("--qp 14 --ref 1 --me hex --subme 5 --keyint 2 --min-keyint 2", 48, id="qp14-every-other-frame-intra"),
    pytest.param("--qp 40 --ref 2 --me dia --subme 4", 4, id="qp40-flat"),
    pytest.param("--qp 26 --ref 1 --me hex --subme 6 --no-cabac", 32, id="rd-thresholds"),
    pytest.param("--qp 30 --ref 1 --me hex --subme 5 --no-chroma-me", 16, id="chroma-analysed-later"),
]


@pytest.fixture(scope="module")
def checker(pcamv):
    return pcamv.build.build_tool("emu_intra_check", os.path.join(ROOT, "tests", "emu", "emu_intra_check.cpp"))


@pytest.mark.skipif(not refrun.have_ref("x264_dump_rd"), reason="oracle/_ref/x264_dump_rd not built")
@pytest.mark.parametrize("args,noise", CASES)
def test_intra_mode_analysis_equals_reference(pcamv, checker, args, noise, tmp_path):
    clip = refrun.synth_clip(pcamv, 352, 288, 4, config=1, stream=8, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--keyint 250 --emrate 0.2 " + args).split(), binary="x264_dump_rd", dump=dump, planes=False, calls=False)
    p = subprocess.run([checker, dump], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr[-1500:]
    n = {k: int(v) for k, v in (kv.split("=") for kv in p.stdout.split())}
    assert n["bad16"] == 0 and n["bad_chroma"] == 0 and n["bad4x4"] == 0
    assert n["luma16x16"] > 1500 and n["in_i_slices"] >= 396 and n["at_picture_border"] > 100 and n["luma4x4_completed"] > 300
