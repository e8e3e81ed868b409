"""csrc/pcamv_cavlc.cuh — the CAVLC size of an inter macroblock as a host/device function — against the reference's RD mode
decision: every macroblock x264_rd_cost_mb (encoder/rdo.c:139-172) sized with x264_macroblock_size_cavlc at --subme 6
--no-cabac ('RDMB' records of oracle/_ref/x264_dump_rd) must get the same bit count.  The first parity-tested piece of
--subme 6 / 7; not on the product path, which still refuses RD mode decision."""
import os
import subprocess

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    pytest.param("--qp 26 --ref 1 --me hex", 32, id="qp26"),
    pytest.param("--qp 26 --ref 3 --me umh --partitions all", 24, id="ref3-sub8x8"),          # ref_idx as te(v), sub_mb_types as ue(v)
    pytest.param("--qp 26 --ref 2 --me hex", 32, id="ref2-one-bit-ref-idx"),
    pytest.param("--qp 4 --ref 1 --me hex --partitions p8x8,p4x4", 64, id="qp4-long-level-codes"),
    pytest.param("--qp 14 --ref 1 --me hex --no-dct-decimate", 48, id="qp14-no-decimate"),
    pytest.param("--qp 40 --ref 2 --me dia", 4, id="qp40-sparse"),
]


@pytest.fixture(scope="module")
def checker(pcamv):
    return pcamv.build.build_tool("emu_cavlc_check", os.path.join(ROOT, "tests", "emu", "emu_cavlc_check.cpp"))


@pytest.mark.skipif(not refrun.have_ref("x264_dump_rd"), reason="oracle/_ref/x264_dump_rd not built")
@pytest.mark.parametrize("args,noise", CASES)
def test_cavlc_macroblock_size_equals_reference_rd(pcamv, checker, args, noise, tmp_path):
    clip = refrun.synth_clip(pcamv, 352, 288, 4, config=1, stream=6, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--keyint 250 --emrate 0.2 --subme 6 --no-cabac " + args).split(), binary="x264_dump_rd", dump=dump,
                   planes=False, calls=False)
    p = subprocess.run([checker, dump], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr[-1500:]
    n = {k: int(v) for k, v in (kv.split("=") for kv in p.stdout.split())}
    assert n["bad"] == 0 and n["inter"] > 2000 and n["coded"] > 100
    if "--partitions all" in args:
        assert n["sub8x8"] > 0 and n["multi_ref"] > 0


RD_CASES = [
    pytest.param("--qp 26 --ref 1 --me hex", 32, id="qp26"),
    pytest.param("--qp 26 --ref 3 --me umh --partitions all", 24, id="ref3-sub8x8"),
    pytest.param("--qp 14 --ref 1 --me hex --no-dct-decimate", 48, id="qp14-no-decimate"),
    pytest.param("--qp 38 --ref 2 --me dia", 4, id="qp38-sparse"),
    pytest.param("--qp 26 --ref 1 --me hex --psy-rd 0:0", 32, id="no-psy"),
]


@pytest.fixture(scope="module")
def rd_checker(pcamv):
    return pcamv.build.build_tool("emu_rd_check", os.path.join(ROOT, "tests", "emu", "emu_rd_check.cpp"))


@pytest.mark.skipif(not refrun.have_ref("x264_dump_rd"), reason="oracle/_ref/x264_dump_rd not built")
@pytest.mark.parametrize("args,noise", RD_CASES)
def test_rd_cost_of_inter_candidates_equals_reference(pcamv, rd_checker, args, noise, tmp_path):
    """Both halves of x264_rd_cost_mb (encoder/rdo.c:139-172) for every inter candidate RD mode decision costed: the distortion
    (SSD + psy term, csrc/pcamv_rd.cuh) of the candidate as the PRODUCT's device code reconstructs it (recon_mb) and its CAVLC
    size (csrc/pcamv_cavlc.cuh); also the set of luma blocks that keep coefficients."""
    clip = refrun.synth_clip(pcamv, 352, 288, 4, config=1, stream=6, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--keyint 250 --emrate 0.2 --subme 6 --no-cabac " + args).split(), binary="x264_dump_rd", dump=dump, frames="1:3")
    p = subprocess.run([rd_checker, dump], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr[-1500:]
    n = {k: int(v) for k, v in (kv.split("=") for kv in p.stdout.split())}
    assert n["candidates"] > 1500 and n["bad_distortion"] == 0 and n["bad_bits"] == 0 and n["bad_kept_mask"] == 0
    assert (n["with_psy"] == 0) == ("--psy-rd 0:0" in args)


@pytest.mark.skipif(not os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")), reason="nvcc not present")
def test_rd_pieces_compile_for_sm_100a(tmp_path):
    """The RD pieces are host/device code: a kernel that runs a candidate through motion compensation, kept levels, the product's
    residual path, distortion and CAVLC size compiles for sm_100a (tools/probes/rd_compile_probe.cu; compile only, nothing launches
    it - the pieces are not on the product path)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    p = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v", "-c",
                        os.path.join(ROOT, "tools", "probes", "rd_compile_probe.cu"), "-o", str(tmp_path / "rd_probe.o")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "k_rd_cost_probe" in p.stderr and "k_intra_probe" in p.stderr and "sm_100a" in p.stderr
