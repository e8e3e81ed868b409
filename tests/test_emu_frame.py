"""Frame-level device code (neighbour cache, MV prediction, P_SKIP probe, partition decision, pass-2 forcing,
PCAMV cost table) checked on the CPU with a lane team of one (-DPCAMV_EMU) against the reference encoder's
records: committed golden fixtures always, live runs of oracle/_ref/x264_dump when that binary is present."""
import os
import subprocess

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(pcamv):
    return pcamv.build.build_tool("emu_frame_check", os.path.join(ROOT, "tests", "emu", "emu_frame_check.cpp"))


def run_checker(checker, dump, async_mode=False, conformant=False, expect_bad=False):
    env = dict(os.environ, PCAMV_EMU_ASYNC="1" if async_mode else "0", PCAMV_EMU_CONFORMANT="1" if conformant else "0")
    p = subprocess.run([checker, dump], capture_output=True, text=True, env=env)
    if expect_bad:
        out = dict(kv.split("=") for kv in p.stdout.split())
        return {k: int(v) for k, v in out.items()}
    if async_mode:
        assert "searches handed out" in p.stderr and int(p.stderr.split("async:")[1].split()[0]) > 1000, p.stderr
    assert p.returncode == 0, p.stdout + p.stderr
    out = dict(kv.split("=") for kv in p.stdout.split())
    assert out["bad_mb_logs"] == "0" and out["bad_decisions"] == "0" and out["bad_ih"] == "0", p.stdout
    # reconstruction + deblocking of every frame whose successor was dumped too == the reference's own reference plane
    assert out["bad_recon"] == "0", p.stdout + p.stderr[-1500:]
    return {k: int(v) for k, v in out.items()}


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_tesa5", "qcif_dia2_lownoise"])
def test_golden_frames(checker, name, tmp_path):
    n = run_checker(checker, refrun.golden_dump_path(name, str(tmp_path)))
    assert n["passes"] >= 2 and n["calls"] > 1000 and n["ih"] > 100
    assert n["recon"] >= 1 or name in ("qcif_esa5", "qcif_tesa5")          # (those two fixtures hold a single frame)


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_dia2_lownoise"])
def test_golden_frames_resumable_analysis(checker, name, tmp_path):
    """The split wavefront's form of the analysis: every search leaves the macroblock as a request, the macroblock's
    persistent bytes are parked and everything else is overwritten, a separate team serves the request, the analysis
    resumes — and the records still equal the reference's."""
    n = run_checker(checker, refrun.golden_dump_path(name, str(tmp_path)), async_mode=True)
    assert n["passes"] >= 2 and n["calls"] > 1000


LIVE = [
    ("--me hex --subme 5 --ref 1", "1:4", 32),
    ("--me umh --subme 5 --ref 3", "3:5", 32),
    ("--me hex --subme 4 --ref 2 --no-fast-pskip", "2:4", 32),
    ("--me dia --subme 2 --ref 1 --qp 32", "1:4", 4),          # subme < 3: skips only through the neighbour-gated probe
    ("--me dia --subme 4 --ref 2 --qp 34", "1:4", 2),          # low noise: a quarter to a third P_SKIP, exercises the pass-2 quirks
    ("--me hex --subme 5 --ref 1 --no-cabac", "1:3", 16),
    ("--me esa --merange 16 --subme 5 --ref 2", "1:3", 32),            # successive elimination on the integral plane
    ("--me tesa --merange 24 --subme 3 --ref 2", "1:3", 32),
    ("--me hex --subme 5 --ref 1 --partitions p8x8,p4x4", "1:3", 32),   # sub-8x8 partitions: P_8x8 with 4x4 / 8x4 / 4x8 splits
    ("--me umh --subme 5 --ref 2 --partitions all", "2:4", 24),
    ("--me dia --subme 1 --ref 1 --partitions all --qp 30", "1:3", 24),           # SATD as fpelcmp, ADS/SAD thresholds, list pruning
]


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,frames,noise", LIVE)
def test_live_reference_frames(pcamv, checker, args, frames, noise, tmp_path):
    clip = refrun.synth_clip(pcamv, 352, 288, 5, config=1, stream=5, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), dump=dump, frames=frames)
    n = run_checker(checker, dump)
    assert n["passes"] >= 2 and n["calls"] > 5000
    if "--partitions" not in args or "p4x4" not in args and "all" not in args:
        run_checker(checker, dump, async_mode=True)


CONFORMANT = [
    ("--me dia --subme 4 --ref 2 --qp 34", "1:4", 2),                      # skip-heavy: macroblocks forced to P_SKIP (quirk q2)
    ("--me hex --subme 5 --ref 1 --partitions p8x8,p4x4 --qp 22", "1:3", 32),   # forced P_8x8 whose own pass-2 analysis chose 16x8 / 8x16
    ("--me umh --subme 5 --ref 2 --partitions all", "2:4", 24),
]


@pytest.mark.skipif(not refrun.have_ref("x264_dump_conformant"), reason="oracle/_ref/x264_dump_conformant not built")
@pytest.mark.parametrize("args,frames,noise", CONFORMANT)
def test_live_conformant_frames(pcamv, checker, args, frames, noise, tmp_path):
    """pcamv_set_conformant (include/pcamv.h): the device logic with the switch on against the reference with the same three
    statements corrected (oracle/_ref/x264_dump_conformant, tools/reftree.py::conformance_switch) - searches, decisions of both
    passes incl. the vectors of forced skips and the partition of forced P_8x8, cost table, reconstructed reference frame.
    Negative controls: the switch off against that dump, and the switch on against the unmodified reference, must both differ."""
    clip = refrun.synth_clip(pcamv, 352, 288, 5, config=1, stream=5, noise16=noise, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), binary="x264_dump_conformant", dump=dump, frames=frames)
    n = run_checker(checker, dump, conformant=True)
    assert n["passes"] >= 2 and n["calls"] > 5000
    if "--partitions" not in args:
        run_checker(checker, dump, async_mode=True, conformant=True)
    off = run_checker(checker, dump, conformant=False, expect_bad=True)
    assert off["bad_decisions"] > 0
    refrun.run_ref(clip, 352, 288, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), binary="x264_dump", dump=dump, frames=frames)
    on = run_checker(checker, dump, conformant=True, expect_bad=True)
    assert on["bad_decisions"] > 0
    run_checker(checker, dump)          # and the default against the unmodified reference stays exact
