"""Search logic of the device code, checked on the CPU with a lane team of one (-DPCAMV_EMU) against
calls recorded from the reference encoder: committed golden fixtures always, and live runs of
oracle/_ref/x264_dump when that binary is present."""
import os
import subprocess

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(pcamv):
    return pcamv.build.build_tool("emu_search_check", os.path.join(ROOT, "tests", "emu", "emu_search_check.cpp"))


def run_checker(checker, dump):
    p = subprocess.run([checker, dump], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    calls = int(p.stdout.split("calls=")[1].split()[0])
    assert "mismatches=0" in p.stdout
    return calls


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2", "qcif_esa5", "qcif_tesa5", "qcif_dia2_lownoise"])
def test_golden_calls(checker, name, tmp_path):
    assert run_checker(checker, refrun.golden_dump_path(name, str(tmp_path))) > 1000


LIVE = [
    ("--me hex --subme 5 --ref 1", "1:3"),
    ("--me umh --subme 7 --ref 1", "1:3"),
    ("--me umh --subme 5 --ref 3", "3:5"),
    ("--me dia --subme 1 --ref 1", "1:3"),
    ("--me hex --subme 3 --ref 2 --partitions all --mixed-refs", "2:4"),
    ("--me esa --merange 24 --subme 4 --ref 1", "1:2"),
    ("--me tesa --merange 16 --subme 5 --ref 1 --partitions all", "1:2"),      # ADS on the 8x8 and the 4x4 integral planes
]


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref/x264_dump not built")
@pytest.mark.parametrize("args,frames", LIVE)
def test_live_reference_calls(pcamv, checker, args, frames, tmp_path):
    clip = refrun.synth_clip(pcamv, 352, 288, 5, config=1, stream=3, workdir=str(tmp_path))
    dump = str(tmp_path / "d.bin")
    refrun.run_ref(clip, 352, 288, ("--qp 26 --keyint 250 --emrate 0.2 " + args).split(), dump=dump, frames=frames)
    assert run_checker(checker, dump) > 5000
