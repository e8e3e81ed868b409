"""Packed-pixel primitives of the device code (pcamv_prims.cuh / pcamv_me.cuh under -DPCAMV_EMU) against the plain-C leaf
oracle on random, near-equal and saturating inputs: the packed-pair 4x4 Hadamard (two 16-bit halves per word must not
overflow at 0-vs-255 blocks), packed SAD, the rounding average and the two-pixels-per-multiply chroma interpolation."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_primitives_equal_leaf_oracle(tmp_path):
    obj = str(tmp_path / "leaf_oracle.o")
    exe = str(tmp_path / "emu_prims_check")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "leaf_oracle.c")])
    subprocess.check_call(["g++", "-O2", "-w", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_prims_check.cpp"), obj])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "mismatches=0" in p.stdout, p.stdout + p.stderr
    assert int(p.stdout.split("checks=")[1].split()[0]) > 2000000
