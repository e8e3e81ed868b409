"""Pins the plain-C restatement oracle/leaf_oracle.c: against the reference's own leaf function tables (needs
oracle/_ref/libx264_wide.a and the reference headers, i.e. this container) and against the reference encoder's planes
in the committed golden dumps (runs anywhere)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PCAMV_REFERENCE", "/root/reference")


def build_oracle_lib():
    out = os.path.join(ROOT, "oracle", "_ref", "libpcamv_oracle.so")
    src = os.path.join(ROOT, "oracle", "leaf_oracle.c")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


@pytest.mark.skipif(not (os.path.isdir(REF) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libx264_wide.a"))),
                    reason="needs the reference headers and oracle/_ref/libx264_wide.a")
def test_restatement_equals_reference_function_tables(tmp_path):
    exe = str(tmp_path / "leaf_check")
    subprocess.check_call(["gcc", "-O2", "-w", "-I" + REF, "-DHAVE_MALLOC_H", "-DARCH_X86_64", "-DSYS_LINUX", "-DHAVE_PTHREAD",
                           "-include", os.path.join(ROOT, "oracle", "leaf_config.h"),
                           "-o", exe, os.path.join(ROOT, "oracle", "leaf_check.c"), os.path.join(ROOT, "oracle", "leaf_oracle.c"),
                           os.path.join(ROOT, "oracle", "_ref", "libx264_wide.a"), "-lm", "-lpthread"])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "mismatches=0" in p.stdout, p.stdout + p.stderr
    assert int(p.stdout.split("checks=")[1].split()[0]) > 20000


@pytest.mark.parametrize("name", ["qcif_hex5", "qcif_umh5_ref2"])
def test_restatement_equals_reference_planes(pcamv, name, tmp_path):
    """Borders + 6-tap half-pel planes of every dumped reference frame, rebuilt from the integer interior alone."""
    lib = build_oracle_lib()
    dump = pcamv.dumpfmt.Dump(refrun.golden_dump_path(name, str(tmp_path)))
    n = 0
    for s in dump.slices():
        if not s.with_planes:
            continue
        H, W, S = s.lines_y, s.width, s.stride_y
        for r in s.refs:
            bufs = [np.zeros((H + 64, S), dtype=np.uint8) for _ in range(4)]
            bufs[0][32:32 + H, 32:32 + W] = r["luma"][0][32:32 + H, 32:32 + W]
            ptrs = (C.c_void_p * 4)(*[b.ctypes.data + 32 * S + 32 for b in bufs])
            lib.pcamv_oracle_frame_planes(ptrs, S, W, H)
            for k in range(4):
                assert np.array_equal(bufs[k][:, :W + 64], r["luma"][k][:, :W + 64]), "plane %d of frame %d" % (k, s.frame)
            for pl in ("u", "v"):
                Sc = s.stride_c
                c = np.zeros((H // 2 + 32, Sc), dtype=np.uint8)
                c[16:16 + H // 2, 16:16 + W // 2] = r[pl][16:16 + H // 2, 16:16 + W // 2]
                lib.pcamv_oracle_chroma_border(C.c_void_p(c.ctypes.data + 16 * Sc + 16), Sc, W // 2, H // 2)
                assert np.array_equal(c[:, :W // 2 + 32], r[pl][:, :W // 2 + 32])
            n += 1
        if n >= 3:
            break
    assert n >= 2
