"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports every
symbol include/pcamv.h declares.  No compute call is made here (there is no GPU in this container)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcamv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcamv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cuda_lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(cuda_lib, s), "libpcamv_cuda.so does not export %s" % s


def test_host_mirror_lists_every_export(pcamv):
    assert sorted(pcamv.host.EXPORTS) == declared_symbols()


def test_abi_version_and_record_layout(pcamv, cuda_lib):
    assert cuda_lib.pcamv_abi_version() == 6
    # sizes of the POD records as laid out by the C compiler (see include/pcamv.h)
    assert pcamv.host.ME_CALL_DTYPE.itemsize == 132
    assert pcamv.host.ME_RESULT_DTYPE.itemsize == 16


def test_open_fails_loudly_without_gpu(pcamv, cuda_lib):
    """No CPU fallback: without a device pcamv_open returns -1 and says why."""
    import torch
    if torch.cuda.is_available():
        return
    cfg = pcamv.host.Cfg()
    cfg.abi_version = cuda_lib.pcamv_abi_version()
    cfg.width, cfg.height, cfg.max_refs = 176, 144, 1
    h = ctypes.c_void_p()
    assert cuda_lib.pcamv_open(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert b"no CUDA device" in cuda_lib.pcamv_last_error(None) or b"cuda" in cuda_lib.pcamv_last_error(None).lower()


def test_open_rejects_bad_arguments(pcamv, cuda_lib):
    cfg = pcamv.host.Cfg()
    cfg.abi_version = 999
    h = ctypes.c_void_p()
    assert cuda_lib.pcamv_open(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert b"ABI" in cuda_lib.pcamv_last_error(None)
    cfg.abi_version = cuda_lib.pcamv_abi_version()
    cfg.width, cfg.height, cfg.max_refs = 100, 100, 1       # not multiples of 16
    assert cuda_lib.pcamv_open(ctypes.byref(h), ctypes.byref(cfg)) == -1


def test_host_encoder_carries_no_cpu_search():
    """No CPU fallback, provable with nm: the GPU host links the reference's C encoder WITHOUT the reference's own motion
    search (x264_me_search_ref / refine_subpel / x264_me_refine_qpel bodies, x264_ih_get_mv_cost) — the only
    x264_me_search_ref in the binary is the replay stub of host/pcamv_x264_glue.c, which pops GPU results."""
    import subprocess
    exe = os.path.join(ROOT, "host", "_build", "x264_pcamv")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("host/_build/x264_pcamv is not built")
    syms = subprocess.run(["nm", exe], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in syms.splitlines() if l.strip()]
    for bad in ("x264_me_search_ref_real", "x264_me_refine_qpel_real", "x264_ih_get_mv_cost_real", "refine_subpel"):
        assert bad not in names, "%s is linked into the GPU host" % bad
    assert "x264_me_search_ref" in names and "pcamv_hook_slice_begin" in names
    # and it is bound to the CUDA library, not to a CPU copy of it
    dyn = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    assert "pcamv_analyse_p" in dyn and "pcamv_open" in dyn
