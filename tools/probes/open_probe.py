#!/usr/bin/env python3
"""Where does the start-up of an encoder process go?  Times, in a fresh process: loading libpcamv_cuda.so, the first CUDA call
(context creation), pcamv_open of a 1080p context, its first analysed pass (module load of the kernels it needs), a second
context.  Run once per CUDA_MODULE_LOADING setting."""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    t = [time.perf_counter()]
    rt = C.CDLL("libcudart.so.12")
    t.append(time.perf_counter())
    rt.cudaFree(None)
    t.append(time.perf_counter())
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import numpy as np
    import pcamv_loader
    pcamv = pcamv_loader.load()
    pcamv.host.load_library()
    t.append(time.perf_counter())
    ctx = pcamv.PcamvContext(1920, 1088, me_method=2, subpel_refine=5, max_refs=1)
    t.append(time.perf_counter())
    ctx2 = pcamv.PcamvContext(1920, 1088, me_method=2, subpel_refine=5, max_refs=1)
    t.append(time.perf_counter())
    names = ["load cudart", "first CUDA call (context)", "import + load libpcamv_cuda.so", "pcamv_open #1", "pcamv_open #2"]
    print(os.environ.get("CUDA_MODULE_LOADING", "(default)"), {n: round(b - a, 3) for n, a, b in zip(names, t, t[1:])})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for mode in (None, "LAZY", "EAGER"):
            env = dict(os.environ)
            if mode:
                env["CUDA_MODULE_LOADING"] = mode
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
