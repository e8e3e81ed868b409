"""nvcc / gcc build recipes (in-tree artefacts; they travel to the GPU box with the snapshot)."""
import os
import subprocess
import sys
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpcamv_cuda.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def cuda_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "pcamv.h"))
    return deps


def build_cuda(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into libpcamv_cuda.so (nvcc cross-compiles without a GPU)."""
    if not force and _newer(LIB, cuda_deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + cuda_sources()
    subprocess.check_call(cmd)
    return LIB


_tool_lock = threading.Lock()


def build_tool(name, src, extra=(), deps=None):
    """gcc/g++ helper for the small host tools (synth generator, emulation checkers).  `deps`: the files the tool is built
    from besides `src` (default: everything under csrc/, which the emulation checkers include).  Safe to call from several
    threads: one build at a time, and the binary is replaced atomically (a running copy is never written to)."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name)
    if deps is None:
        deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    with _tool_lock:
        if _newer(out, [src] + list(deps)):
            return out
        cc = "g++" if src.endswith(".cpp") else "gcc"
        std = ["-std=c++17"] if cc == "g++" else []
        tmp = "%s.tmp.%d" % (out, os.getpid())
        subprocess.check_call([cc, "-O2", "-w"] + std + ["-o", tmp, src] + list(extra))
        os.replace(tmp, out)
    return out


def build_synth():
    return build_tool("pcamv_synth", os.path.join(ROOT, "synth", "pcamv_synth.c"), deps=[])


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
    print(build_synth())
