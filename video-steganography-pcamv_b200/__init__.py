"""pcamv_b200 — B200 (sm_100a) implementation of the PCAMV encoder's motion-estimation hot path.

The directory is named ``video-steganography-pcamv_b200`` (not importable by name); load it with
``pcamv_loader.load()`` from the repo root, which registers it as the module ``pcamv_b200``.

Contents: ``csrc/`` (CUDA kernels + the C-ABI of libpcamv_cuda.so, declared in include/pcamv.h),
``host.py`` (ctypes mirror of that ABI — the binding a Python host would use; the reference's
host is C and binds the same symbols directly), ``dumpfmt.py`` (parser for the
instrumented-reference dumps used by tests and bench), ``build.py`` (nvcc recipes).
"""
from . import build, dumpfmt, host  # noqa: F401  (shard imports torch; load it on demand: from pcamv_b200 import shard)
from .host import PcamvContext, PcamvError, load_library  # noqa: F401
