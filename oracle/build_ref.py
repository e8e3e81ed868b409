#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): build the reference encoder from its own sources.

Compiles the C sources where they lie under /root/reference (read-only) into oracle/_ref/:

  x264_ref    unmodified sources (CIF-only: the reference hard-wires 396 macroblocks,
              common/common.h:603-619)
  x264_wide   same, with the CIF constants widened (396 -> MAX_MB, 6336 -> 16*MAX_MB,
              uint16 MV counters -> uint32) so 720p/1080p/4K run
  x264_dump   x264_wide + oracle/ref_hooks.c instrumentation (dumps, counters, timers)
  x264_dump_conformant   x264_dump with tools/reftree.py::conformance_switch ON: the three statements of pass 2 that make the
              reference's embedding streams unreadable for a standard decoder are corrected (straight vector copy,
              i_partition of a forced P_8x8, vector cache of a forced P_SKIP).  NOT the parity reference — its bitstream
              differs from the reference's by design; it is the checker for the decoder side (tests/test_bitstream.py)
  x264_dump_rd   x264_dump + a report from x264_rd_cost_mb (encoder/rdo.c:139-172): what x264_macroblock_size_cavlc saw and returned
              for every macroblock RD mode decision sized (--subme 6 --no-cabac); the checker of csrc/pcamv_cavlc.cuh

The reference does not link as shipped: the MSVC-only `sscanf_s`/`_strdup`
are mapped with -D, and the un-vendored S-UNIWARD.lib symbol comes from oracle/ref_stub.c.
Flags are those the reference's `configure --disable-asm` produces on x86-64 Linux
(configure:205-209,397-410): -O4 -ffast-math -fomit-frame-pointer, C only.

The reference's own build system is not run.  A scratch copy of the needed sources is made
under oracle/_ref/build/ (git-ignored), patched there, compiled, and deleted again; no
reference source is ever written into tracked files.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import reftree  # noqa: E402
from reftree import REF  # noqa: E402

OUT = os.path.join(HERE, "_ref")

HOOK_DECL = ("void pcamv_hook_open( x264_t *h ); void pcamv_hook_close( x264_t *h );\n"
             "void pcamv_hook_slice_begin( x264_t *h ); void pcamv_hook_slice_end( x264_t *h );\n"
             "void pcamv_hook_analyse_begin( x264_t *h ); void pcamv_hook_analyse_end( x264_t *h );\n"
             "void pcamv_hook_embed( x264_t *h, int an ); void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me );\n"
             "void pcamv_hook_encoded( x264_t *h );\n"
             "void pcamv_hook_ih_begin( void ); void pcamv_hook_ih_end( void );\n")
# timing wrapper around the cost-table routine
IH_WRAPPER = ("{ int r; pcamv_hook_ih_begin(); r = x264_ih_get_mv_cost_real( h, analysis, m, m_x, m_y, d_mv, d_mv_1_neighborhood, mb_xy );\n"
              "  pcamv_hook_ih_end(); return r; }\n")


def compile_variant(name, wide, hooks, jobs=8, conformant=False, rd=False):
    tree = os.path.join(OUT, "build", name)
    reftree.copy_tree(tree)
    if wide:
        reftree.widen(tree)
    if hooks:
        reftree.hook_call_sites(tree, HOOK_DECL, IH_WRAPPER)
    if conformant:
        reftree.conformance_switch(tree, "static inline int pcamv_conformant( void ) { return 1; }\n")
    if rd:
        reftree.rd_hook(tree)
        reftree.intra_hook(tree)
    extra = [os.path.join(HERE, "ref_stub.c")] + ([os.path.join(HERE, "ref_hooks.c")] if hooks else [])
    exe = os.path.join(OUT, name)
    # libx264-equivalent archive of the wide build, for leaf-level differential tests
    reftree.compile_tree(tree, exe, extra_sources=extra, jobs=jobs,
                         archive=os.path.join(OUT, "libx264_wide.a") if name == "x264_wide" else None)
    shutil.rmtree(tree)
    return exe


def main():
    if not os.path.isdir(REF):
        print("build_ref: %s not present; keeping prebuilt oracle/_ref/ as is" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    want = sys.argv[1:] or ["x264_ref", "x264_wide", "x264_dump", "x264_dump_conformant", "x264_dump_rd"]
    for name in want:
        exe = compile_variant(name, wide=name != "x264_ref", hooks=name.startswith("x264_dump"), conformant=name.endswith("_conformant"),
                              rd=name.endswith("_rd"))
        print("build_ref: built", exe)
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
