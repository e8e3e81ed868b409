#!/usr/bin/env python3
"""Build the reference's C host encoder with the CUDA shim bound in: host/_build/x264_pcamv.

The host side of this project IS the reference's own C code (encoder/analyse.c, encoder/me.c, encoder/encoder.c keep
their x264_encoder_encode and --emrate interface); this script compiles it from /root/reference with
  * the array widening every non-CIF resolution needs,
  * calls to host/pcamv_x264_glue.c inserted at the anchored lines INTEGRATION.md lists, and the three search entry
    points (x264_me_search_ref, x264_me_refine_qpel, x264_ih_get_mv_cost) replaced by the glue's replay versions,
and links it against video-steganography-pcamv_b200/libpcamv_cuda.so (rpath relative to the binary, so the pair
travels).  The scratch copy of the sources is deleted after the build; nothing of the reference is written into
tracked files.  Needs /root/reference (this container); the GPU box uses the prebuilt binary.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import reftree  # noqa: E402

OUT = os.path.join(HERE, "_build")
PKG = os.path.join(ROOT, "video-steganography-pcamv_b200")

HOOK_DECL = ("void pcamv_hook_open( x264_t *h ); void pcamv_hook_close( x264_t *h );\n"
             "void pcamv_hook_slice_begin( x264_t *h ); void pcamv_hook_slice_end( x264_t *h );\n"
             "void pcamv_hook_analyse_begin( x264_t *h ); void pcamv_hook_analyse_end( x264_t *h );\n"
             "void pcamv_hook_embed( x264_t *h, int an ); void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me );\n")
# encoder/analyse.c additions: a way for the glue to have the lambda*bits tables built before the first macroblock
# (the reference builds them lazily inside x264_mb_analyse_load_costs, analyse.c:193-229), and the replay version of
# x264_ih_get_mv_cost — the cost comes from the GPU log, the host state ends up where the reference leaves it
# (x264_analyse_update_cache with the original vector, analyse.c:2547-2548)
ANALYSE_EXTRA = ("struct x264_me_t_tag; int pcamv_glue_ih_cost( x264_t *h, x264_me_t *m, int16_t *m_x, int16_t *m_y );\n"
                 "void pcamv_glue_load_costs( x264_t *h, int qp )\n"
                 "{ x264_mb_analysis_t a; memset( &a, 0, sizeof(a) ); a.i_qp = qp; a.i_lambda = x264_lambda_tab[qp];\n"
                 "  x264_mb_analyse_load_costs( h, &a ); }\n")
IH_WRAPPER = ("{ int r = pcamv_glue_ih_cost( h, m, m_x, m_y ); x264_analyse_update_cache( h, analysis ); return r; }\n")


def main():
    if not os.path.isdir(reftree.REF):
        print("build_host: %s not present; keeping prebuilt host/_build/ as is" % reftree.REF)
        return 0
    lib = os.path.join(PKG, "libpcamv_cuda.so")
    if not os.path.exists(lib):
        raise SystemExit("build_host: %s is not built (run __graft_entry__.build())" % lib)
    os.makedirs(OUT, exist_ok=True)
    tree = os.path.join(OUT, "tree")
    reftree.copy_tree(tree)
    reftree.widen(tree)
    reftree.hook_call_sites(tree, HOOK_DECL, IH_WRAPPER, analyse_extra=ANALYSE_EXTRA)
    exe = os.path.join(OUT, "x264_pcamv")
    reftree.compile_tree(tree, exe,
                         extra_sources=[os.path.join(HERE, "ref_stub.c"), os.path.join(HERE, "pcamv_x264_glue.c")],
                         extra_cflags=["-I" + os.path.join(ROOT, "include")],
                         extra_ldflags=["-L" + PKG, "-lpcamv_cuda", "-Wl,-rpath,$ORIGIN/../../video-steganography-pcamv_b200"])
    shutil.rmtree(tree)
    print("build_host: built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
