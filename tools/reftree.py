"""Helpers shared by oracle/build_ref.py (the reference as checker) and host/build_host.py (the reference's C host
with the CUDA shim bound in): make a scratch copy of the reference's C sources, apply anchored edits there, compile
with the flags the reference's own `configure --disable-asm` produces.  No reference source is ever written into
tracked files; the scratch tree is deleted after the build."""
import os
import re
import shutil
import subprocess

REF = os.environ.get("PCAMV_REFERENCE", "/root/reference")
MAX_MB = 32400          # 3840x2160 = 240x135 macroblocks

SRCS = """common/mc.c common/predict.c common/pixel.c common/macroblock.c common/frame.c common/dct.c
common/cpu.c common/cabac.c common/common.c common/mdate.c common/set.c common/quant.c common/vlc.c
encoder/analyse.c encoder/me.c encoder/ratecontrol.c encoder/set.c encoder/macroblock.c encoder/cabac.c
encoder/cavlc.c encoder/encoder.c""".split()
SRCCLI = "x264.c matroska.c muxers.c".split()

CFLAGS = ("-O4 -ffast-math -Wall -I. -DHAVE_MALLOC_H -DARCH_X86_64 -DSYS_LINUX -DHAVE_PTHREAD "
          "-fomit-frame-pointer -Dsscanf_s=sscanf -D_strdup=strdup -w").split()
LDFLAGS = "-lm -lpthread".split()

CONFIG_H = '#define fseek fseeko\n#define ftell ftello\n#define X264_VERSION ""\n#define X264_POINTVER "0.66.x"\n'


def sub_exact(text, pattern, repl, count, what, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n != count:
        raise SystemExit("reftree: anchor %r matched %d times, expected %d" % (what, n, count))
    return new


def read(path):
    with open(path, "rb") as f:
        # byte-transparent (sources are GB18030); CRLF -> LF so the anchors are uniform
        return f.read().decode("latin-1").replace("\r\n", "\n")


def write(path, text):
    with open(path, "wb") as f:
        f.write(text.encode("latin-1"))


def widen(tree):
    """CIF constants -> MAX_MB macroblocks; leaves the CIF bitstream unchanged."""
    p = os.path.join(tree, "common/common.h")
    t = read(p)
    t = sub_exact(t, r"cache\[396\]", "cache[%d]" % MAX_MB, 1, "cache[396]")
    t = sub_exact(t, r"\[6336\]", "[%d]" % (16 * MAX_MB), 8, "[6336]")
    t = sub_exact(t, r"uint16_t i_mv_no", "uint32_t i_mv_no", 1, "i_mv_no")
    t = sub_exact(t, r"uint16_t num_mv_modify_real", "uint32_t num_mv_modify_real", 1, "num_mv_modify_real")
    write(p, t)
    p = os.path.join(tree, "encoder/encoder.c")
    t = read(p)
    t = sub_exact(t, r"i < 396;", "i < %d;" % MAX_MB, 1, "i < 396")
    t = sub_exact(t, r"\* 6336\)", "* %d)" % (16 * MAX_MB), 8, "* 6336)")
    write(p, t)


def copy_tree(dst):
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for d in ("common", "encoder", "extras"):
        os.makedirs(os.path.join(dst, d))
    for f in os.listdir(REF):
        if f.endswith((".c", ".h")):
            shutil.copy(os.path.join(REF, f), os.path.join(dst, f))
    for d in ("common", "encoder", "extras"):
        for f in os.listdir(os.path.join(REF, d)):
            if f.endswith((".c", ".h")):
                shutil.copy(os.path.join(REF, d, f), os.path.join(dst, d, f))
    for root, _, files in os.walk(dst):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    write(os.path.join(dst, "config.h"), CONFIG_H)


def compile_tree(tree, exe, extra_sources=(), extra_cflags=(), extra_ldflags=(), jobs=8, archive=None):
    """gcc every reference source of `tree` plus extra_sources, link `exe`; optionally ar the library objects."""
    objs, procs = [], []
    srcs = SRCS + SRCCLI

    def reap(s0, p0):
        _, err = p0.communicate()
        if p0.returncode:
            raise SystemExit("reftree: %s failed:\n%s" % (s0, err.decode("latin-1")[-4000:]))

    for s in srcs + list(extra_sources):
        o = os.path.join(tree, os.path.basename(s).replace(".c", "") + "_" + str(len(objs)) + ".o")
        objs.append(o)
        cmd = ["gcc"] + CFLAGS + list(extra_cflags) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, cwd=tree, stderr=subprocess.PIPE)))
        if len(procs) >= jobs:
            reap(*procs.pop(0))
    for s0, p0 in procs:
        reap(s0, p0)
    subprocess.check_call(["gcc", "-o", exe] + objs + LDFLAGS + list(extra_ldflags), cwd=tree)
    if archive:
        if os.path.exists(archive):
            os.remove(archive)
        libobjs = objs[:len(SRCS)] + objs[len(srcs):len(srcs) + 1]
        subprocess.check_call(["ar", "rcs", archive] + libobjs, cwd=tree)
    return exe


def hook_call_sites(tree, hook_decl, ih_wrapper_body, analyse_extra="", drop_real=False, pass1_on_device=False):
    """The anchored edits both instrumented builds share: calls to pcamv_hook_* at the frame-level points, the search
    entry points renamed *_real (the hook file defines the originals' names), and a wrapper in front of
    x264_ih_get_mv_cost.  (file:line of the reference in oracle/ref_hooks.c / host/pcamv_x264_glue.c.)
    drop_real: the GPU host never calls the reference's own searches, so there the renamed originals become unused statics
    and the compiler drops them (and refine_subpel with them) — `nm x264_pcamv` then shows no CPU search code at all; the
    instrumented oracle keeps them (its hooks time and count the real functions).
    pass1_on_device: x264_slice_write returns right after the slice-begin hook when the hook says the first pass of an embedding
    frame is complete without the host (int pcamv_hook_pass1_on_device( x264_t * )); x264_slices_write then finds no NAL unit of
    that pass and must not index nal[-1] (encoder/encoder.c:2091)."""
    p = os.path.join(tree, "encoder/encoder.c")
    t = read(p)
    t = sub_exact(t, r'(#include "common/common.h"\n)', r"\1" + hook_decl.replace("\\", "\\\\"), 1, "encoder.c include")
    # x264_encoder_open: the first mbcmp_init( h ) call (encoder/encoder.c:766); the second is reconfig
    idx = t.index("    mbcmp_init( h );")
    t = t[:idx] + "    mbcmp_init( h ); pcamv_hook_open( h );" + t[idx + len("    mbcmp_init( h );"):]
    t = sub_exact(t, r"(    /\* init stats \*/\n    memset\( &h->stat\.frame, 0, sizeof\(h->stat\.frame\) \);)",
                  r"\1 pcamv_hook_slice_begin( h );" + (" if( pcamv_hook_pass1_on_device( h ) ) return;" if pass1_on_device else ""), 1, "slice begin")
    if pass1_on_device:
        # x264_fdec_filter_row (encoder/encoder.c:1041-1046): the half-pel planes of a frame the GPU rebuilds as a reference come
        # back from the GPU (pcamv_hook_slice_end), the host does not filter them itself
        t = sub_exact(t, r"(        if\( h->param\.analyse\.i_subpel_refine) \)\n(        \{\n[^\n]*\n            x264_frame_filter\( h, h->fdec, min_y, b_end \);)",
                      r"\1 && !pcamv_hook_skip_hpel( h ) )\n\2", 1, "half-pel filter of fdec")
        t = sub_exact(t, r"i_frame_size = h->out\.nal\[h->out\.i_nal-1\]\.i_payload;",
                      "i_frame_size = h->out.i_nal ? h->out.nal[h->out.i_nal-1].i_payload : 0;", 1, "frame size of an empty pass")
    t = sub_exact(t, r"\n(\t\tx264_macroblock_analyse\( h \);)",
                  r"\n\t\tpcamv_hook_analyse_begin( h ); x264_macroblock_analyse( h ); pcamv_hook_analyse_end( h );", 1, "analyse call")
    # right after the macroblock has been reconstructed in h->mb.pic.p_fdec (encoder/encoder.c:1881), before anything filters it
    t = sub_exact(t, r"\n(\t\tx264_macroblock_encode\( h \);)", r"\n\1 pcamv_hook_encoded( h );", 1, "macroblock_encode call")
    # after the filp loop of the embed stage (encoder/encoder.c:1848-1855): hook before the DEGUG print
    t = sub_exact(t, r"(\t\t\t\t// [^\n]*\n\t\t\t\tif \(DEGUG_LIJUN\)\n\t\t\t\t\{\n\t\t\t\t\tprintf\(\"1)",
                  r"\t\t\t\tpcamv_hook_embed( h, an );\n\1", 1, "embed end")
    # end of x264_slice_write: the MB loop is followed by the cabac flush
    t = sub_exact(t, r"(\n    if\( h->param\.b_cabac \)[^\n]*\n    \{\n        x264_cabac_encode_flush\( h, &h->cabac \);)",
                  r"\n    pcamv_hook_slice_end( h );\1", 1, "slice end")
    t = sub_exact(t, r"(void    x264_encoder_close  \( x264_t \*h \)\n\{)", r"\1 pcamv_hook_close( h );", 1, "close")
    write(p, t)

    p = os.path.join(tree, "encoder/me.c")
    t = read(p)
    lead = "\nstatic __attribute__((unused)) void " if drop_real else "\nvoid "
    t = sub_exact(t, r"\nvoid x264_me_search_ref\(", lead + "x264_me_search_ref_real(", 1, "me_search_ref def")
    t = sub_exact(t, r"\nvoid x264_me_refine_qpel\(", lead + "x264_me_refine_qpel_real(", 1, "me_refine_qpel def")
    write(p, t)

    p = os.path.join(tree, "encoder/analyse.c")
    t = read(p)
    t = sub_exact(t, r'(#include "common/common.h"\n)', r"\1" + hook_decl.replace("\\", "\\\\"), 1, "analyse.c include")
    t = sub_exact(t, r"(#define MV_SATD_FDEC_IH\(mx, my\)\\\n\{\\\n)", r"\1\tpcamv_hook_ih_satd( m->i_pixel, h->mb.b_chroma_me && m->i_pixel <= PIXEL_8x8 );\\\n", 1, "MV_SATD_FDEC_IH")
    # x264_ih_get_mv_cost (encoder/analyse.c:2391): rename the definition and put a wrapper of the same name in front of
    # x264_macroblock_analyse (its only caller, encoder/analyse.c:3557-3673)
    t = sub_exact(t, r"\nstatic inline int x264_ih_get_mv_cost\(", "\nstatic inline __attribute__((unused)) int x264_ih_get_mv_cost_real(", 1, "ih_get_mv_cost def")
    wrapper = ("static int x264_ih_get_mv_cost( x264_t *h, x264_mb_analysis_t *analysis, x264_me_t *m, int16_t *m_x, int16_t *m_y,\n"
               "    int8_t d_mv[][2], int8_t d_mv_1_neighborhood[][2], int mb_xy )\n" + ih_wrapper_body)
    idx = t.index("\nvoid x264_macroblock_analyse( x264_t *h )\n")
    t = t[:idx] + "\n" + analyse_extra + wrapper + t[idx:]
    write(p, t)


def conformance_switch(tree, switch_def):
    """Three anchored edits in encoder/analyse.c behind `int pcamv_conformant( void )` (`switch_def` declares or defines it).
    With the switch OFF the code is the reference's, statement for statement.  With it ON pass 2 of an embedding P frame
    writes a stream a standard H.264 decoder reads back to the encoder's own vectors, so the payload can be extracted from
    the .264 alone (host/pcamv_bitstream.c; SURVEY.md 8(f) row 4 asks for "a documented fix/flag ... off by default"):
      1. the copy of the macroblock's vectors into info.cache[].mv (analyse.c:3537-3543, 3626-3632) is a straight copy instead
         of the unsequenced `idx++` one, so a partition's cover bit is the LSB of its own vector;
      2. a macroblock pass 2 forces to P_8x8 also gets h->mb.i_partition = D_8x8 (analyse.c:2871-2875 sets it for P_L0 only;
         x264_mb_predict_mv, common/macroblock.c:51-79, applies the 16x8 / 8x16 shortcuts to the 8x8 blocks of a macroblock
         whose i_partition is still what pass 2's own analysis chose — the decoder predicts differently and reads other vectors);
      3. a macroblock pass 2 forces to P_SKIP leaves through x264_analyse_update_cache like every other skip (analyse.c:2677-2680
         returns with the vector cache of whatever was analysed last: the encoder then motion-compensates and predicts its
         neighbours from a vector the decoder never sees — quirk q2)."""
    p = os.path.join(tree, "encoder/analyse.c")
    t = read(p)
    t = sub_exact(t, r'(#include "common/common.h"\n)', r"\1" + switch_def.replace("\\", "\\\\"), 1, "analyse.c include (conformance)")
    t = sub_exact(t, r"for \(int idx = 0; idx < 16;\) \{",
                  "for (int idx = 0; idx < 16;) { if( pcamv_conformant() ) { *(uint32_t*)&h->info.cache[mb_xy].mv[idx][0] = "
                  "*(uint32_t*)&h->mb.cache.mv[0][x264_scan8[idx]][0]; idx++; continue; }", 2, "unsequenced mv copy")
    t = sub_exact(t, r"(\n\t\t\t\tif \(i_type==P_L0\)\n\t\t\t\t\{\n\t\t\t\t\th->mb\.i_partition = h->info\.cache\[mb_xy\]\.i_partition;)",
                  r"\n\t\t\t\tif( i_type == P_8x8 && pcamv_conformant() ) h->mb.i_partition = D_8x8;\1", 1, "forced P_8x8 partition")
    t = sub_exact(t, r"(\n\t\t\tif \(h->mb\.i_type == P_SKIP\)[^\n]*\n\t\t\t\{\n)(\t\t\t\treturn;)",
                  r"\1\t\t\t\tif( pcamv_conformant() && h->info.embed_flag && !h->info.firstTime ) x264_analyse_update_cache( h, &analysis );\n\2",
                  1, "forced P_SKIP cache")
    write(p, t)


def rd_hook(tree):
    """oracle variant x264_dump_rd: x264_rd_cost_mb (encoder/rdo.c:139-172, #included by encoder/analyse.c) reports every macroblock it
    sizes with CAVLC to oracle/ref_hooks.c::pcamv_hook_rd_mb - the inputs and the result of x264_macroblock_size_cavlc."""
    p = os.path.join(tree, "encoder/rdo.c")
    t = read(p)
    t = sub_exact(t, r"(        x264_macroblock_size_cavlc\( h, &bs_tmp \);\n)",
                  r"\1        pcamv_hook_rd_mb( h, i_ssd, bs_tmp.i_bits_encoded, i_lambda2 );\n", 1, "rd_cost_mb cavlc size")
    t = "void pcamv_hook_rd_mb( x264_t *h, int i_ssd, int i_bits_encoded, int i_lambda2 );\n" + t
    write(p, t)


def intra_hook(tree):
    """oracle variant x264_dump_rd: x264_mb_analyse_intra (encoder/analyse.c:628-879) is wrapped; after every call the wrapper reports
    what the analysis saw and decided to oracle/ref_hooks.c::pcamv_hook_intra."""
    p = os.path.join(tree, "encoder/analyse.c")
    t = read(p)
    t = sub_exact(t, r"\nstatic void x264_mb_analyse_intra\( x264_t \*h, x264_mb_analysis_t \*a, int i_satd_inter \)\n\{",
                  "\nstatic void x264_mb_analyse_intra_real( x264_t *h, x264_mb_analysis_t *a, int i_satd_inter )\n{", 1, "analyse_intra def")
    wrapper = ("void pcamv_hook_intra( x264_t *h, int lambda, int i_satd_inter, int satd16, int pred16, const int *dir16, int satd_c, int pred_c, int satd4, const int *pred4,\n"
               "                       int i_qp, int i_mbrd, int b_fast_intra, int i_satd_i8x8 );\n"
               "static void x264_mb_analyse_intra( x264_t *h, x264_mb_analysis_t *a, int i_satd_inter )\n"
               "{\n    x264_mb_analyse_intra_real( h, a, i_satd_inter );\n"
               "    pcamv_hook_intra( h, a->i_lambda, i_satd_inter, a->i_satd_i16x16, a->i_predict16x16, a->i_satd_i16x16_dir, a->i_satd_i8x8chroma,\n"
               "                      a->i_predict8x8chroma, a->i_satd_i4x4, a->i_predict4x4, a->i_qp, a->i_mbrd, a->b_fast_intra, a->i_satd_i8x8 );\n}\n\n")
    t = sub_exact(t, r"\nstatic void x264_intra_rd\( x264_t \*h, x264_mb_analysis_t \*a, int i_satd_thresh \)\n", "\n" + wrapper.replace("\\", "\\\\") + "static void x264_intra_rd( x264_t *h, x264_mb_analysis_t *a, int i_satd_thresh )\n", 1, "intra wrapper")
    write(p, t)
