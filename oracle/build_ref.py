#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): build the reference encoder from its own sources.

Compiles the C sources where they lie under /root/reference (read-only) into oracle/_ref/:

  x264_ref    unmodified sources (CIF-only: the reference hard-wires 396 macroblocks,
              common/common.h:603-619)
  x264_wide   same, with the CIF constants widened (396 -> MAX_MB, 6336 -> 16*MAX_MB,
              uint16 MV counters -> uint32) so 720p/1080p/4K run
  x264_dump   x264_wide + oracle/ref_hooks.c instrumentation (dumps, counters, timers)

The reference does not link as shipped: the MSVC-only `sscanf_s`/`_strdup`
are mapped with -D, and the un-vendored S-UNIWARD.lib symbol comes from oracle/ref_stub.c.
Flags are those the reference's `configure --disable-asm` produces on x86-64 Linux
(configure:205-209,397-410): -O4 -ffast-math -fomit-frame-pointer, C only.

The reference's own build system is not run.  A scratch copy of the needed sources is made
under oracle/_ref/build/ (git-ignored), patched there, compiled, and deleted again; no
reference source is ever written into tracked files.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCAMV_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
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
        raise SystemExit("build_ref: anchor %r matched %d times, expected %d" % (what, n, count))
    return new


def read(path):
    with open(path, "rb") as f:
        # byte-transparent (sources are GB18030); CRLF -> LF so the anchors below are uniform
        return f.read().decode("latin-1").replace("\r\n", "\n")


def write(path, text):
    with open(path, "wb") as f:
        f.write(text.encode("latin-1"))


def widen(tree):
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


def instrument(tree):
    hook_decl = ("void pcamv_hook_open( x264_t *h ); void pcamv_hook_close( x264_t *h );\n"
                 "void pcamv_hook_slice_begin( x264_t *h ); void pcamv_hook_slice_end( x264_t *h );\n"
                 "void pcamv_hook_analyse_begin( x264_t *h ); void pcamv_hook_analyse_end( x264_t *h );\n"
                 "void pcamv_hook_embed( x264_t *h, int an ); void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me );\n"
                 "void pcamv_hook_ih_begin( void ); void pcamv_hook_ih_end( void );\n")
    p = os.path.join(tree, "encoder/encoder.c")
    t = read(p)
    t = sub_exact(t, r'(#include "common/common.h"\n)', r"\1" + hook_decl.replace("\\", "\\\\"), 1, "encoder.c include")
    # x264_encoder_open: the first mbcmp_init( h ) call (encoder/encoder.c:766); the second is reconfig
    idx = t.index("    mbcmp_init( h );")
    t = t[:idx] + "    mbcmp_init( h ); pcamv_hook_open( h );" + t[idx + len("    mbcmp_init( h );"):]
    t = sub_exact(t, r"(    /\* init stats \*/\n    memset\( &h->stat\.frame, 0, sizeof\(h->stat\.frame\) \);)",
                  r"\1 pcamv_hook_slice_begin( h );", 1, "slice begin")
    t = sub_exact(t, r"\n(\t\tx264_macroblock_analyse\( h \);)",
                  r"\n\t\tpcamv_hook_analyse_begin( h ); x264_macroblock_analyse( h ); pcamv_hook_analyse_end( h );", 1, "analyse call")
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
    t = sub_exact(t, r"\nvoid x264_me_search_ref\(", "\nvoid x264_me_search_ref_real(", 1, "me_search_ref def")
    t = sub_exact(t, r"\nvoid x264_me_refine_qpel\(", "\nvoid x264_me_refine_qpel_real(", 1, "me_refine_qpel def")
    write(p, t)

    p = os.path.join(tree, "encoder/analyse.c")
    t = read(p)
    t = sub_exact(t, r'(#include "common/common.h"\n)', r"\1" + hook_decl.replace("\\", "\\\\"), 1, "analyse.c include")
    t = sub_exact(t, r"(#define MV_SATD_FDEC_IH\(mx, my\)\\\n\{\\\n)", r"\1\tpcamv_hook_ih_satd( m->i_pixel, h->mb.b_chroma_me && m->i_pixel <= PIXEL_8x8 );\\\n", 1, "MV_SATD_FDEC_IH")
    # time spent inside x264_ih_get_mv_cost (encoder/analyse.c:2391): rename the definition and put a timing wrapper
    # of the same name in front of x264_macroblock_analyse (its only caller, encoder/analyse.c:3557-3673)
    t = sub_exact(t, r"\nstatic inline int x264_ih_get_mv_cost\(", "\nstatic inline int x264_ih_get_mv_cost_real(", 1, "ih_get_mv_cost def")
    wrapper = ("static int x264_ih_get_mv_cost( x264_t *h, x264_mb_analysis_t *analysis, x264_me_t *m, int16_t *m_x, int16_t *m_y,\n"
               "    int8_t d_mv[][2], int8_t d_mv_1_neighborhood[][2], int mb_xy )\n"
               "{ int r; pcamv_hook_ih_begin(); r = x264_ih_get_mv_cost_real( h, analysis, m, m_x, m_y, d_mv, d_mv_1_neighborhood, mb_xy );\n"
               "  pcamv_hook_ih_end(); return r; }\n")
    idx = t.index("\nvoid x264_macroblock_analyse( x264_t *h )\n")
    t = t[:idx] + "\n" + wrapper + t[idx:]
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


def compile_variant(name, wide, hooks, jobs=8):
    tree = os.path.join(OUT, "build", name)
    copy_tree(tree)
    if wide:
        widen(tree)
    if hooks:
        instrument(tree)
    objs = []
    procs = []
    srcs = SRCS + SRCCLI
    extra = [os.path.join(HERE, "ref_stub.c")] + ([os.path.join(HERE, "ref_hooks.c")] if hooks else [])
    for s in srcs + extra:
        o = os.path.join(tree, os.path.basename(s).replace(".c", "") + "_" + str(len(objs)) + ".o")
        objs.append(o)
        cmd = ["gcc"] + CFLAGS + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, cwd=tree, stderr=subprocess.PIPE)))
        if len(procs) >= jobs:
            s0, p0 = procs.pop(0)
            _, err = p0.communicate()
            if p0.returncode:
                raise SystemExit("build_ref: %s failed:\n%s" % (s0, err.decode("latin-1")[-4000:]))
    for s0, p0 in procs:
        _, err = p0.communicate()
        if p0.returncode:
            raise SystemExit("build_ref: %s failed:\n%s" % (s0, err.decode("latin-1")[-4000:]))
    exe = os.path.join(OUT, name)
    subprocess.check_call(["gcc", "-o", exe] + objs + LDFLAGS, cwd=tree)
    # libx264-equivalent archive of the wide build, for leaf-level differential tests
    if name == "x264_wide":
        lib = os.path.join(OUT, "libx264_wide.a")
        if os.path.exists(lib):
            os.remove(lib)
        libobjs = objs[:len(SRCS)] + [objs[len(srcs)]]
        subprocess.check_call(["ar", "rcs", lib] + libobjs, cwd=tree)
    shutil.rmtree(tree)
    return exe


def main():
    if not os.path.isdir(REF):
        print("build_ref: %s not present; keeping prebuilt oracle/_ref/ as is" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    want = sys.argv[1:] or ["x264_ref", "x264_wide", "x264_dump"]
    for name in want:
        exe = compile_variant(name, wide=name != "x264_ref", hooks=name == "x264_dump")
        print("build_ref: built", exe)
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
