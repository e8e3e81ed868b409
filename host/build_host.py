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
             "void pcamv_hook_embed( x264_t *h, int an ); void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me );\n"
             "void pcamv_hook_encoded( x264_t *h ); int pcamv_hook_pass1_on_device( x264_t *h );\n"
             "int pcamv_hook_skip_hpel( x264_t *h );\n")
# encoder/analyse.c additions: a way for the glue to have the lambda*bits tables built before the first macroblock
# (the reference builds them lazily inside x264_mb_analyse_load_costs, analyse.c:193-229), and the replay version of
# x264_ih_get_mv_cost — the cost comes from the GPU log, the host state ends up where the reference leaves it
# (x264_analyse_update_cache with the original vector, analyse.c:2547-2548)
ANALYSE_EXTRA = ("struct x264_me_t_tag; int pcamv_glue_ih_cost( x264_t *h, x264_me_t *m, int16_t *m_x, int16_t *m_y );\n"
                 "void pcamv_glue_load_costs( x264_t *h, int qp )\n"
                 "{ x264_mb_analysis_t a; memset( &a, 0, sizeof(a) ); a.i_qp = qp; a.i_lambda = x264_lambda_tab[qp];\n"
                 "  x264_mb_analyse_load_costs( h, &a ); }\n")
IH_WRAPPER = ("{ int r = pcamv_glue_ih_cost( h, m, m_x, m_y ); x264_analyse_update_cache( h, analysis ); return r; }\n")

# x264.c: `x264_pcamv --shards N --shard-frames K <x264 arguments>` encodes N IDR-bounded shards of K frames each
# (frames [g*K, (g+1)*K) = the CLI's own --seek / --frames) on N threads of one process, which share the
# GPU through an encoder group, and concatenates the shards' NAL streams in order.  Without --shards the CLI is unchanged.
SHARD_MAIN = r'''
/* ---- appended by host/build_host.py: GOP-sharded encoding on threads of one process -------------------------------- */
#include <pthread.h>
#include <unistd.h>
void pcamv_glue_set_shards( int n );
void pcamv_glue_shard_done( void );
void pcamv_glue_set_shard_index( int i, int gop );
typedef struct { x264_param_t param; cli_opt_t opt; int ret, index, gop; char out[1024]; } pcamv_shard_t;
/* `x264_pcamv --extract STEGO -o MESSAGE`: STEGO holds, per embedded frame in coding order, int32 frame, length, an and the
 * `length` stego LSBs (what the encoder writes with PCAMV_STEGO=<file>; a decoder-side MV parser would deliver the same
 * vector); MESSAGE receives int32 frame, an and the `an` recovered payload bits.  Frames that carried nothing (an <= 0) are
 * passed through with an = 0. */
int pcamv_stc_extract( const uint8_t *stego, int n, uint8_t *message, int an, int matrixheight );
int pcamv_stc_columns( int width, int height, uint32_t *out );
int pcamv_bitstream_main( int argc, char **argv );
static int pcamv_extract_main( int argc, char **argv )
{
    const char *in = argv[2], *out = NULL;
    FILE *fi, *fo;
    int32_t hd[3];
    int i, frames = 0, bits = 0, skipped = 0;
    for( i = 3; i < argc - 1; i++ )
        if( !strcmp( argv[i], "-o" ) || !strcmp( argv[i], "--output" ) ) out = argv[i + 1];
    if( !out ) { fprintf( stderr, "x264 [error]: --extract needs -o\n" ); return -1; }
    fi = fopen( in, "rb" ); fo = fopen( out, "wb" );
    if( !fi || !fo ) { fprintf( stderr, "x264 [error]: cannot open %s / %s\n", in, out ); return -1; }
    while( fread( hd, 4, 3, fi ) == 3 )
    {
        const int length = hd[1], an = hd[2] > 0 ? hd[2] : 0;
        uint8_t *stego = malloc( length > 0 ? length : 1 ), *msg = malloc( an + 1 );
        int32_t oh[2] = { hd[0], an };
        if( length < 0 || (int)fread( stego, 1, length, fi ) != length ) { fprintf( stderr, "x264 [error]: truncated stego file\n" ); return -1; }
        if( an > 0 && pcamv_stc_extract( stego, length, msg, an, 10 ) < 0 )
        {
            /* a frame that cannot carry what its header claims: pass it by (an = 0) instead of giving up on the whole stream */
            fprintf( stderr, "x264 [warning]: frame %d: cannot extract %d bits from %d carriers, skipped\n", hd[0], an, length );
            oh[1] = 0; skipped++;
        }
        fwrite( oh, 4, 2, fo );
        fwrite( msg, 1, oh[1], fo );
        free( stego ); free( msg );
        frames++; bits += oh[1];
    }
    fclose( fi ); fclose( fo );
    fprintf( stderr, "x264 [info]: extracted %d payload bits from %d frames (%d skipped)\n", bits, frames, skipped );
    return 0;
}
static void *pcamv_shard_thread( void *p )
{
    pcamv_shard_t *s = (pcamv_shard_t *)p;
    pcamv_glue_set_shard_index( s->index, s->gop );
    s->ret = Encode( &s->param, &s->opt );
    pcamv_glue_shard_done();
    return NULL;
}
int main( int argc, char **argv )
{
    int n, k, g, i, o_at = -1, ret = 0, first = 0, step = 1, keep = 0, a0 = 5;
    pcamv_shard_t *sh;
    pthread_t *th;
    FILE *fo;
    if( argc >= 3 && !strcmp( argv[1], "--extract" ) )
        return pcamv_extract_main( argc, argv );
    /* the decoder side (host/pcamv_bitstream.c): motion vectors / payload from the .264 alone */
    if( argc >= 3 && ( !strcmp( argv[1], "--parse-mv" ) || !strcmp( argv[1], "--extract-264" ) ) )
        return pcamv_bitstream_main( argc, argv );
    if( argc == 4 && !strcmp( argv[1], "--stc-columns" ) )
    {
        uint32_t cols[64]; int w = atoi( argv[2] ), hh = atoi( argv[3] ), i;
        if( w < 1 || w > 64 || pcamv_stc_columns( w, hh, cols ) ) return -1;
        for( i = 0; i < w; i++ ) printf( "%u\n", cols[i] );
        return 0;
    }
    if( argc < 6 || strcmp( argv[1], "--shards" ) || strcmp( argv[3], "--shard-frames" ) )
        return x264_cli_main( argc, argv );
    n = atoi( argv[2] ); k = atoi( argv[4] );
    if( n < 1 || k < 1 ) { fprintf( stderr, "x264 [error]: bad --shards / --shard-frames\n" ); return -1; }
    /* optional, right after --shard-frames K: --shard-first F --shard-step S (shard i encodes GOP F + i*S: the share of one
     * rank when the GOPs of a job are dealt round-robin to several processes / GPUs), --shard-keep (leave the per-GOP
     * streams "<out>.<gop>" in place for the launcher to concatenate) */
    while( a0 < argc - 1 )
    {
        if( !strcmp( argv[a0], "--shard-first" ) ) { first = atoi( argv[a0 + 1] ); a0 += 2; }
        else if( !strcmp( argv[a0], "--shard-step" ) ) { step = atoi( argv[a0 + 1] ); a0 += 2; }
        else if( !strcmp( argv[a0], "--shard-keep" ) ) { keep = 1; a0 += 1; }
        else break;
    }
    if( first < 0 || step < 1 ) { fprintf( stderr, "x264 [error]: bad --shard-first / --shard-step\n" ); return -1; }
    for( i = a0; i < argc - 1; i++ )
        if( !strcmp( argv[i], "-o" ) || !strcmp( argv[i], "--output" ) ) o_at = i + 1;
    if( o_at < 0 ) { fprintf( stderr, "x264 [error]: --shards needs -o\n" ); return -1; }
    sh = calloc( n, sizeof(*sh) ); th = calloc( n, sizeof(*th) );
    for( g = 0; g < n; g++ )
    {
        /* the shard's command line: the user's arguments, its own output file, --seek gop*K --frames K */
        char **av = calloc( argc + 8, sizeof(char *) ), seek[32], frames[32];
        int ac = 0;
        const int gop = first + g * step;
        av[ac++] = argv[0];
        snprintf( sh[g].out, sizeof(sh[g].out), "%s.%d", argv[o_at], gop );
        snprintf( seek, sizeof(seek), "%d", gop * k ); snprintf( frames, sizeof(frames), "%d", k );
        av[ac++] = "--seek"; av[ac++] = strdup( seek ); av[ac++] = "--frames"; av[ac++] = strdup( frames );
        for( i = a0; i < argc; i++ ) av[ac++] = i == o_at ? sh[g].out : argv[i];
        optind = 0;                                   /* getopt state is global: shards are parsed one after the other */
        x264_param_default( &sh[g].param );
        if( Parse( ac, av, &sh[g].param, &sh[g].opt ) < 0 ) return -1;
        sh[g].opt.b_progress = 0;
        sh[g].index = g;
        sh[g].gop = gop;
    }
    {
        /* the reference fills several global tables the first time an encoder opens (x264_rdo_init, x264_init_vlc_tables,
         * x264_dct_init_weights, the lambda*bits tables): open and close one encoder before any thread runs */
        x264_param_t warm = sh[0].param;
        x264_t *hw = x264_encoder_open( &warm );
        if( !hw ) return -1;
        x264_encoder_close( hw );
    }
    pcamv_glue_set_shards( n );
    signal( SIGINT, SigIntHandler );
    for( g = 0; g < n; g++ ) pthread_create( &th[g], NULL, pcamv_shard_thread, &sh[g] );
    for( g = 0; g < n; g++ ) { pthread_join( th[g], NULL ); ret |= sh[g].ret; }
    fo = fopen( argv[o_at], "wb" );
    if( !fo ) return -1;
    for( g = 0; g < n; g++ )
    {
        FILE *fi = fopen( sh[g].out, "rb" );
        char buf[1 << 16]; size_t r;
        if( !fi ) { ret = -1; continue; }
        while( ( r = fread( buf, 1, sizeof(buf), fi ) ) > 0 ) fwrite( buf, 1, r, fo );
        fclose( fi ); if( !keep ) remove( sh[g].out );
    }
    fclose( fo );
    return ret;
}
'''


def shard_driver(tree):
    p = os.path.join(tree, "x264.c")
    t = reftree.read(p)
    t = reftree.sub_exact(t, r"\nint main\( int argc, char \*\*argv \)\n", "\nint x264_cli_main( int argc, char **argv )\n", 1, "main")
    # the NAL staging buffer of the CLI is a global: one per encoder thread
    t = reftree.sub_exact(t, r"\nuint8_t \*mux_buffer = NULL;\nint mux_buffer_size = 0;", "\n__thread uint8_t *mux_buffer = NULL;\n__thread int mux_buffer_size = 0;", 1, "mux_buffer")
    t += SHARD_MAIN
    reftree.write(p, t)
    # payload bits: rand() & 1 (encoder/encoder.c:1838-1840) -> the same seed-1 stream, but per encoder thread
    p = os.path.join(tree, "encoder/encoder.c")
    t = reftree.read(p)
    t = reftree.sub_exact(t, r"h->info\.message\[i\] = rand\(\) & 0x01;", "h->info.message[i] = pcamv_tls_rand() & 0x01;", 1, "rand")
    t = t.replace("void pcamv_hook_open( x264_t *h );", "int pcamv_tls_rand( void ); void pcamv_hook_open( x264_t *h );", 1)
    # x264_encoder_close frees the process-wide lambda*bits tables (encoder/encoder.c:2944-2952) although the function-static
    # pointers in encoder/analyse.c:195 keep referring to them: with more than one encoder per process they must stay
    t = reftree.sub_exact(t, r"x264_free\(g_x264_cost_mv_fpel\[i\]\[j\]\);", ";", 1, "free fpel tables")
    t = reftree.sub_exact(t, r"x264_free\(g_cost_mv\[i\]\);", ";", 1, "free cost_mv tables")
    # the embedder's trellis runs on the GPU: the one stc_embed call of the embed stage (encoder/encoder.c:1843) goes to the
    # wrapper in host/pcamv_stc_extract.c (PCAMV_HOST_STC=1 at build time keeps the reference's CPU routine, for A/B timing)
    if not os.environ.get("PCAMV_HOST_STC"):
        t = reftree.sub_exact(t, r"\n(\s*)stc_embed\(h->info\.cover, h->info\.length, h->info\.message, an, h->info\.rho_final, h->info\.stego, 10\);",
                              r"\n\1pcamv_glue_stc_embed( h, an );", 1, "stc_embed call")
        t = t.replace("void pcamv_hook_open( x264_t *h );", "void pcamv_glue_stc_embed( x264_t *h, int an ); void pcamv_hook_open( x264_t *h );", 1)
    # the extractor the reference lacks (host/pcamv_stc_extract.c) joins the translation unit that owns getMatrix (embed.h)
    t += "\n" + open(os.path.join(HERE, "pcamv_stc_extract.c")).read()
    reftree.write(p, t)
    # the LCG behind the STC sub-matrices for widths outside the built-in tables (embed.h:134-139) keeps static state
    p = os.path.join(tree, "embed.h")
    t = reftree.read(p)
    t = reftree.sub_exact(t, r"static long myholdrand = 1L;", "static __thread long myholdrand = 1L;", 1, "myholdrand")
    reftree.write(p, t)


def stats_only_intra(tree):
    """Intra macroblocks are disabled in P slices (the COPY2_IF_LT that would pick them is commented out, encoder/analyse.c:2863;
    SURVEY fact 10): the intra analysis of a P macroblock (analyse.c:2812-2825) only feeds frame statistics — except that its
    predictions stay in the fdec scratch block, which the macroblocks of quirk q1 are coded against.  The host runs it for those
    (the GPU's records name them: early_skip == 2) and skips it for the others; PCAMV_HOST_INTRA=1 runs it everywhere."""
    p = os.path.join(tree, "encoder/analyse.c")
    t = reftree.read(p)
    t = reftree.sub_exact(t, r"\n            if\( h->mb\.b_chroma_me \)\n            \{\n([^\n]*\n)                x264_mb_analyse_intra_chroma\( h, &analysis \);",
                          r"\n            if( !pcamv_hook_want_intra( h ) ) analysis.i_satd_i16x16 = i_cost;\n            else if( h->mb.b_chroma_me )\n            {\n\1                x264_mb_analyse_intra_chroma( h, &analysis );",
                          1, "P-slice intra analysis")
    t = t.replace("void pcamv_hook_open( x264_t *h );", "int pcamv_hook_want_intra( x264_t *h ); void pcamv_hook_open( x264_t *h );", 1)
    reftree.write(p, t)


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
    reftree.hook_call_sites(tree, HOOK_DECL, IH_WRAPPER, analyse_extra=ANALYSE_EXTRA, drop_real=True, pass1_on_device=True)
    shard_driver(tree)
    stats_only_intra(tree)
    reftree.conformance_switch(tree, "int pcamv_conformant( void );\n")      # PCAMV_CONFORMANT=1, host/pcamv_x264_glue.c
    exe = os.path.join(OUT, "x264_pcamv")
    reftree.compile_tree(tree, exe,
                         extra_sources=[os.path.join(HERE, "ref_stub.c"), os.path.join(HERE, "pcamv_x264_glue.c"), os.path.join(HERE, "pcamv_bitstream.c")],
                         extra_cflags=["-I" + os.path.join(ROOT, "include")],
                         extra_ldflags=["-L" + PKG, "-lpcamv_cuda", "-Wl,-rpath,$ORIGIN/../../video-steganography-pcamv_b200"])
    print("build_host: built", exe)
    if "--profile" in sys.argv[1:]:
        # gprof twin (same flags + -pg): where a single-stream encode spends its HOST time; writes gmon.out into the cwd
        exe_pg = os.path.join(OUT, "x264_pcamv_pg")
        reftree.compile_tree(tree, exe_pg,
                             extra_sources=[os.path.join(HERE, "ref_stub.c"), os.path.join(HERE, "pcamv_x264_glue.c"), os.path.join(HERE, "pcamv_bitstream.c")],
                             extra_cflags=["-I" + os.path.join(ROOT, "include"), "-pg", "-fno-omit-frame-pointer"],
                             extra_ldflags=["-pg", "-L" + PKG, "-lpcamv_cuda", "-Wl,-rpath,$ORIGIN/../../video-steganography-pcamv_b200"])
        print("build_host: built", exe_pg)
    shutil.rmtree(tree)
    return 0


if __name__ == "__main__":
    sys.exit(main())
