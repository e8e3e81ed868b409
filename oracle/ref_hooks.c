/* TEST INFRASTRUCTURE ONLY (oracle/): instrumentation linked into the "x264_dump" twin of the
 * reference encoder (see oracle/build_ref.py).  This file is ours; it includes the reference's
 * headers from /root/reference at build time and never ships reference code.
 *
 * What it records (all switched by environment variables, off by default):
 *   PCAMV_DUMP=<file>        tagged binary records, see the layout comments below
 *   PCAMV_DUMP_PLANES=1      also write fenc + reference planes at every P-slice begin
 *   PCAMV_DUMP_CALLS=0       suppress per-call search records (frame-level records stay)
 *   PCAMV_DUMP_FRAMES=a:b    only frames a <= h->i_frame < b get records
 *   PCAMV_COUNT=1            count block-distortion evaluations made inside the ME path
 *   PCAMV_STATS=<file>       one JSON object with counters and timers, written at close
 *
 * Hook call sites (inserted by build_ref.py at anchored lines of the build copy):
 *   pcamv_hook_open          after mbcmp_init() in x264_encoder_open   (encoder/encoder.c:766)
 *   pcamv_hook_slice_begin   top of the MB loop setup in x264_slice_write (encoder/encoder.c:1176)
 *   pcamv_hook_analyse_begin/_end  around x264_macroblock_analyse(h)    (encoder/encoder.c:1273)
 *   pcamv_hook_embed         after filp[] is final                      (encoder/encoder.c:1855)
 *   pcamv_hook_slice_end     after the MB loop of x264_slice_write
 *   pcamv_hook_close         in x264_encoder_close                      (encoder/encoder.c:2670)
 *   x264_me_search_ref / x264_me_refine_qpel are wrapped (the real ones are renamed *_real)
 *   pcamv_hook_ih_satd       inside MV_SATD_FDEC_IH                      (encoder/analyse.c:2364)
 */
#include "common/common.h"
#include "encoder/me.h"
#include "encoder/macroblock.h"
#include <time.h>

void x264_me_search_ref_real( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh );
void x264_me_refine_qpel_real( x264_t *h, x264_me_t *m );

static FILE *g_dump;
static int g_dump_planes, g_dump_calls = 1, g_count;
static int g_frame_lo = 0, g_frame_hi = 0x7fffffff;
static const char *g_stats_path;

static int g_in_me;                 /* inside x264_me_search_ref / x264_me_refine_qpel */
static int g_pass;                  /* 1 = first (pre-)encode of a P frame, 2 = second, 0 = no embedding */
static int g_pslice;
static uint64_t g_cnt_sad, g_cnt_satd, g_cnt_ih_luma, g_cnt_ih_chroma;
static uint64_t g_cnt_search, g_cnt_refine, g_cnt_mb_p, g_cnt_pframes, g_cnt_ads;
static uint64_t g_pix_sad, g_pix_satd;  /* pixel-normalised (w*h) */
/* further work of the P-slice analysis, pixel-normalised, for the integer-op roofline (SURVEY.md 8(d) weights): quarter-pel
 * averaging in get_ref / mc_luma (2 ops/pixel), chroma bilinear MC (8 ops/pixel), DCT + quant + dequant + IDCT of the
 * macroblock reconstructions and P_SKIP probes (14 ops/pixel).  Counted only inside x264_macroblock_analyse of P slices. */
static uint64_t g_pix_avg, g_pix_chroma_mc, g_pix_dct;
static uint64_t g_ih_pix_avg, g_ih_pix_chroma_mc, g_ih_pix_dct, g_ih_pix_satd;     /* the share of x264_ih_get_mv_cost (the cost-table kernel) */
static int g_in_analyse_p, g_in_ih;
static double g_t_analyse_p, g_t_me, g_t_ih, g_t_total0;
static uint64_t g_cnt_ih_calls;
static struct timespec g_ts_an, g_ts_me, g_ts_ih;
static uint64_t g_slice_c0[9];      /* counters at the start of the current slice pass */
static uint64_t g_slice_x0[7];      /* ... and of pix_avg, pix_chroma_mc, pix_dct, and the cost table's share of avg / chroma / dct / satd */
static double g_slice_t0[3];

static double now_s( void )
{
    struct timespec ts; clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + 1e-9*ts.tv_nsec;
}

static int dump_on( x264_t *h )
{
    return g_dump && h->i_frame >= g_frame_lo && h->i_frame < g_frame_hi;
}

static void rec_begin( const char tag[4], uint32_t nbytes )
{
    fwrite( tag, 1, 4, g_dump );
    fwrite( &nbytes, 4, 1, g_dump );
}

/* ---- counting wrappers around the pixel-compare tables the ME path uses ---------------------- */
static x264_pixel_function_t g_orig;
static const int g_pw[7] = {16,16,8,8,8,4,4}, g_ph[7] = {16,8,16,8,4,8,4};

#define WRAP1(tab, idx, kind) \
static int w_##tab##_##idx( uint8_t *a, int sa, uint8_t *b, int sb ) \
{ if( g_in_me ) { g_cnt_##kind++; g_pix_##kind += g_pw[idx]*g_ph[idx]; } return g_orig.tab[idx]( a, sa, b, sb ); }
#define WRAPX3(tab, idx, kind) \
static void w_##tab##_##idx( uint8_t *f, uint8_t *p0, uint8_t *p1, uint8_t *p2, int s, int sc[3] ) \
{ if( g_in_me ) { g_cnt_##kind += 3; g_pix_##kind += 3*g_pw[idx]*g_ph[idx]; } g_orig.tab[idx]( f, p0, p1, p2, s, sc ); }
#define WRAPX4(tab, idx, kind) \
static void w_##tab##_##idx( uint8_t *f, uint8_t *p0, uint8_t *p1, uint8_t *p2, uint8_t *p3, int s, int sc[4] ) \
{ if( g_in_me ) { g_cnt_##kind += 4; g_pix_##kind += 4*g_pw[idx]*g_ph[idx]; } g_orig.tab[idx]( f, p0, p1, p2, p3, s, sc ); }
#define WRAP7(M, tab, kind) M(tab,0,kind) M(tab,1,kind) M(tab,2,kind) M(tab,3,kind) M(tab,4,kind) M(tab,5,kind) M(tab,6,kind)

/* fpelcmp is SAD unless --me tesa; mbcmp is SATD when subme > 1 (encoder/encoder.c:615-625).
 * They are counted by what the table actually holds, decided at install time. */
WRAP7(WRAP1, sad, sad)
WRAP7(WRAP1, satd, satd)
WRAP7(WRAPX3, sad_x3, sad)
WRAP7(WRAPX4, sad_x4, sad)
WRAP7(WRAPX3, satd_x3, satd)
WRAP7(WRAPX4, satd_x4, satd)

static x264_mc_functions_t g_orig_mc;
static x264_dct_function_t g_orig_dct;
static void w_mc_luma( uint8_t *dst, int i_dst, uint8_t **src, int i_src, int mvx, int mvy, int w, int hh )
{
    if( g_in_analyse_p && ( ((mvy&3)<<2) + (mvx&3) ) & 5 ) { g_pix_avg += w*hh; if( g_in_ih ) g_ih_pix_avg += w*hh; }      /* the positions mc.c:206,232 average */
    g_orig_mc.mc_luma( dst, i_dst, src, i_src, mvx, mvy, w, hh );
}
static uint8_t *w_get_ref( uint8_t *dst, int *i_dst, uint8_t **src, int i_src, int mvx, int mvy, int w, int hh )
{
    if( g_in_analyse_p && ( ((mvy&3)<<2) + (mvx&3) ) & 5 ) { g_pix_avg += w*hh; if( g_in_ih ) g_ih_pix_avg += w*hh; }
    return g_orig_mc.get_ref( dst, i_dst, src, i_src, mvx, mvy, w, hh );
}
static void w_mc_chroma( uint8_t *dst, int i_dst, uint8_t *src, int i_src, int mvx, int mvy, int w, int hh )
{
    if( g_in_analyse_p ) { g_pix_chroma_mc += w*hh; if( g_in_ih ) g_ih_pix_chroma_mc += w*hh; }
    g_orig_mc.mc_chroma( dst, i_dst, src, i_src, mvx, mvy, w, hh );
}
static void w_sub16x16_dct( int16_t dct[16][4][4], uint8_t *a, uint8_t *b ) { if( g_in_analyse_p ) { g_pix_dct += 256; if( g_in_ih ) g_ih_pix_dct += 256; } g_orig_dct.sub16x16_dct( dct, a, b ); }
static void w_sub8x8_dct( int16_t dct[4][4][4], uint8_t *a, uint8_t *b ) { if( g_in_analyse_p ) { g_pix_dct += 64; if( g_in_ih ) g_ih_pix_dct += 64; } g_orig_dct.sub8x8_dct( dct, a, b ); }
/* (sub4x4_dct is left alone: on this path only the host-side intra 4x4 analysis calls it) */

static void install_count_wrappers( x264_t *h )
{
    int i;
    g_orig_mc = h->mc; g_orig_dct = h->dctf;
    h->mc.mc_luma = w_mc_luma; h->mc.get_ref = w_get_ref; h->mc.mc_chroma = w_mc_chroma;
    h->dctf.sub16x16_dct = w_sub16x16_dct; h->dctf.sub8x8_dct = w_sub8x8_dct;
    x264_pixel_cmp_t wsad[7]  = { w_sad_0, w_sad_1, w_sad_2, w_sad_3, w_sad_4, w_sad_5, w_sad_6 };
    x264_pixel_cmp_t wsatd[7] = { w_satd_0, w_satd_1, w_satd_2, w_satd_3, w_satd_4, w_satd_5, w_satd_6 };
    x264_pixel_cmp_x3_t wsad3[7]  = { w_sad_x3_0, w_sad_x3_1, w_sad_x3_2, w_sad_x3_3, w_sad_x3_4, w_sad_x3_5, w_sad_x3_6 };
    x264_pixel_cmp_x4_t wsad4[7]  = { w_sad_x4_0, w_sad_x4_1, w_sad_x4_2, w_sad_x4_3, w_sad_x4_4, w_sad_x4_5, w_sad_x4_6 };
    x264_pixel_cmp_x3_t wsatd3[7] = { w_satd_x3_0, w_satd_x3_1, w_satd_x3_2, w_satd_x3_3, w_satd_x3_4, w_satd_x3_5, w_satd_x3_6 };
    x264_pixel_cmp_x4_t wsatd4[7] = { w_satd_x4_0, w_satd_x4_1, w_satd_x4_2, w_satd_x4_3, w_satd_x4_4, w_satd_x4_5, w_satd_x4_6 };
    g_orig = h->pixf;
    for( i = 0; i < 7; i++ )
    {
        /* re-point only the aliases the ME path calls through; sad[]/satd[] themselves stay
         * untouched so intra analysis and everything else is unaffected. */
        h->pixf.fpelcmp[i]    = h->pixf.fpelcmp[i]    == g_orig.satd[i]    ? wsatd[i]  : wsad[i];
        h->pixf.fpelcmp_x3[i] = h->pixf.fpelcmp_x3[i] == g_orig.satd_x3[i] ? wsatd3[i] : wsad3[i];
        h->pixf.fpelcmp_x4[i] = h->pixf.fpelcmp_x4[i] == g_orig.satd_x4[i] ? wsatd4[i] : wsad4[i];
        h->pixf.mbcmp[i]           = h->pixf.mbcmp[i] == g_orig.satd[i] ? wsatd[i] : wsad[i];
        h->pixf.mbcmp_unaligned[i] = g_orig.mbcmp_unaligned[i] == g_orig.satd[i] ? wsatd[i]
                                   : g_orig.mbcmp_unaligned[i] == g_orig.sad[i] ? wsad[i]
                                   : g_orig.mbcmp_unaligned[i]; /* (only differs with asm) */
    }
    /* TESA calls h->pixf.sad / sad_x3 directly (encoder/me.c:536-566) */
    if( h->param.analyse.i_me_method == X264_ME_TESA )
        for( i = 0; i < 7; i++ )
        {
            h->pixf.sad[i] = wsad[i];
            h->pixf.sad_x3[i] = wsad3[i];
        }
}

/* ---- open / close ---------------------------------------------------------------------------- */
void pcamv_hook_open( x264_t *h )
{
    const char *s;
    g_t_total0 = now_s();
    if( (s = getenv( "PCAMV_DUMP" )) && *s )
        g_dump = fopen( s, "wb" );
    g_dump_planes = (s = getenv( "PCAMV_DUMP_PLANES" )) && atoi( s );
    if( (s = getenv( "PCAMV_DUMP_CALLS" )) ) g_dump_calls = atoi( s );
    if( (s = getenv( "PCAMV_DUMP_FRAMES" )) ) sscanf( s, "%d:%d", &g_frame_lo, &g_frame_hi );
    g_count = (s = getenv( "PCAMV_COUNT" )) && atoi( s );
    g_stats_path = getenv( "PCAMV_STATS" );
    if( g_count )
        install_count_wrappers( h );
    if( g_dump )
    {
        /* 'CFG0': 24 x int32 */
        int32_t c[24] = {
            h->param.i_width, h->param.i_height, h->sps->i_mb_width, h->sps->i_mb_height,
            h->param.analyse.i_me_method, h->param.analyse.i_me_range, h->param.analyse.i_subpel_refine,
            h->param.i_frame_reference, h->param.analyse.b_chroma_me, h->param.analyse.i_mv_range,
            h->param.rc.i_qp_constant, h->param.b_cabac, (int32_t)h->param.analyse.inter,
            h->param.analyse.b_mixed_references, h->param.analyse.b_fast_pskip, h->param.analyse.b_dct_decimate,
            h->param.i_keyint_max, h->param.analyse.b_transform_8x8, h->param.analyse.i_trellis,
            h->param.i_threads, h->mb.i_b4_stride, h->mb.i_b8_stride, h->mb.i_mb_stride,
            (int32_t)(h->param.eparam.iEmRate*1000000.0f + 0.5f) };
        rec_begin( "CFG0", sizeof(c) );
        fwrite( c, 1, sizeof(c), g_dump );
    }
}

void pcamv_hook_close( x264_t *h )
{
    (void)h;
    if( g_stats_path && *g_stats_path )
    {
        FILE *f = fopen( g_stats_path, "w" );
        if( f )
        {
            fprintf( f, "{\"sad\": %llu, \"satd\": %llu, \"ih_luma\": %llu, \"ih_chroma\": %llu, "
                        "\"pix_sad\": %llu, \"pix_satd\": %llu, \"pix_avg\": %llu, \"pix_chroma_mc\": %llu, \"pix_dct\": %llu, "
                        "\"searches\": %llu, \"refines\": %llu, \"p_mb_passes\": %llu, \"p_frames\": %llu, "
                        "\"ih_calls\": %llu, \"t_ih\": %.6f, "
                        "\"t_analyse_p\": %.6f, \"t_me\": %.6f, \"t_total\": %.6f}\n",
                     (unsigned long long)g_cnt_sad, (unsigned long long)g_cnt_satd,
                     (unsigned long long)g_cnt_ih_luma, (unsigned long long)g_cnt_ih_chroma,
                     (unsigned long long)g_pix_sad, (unsigned long long)g_pix_satd,
                     (unsigned long long)g_pix_avg, (unsigned long long)g_pix_chroma_mc, (unsigned long long)g_pix_dct,
                     (unsigned long long)g_cnt_search, (unsigned long long)g_cnt_refine,
                     (unsigned long long)g_cnt_mb_p, (unsigned long long)g_cnt_pframes,
                     (unsigned long long)g_cnt_ih_calls, g_t_ih,
                     g_t_analyse_p, g_t_me, now_s() - g_t_total0 );
            fclose( f );
        }
    }
    if( g_dump ) { fclose( g_dump ); g_dump = NULL; }
}

/* ---- slice begin: header + (optionally) planes ------------------------------------------------ */
static void dump_plane( const uint8_t *p, int stride, int rows )
{
    fwrite( p, 1, (size_t)stride*rows, g_dump );
}

void pcamv_hook_slice_begin( x264_t *h )
{
    g_pslice = h->sh.i_type == SLICE_TYPE_P;
    g_pass = !h->info.embed_flag ? 0 : h->info.firstTime ? 1 : 2;
    if( g_pslice && g_pass != 2 )
        g_cnt_pframes++;
    {
        uint64_t c[9] = { g_cnt_sad, g_cnt_satd, g_cnt_ih_luma, g_cnt_ih_chroma, g_pix_sad, g_pix_satd,
                          g_cnt_search, g_cnt_refine, g_cnt_ih_calls };
        double t[3] = { g_t_me, g_t_ih, g_t_analyse_p };
        memcpy( g_slice_c0, c, sizeof(c) );
        memcpy( g_slice_t0, t, sizeof(t) );
        g_slice_x0[0] = g_pix_avg; g_slice_x0[1] = g_pix_chroma_mc; g_slice_x0[2] = g_pix_dct;
        g_slice_x0[3] = g_ih_pix_avg; g_slice_x0[4] = g_ih_pix_chroma_mc; g_slice_x0[5] = g_ih_pix_dct; g_slice_x0[6] = g_ih_pix_satd;
    }
    if( !dump_on( h ) )
        return;
    {
        /* 'SLCB': 16 x int32 header, then (if planes) fenc Y,U,V (unpadded rows, frame stride) and per
         * reference: 4 padded luma planes (whole buffer incl. 32-px borders), padded U, padded V. */
        x264_frame_t *fe = h->fenc;
        int nref = g_pslice ? h->i_ref0 : 0;
        int with_planes = g_dump_planes && g_pslice;
        int32_t hd[16] = { h->i_frame, g_pass, h->sh.i_type, h->sh.i_qp, nref, with_planes,
                           fe->i_stride[0], fe->i_stride[1], fe->i_lines[0], fe->i_lines[1],
                           fe->i_width[0], fe->i_frame, h->fdec->i_poc, h->sh.i_first_mb, h->sh.i_last_mb, 0 };
        size_t luma_sz = (size_t)fe->i_stride[0]*(fe->i_lines[0] + 2*PADV);
        size_t chroma_sz = (size_t)fe->i_stride[1]*(fe->i_lines[1] + PADV);  /* 16 rows above + 16 below */
        size_t n = sizeof(hd);
        int r, k;
        if( with_planes )
            n += (size_t)fe->i_stride[0]*fe->i_lines[0] + 2*(size_t)fe->i_stride[1]*fe->i_lines[1]
               + (size_t)nref*(4*luma_sz + 2*chroma_sz + 8);
        rec_begin( "SLCB", (uint32_t)n );
        fwrite( hd, 1, sizeof(hd), g_dump );
        if( with_planes )
        {
            dump_plane( fe->plane[0], fe->i_stride[0], fe->i_lines[0] );
            dump_plane( fe->plane[1], fe->i_stride[1], fe->i_lines[1] );
            dump_plane( fe->plane[2], fe->i_stride[2], fe->i_lines[2] );
            for( r = 0; r < nref; r++ )
            {
                x264_frame_t *fr = h->fref0[r];
                int32_t rh[2] = { fr->i_poc, fr->i_frame };
                fwrite( rh, 1, sizeof(rh), g_dump );
                for( k = 0; k < 4; k++ )
                    dump_plane( fr->filtered[k] - fr->i_stride[0]*PADV - PADH, fr->i_stride[0], fr->i_lines[0] + 2*PADV );
                for( k = 1; k < 3; k++ )
                    dump_plane( fr->plane[k] - fr->i_stride[k]*(PADV/2) - PADH/2, fr->i_stride[k], fr->i_lines[k] + PADV );
            }
        }
    }
    if( g_pslice )
    {
        /* 'SLCX': what the P-slice analysis reads besides pixels: int32 cur_poc, n_ref, ref_poc[16], col_n_ref,
         * col_inv_ref_poc[16], n_mb; then the co-located frame's (fref0[0]) int8 ref[4*n_mb] and int16 mv[16*n_mb][2]
         * (temporal MV candidates, common/macroblock.c:444-467); then 'QNT0' once per QP with the quantiser tables. */
        x264_frame_t *l0 = h->fref0[0];
        int n_mb = h->mb.i_mb_count, i;
        int32_t hd[36];
        memset( hd, 0, sizeof(hd) );
        hd[0] = h->fdec->i_poc; hd[1] = h->i_ref0;
        for( i = 0; i < h->i_ref0 && i < 16; i++ ) hd[2+i] = h->fref0[i]->i_poc;
        hd[18] = l0->i_ref[0];
        for( i = 0; i < 16; i++ ) hd[19+i] = l0->inv_ref_poc[i];
        hd[35] = n_mb;
        rec_begin( "SLCX", (uint32_t)(sizeof(hd) + 4*n_mb + 64*n_mb) );
        fwrite( hd, 1, sizeof(hd), g_dump );
        fwrite( l0->ref[0], 1, 4*n_mb, g_dump );
        fwrite( l0->mv[0], 1, 64*n_mb, g_dump );
        {
            static int done[52];
            int qp = h->sh.i_qp, qpc = h->chroma_qp_table[qp];
            if( !done[qp] )
            {
                /* 'QNT0': int32 qp, chroma_qp, lambda2[chroma_qp]; uint16 mf_y[16], bias_y[16], mf_c[16], bias_c[16];
                 * int32 dequant_y[6][16], dequant_c[6][16]  (CQM_4PY / CQM_4PC) */
                int32_t q[3] = { qp, qpc, x264_lambda2_tab[qpc] };
                done[qp] = 1;
                rec_begin( "QNT0", sizeof(q) + 4*32 + 2*96*4 );
                fwrite( q, 1, sizeof(q), g_dump );
                fwrite( h->quant4_mf[CQM_4PY][qp], 2, 16, g_dump );
                fwrite( h->quant4_bias[CQM_4PY][qp], 2, 16, g_dump );
                fwrite( h->quant4_mf[CQM_4PC][qpc], 2, 16, g_dump );
                fwrite( h->quant4_bias[CQM_4PC][qpc], 2, 16, g_dump );
                fwrite( h->dequant4_mf[CQM_4PY], 4, 96, g_dump );
                fwrite( h->dequant4_mf[CQM_4PC], 4, 96, g_dump );
            }
        }
    }
}

/* ---- slice end: final per-MB decisions as stored in the frame arrays -------------------------- */
extern int16_t *g_cost_mv[52];            /* reference encoder/analyse.c:189 (malloc base, centre at +2*4*2048) */
extern uint16_t x264_cost_ref[52][3][33]; /* reference encoder/analyse.c:188 */
static int g_cmv_dumped[52];

void pcamv_hook_slice_end( x264_t *h )
{
    if( !dump_on( h ) || !g_pslice )
        return;
    if( g_cost_mv[h->sh.i_qp] && !g_cmv_dumped[h->sh.i_qp] )
    {
        /* 'CMV0': int32 qp, lambda; int16 cost_mv[32769] (index 0 <-> mv delta -16384); uint16 cost_ref[3][33] */
        int32_t hd[2] = { h->sh.i_qp, x264_lambda_tab[h->sh.i_qp] };
        g_cmv_dumped[h->sh.i_qp] = 1;
        rec_begin( "CMV0", sizeof(hd) + 32769*2 + 3*33*2 );
        fwrite( hd, 1, sizeof(hd), g_dump );
        fwrite( g_cost_mv[h->sh.i_qp], 2, 32769, g_dump );
        fwrite( x264_cost_ref[h->sh.i_qp], 2, 3*33, g_dump );
    }
    {
        /* 'CNT0': work of THIS slice pass (uint64 each): sad, satd, ih_luma, ih_chroma, pix_sad, pix_satd, searches,
         * refines, ih_calls; then double t_me, t_ih, t_analyse_p (seconds spent in this pass). */
        uint64_t c[9] = { g_cnt_sad, g_cnt_satd, g_cnt_ih_luma, g_cnt_ih_chroma, g_pix_sad, g_pix_satd,
                          g_cnt_search, g_cnt_refine, g_cnt_ih_calls };
        double t[3] = { g_t_me, g_t_ih, g_t_analyse_p };
        int k;
        for( k = 0; k < 9; k++ ) c[k] -= g_slice_c0[k];
        for( k = 0; k < 3; k++ ) t[k] -= g_slice_t0[k];
        rec_begin( "CNT0", sizeof(c) + sizeof(t) );
        fwrite( c, 1, sizeof(c), g_dump );
        fwrite( t, 1, sizeof(t), g_dump );
        {
            /* 'CNT1': pixels of this slice pass that went through quarter-pel averaging, chroma MC, DCT/quant/IDCT (uint64 each), then
             * the share of x264_ih_get_mv_cost in those three and in pix_satd */
            uint64_t x[7] = { g_pix_avg - g_slice_x0[0], g_pix_chroma_mc - g_slice_x0[1], g_pix_dct - g_slice_x0[2],
                              g_ih_pix_avg - g_slice_x0[3], g_ih_pix_chroma_mc - g_slice_x0[4], g_ih_pix_dct - g_slice_x0[5], g_ih_pix_satd - g_slice_x0[6] };
            rec_begin( "CNT1", sizeof(x) );
            fwrite( x, 1, sizeof(x), g_dump );
        }
    }
    {
        /* 'SLCE': int32 frame, pass, n_mb; then int8 type[n_mb]; int8 ref[4*n_mb] in b8 raster;
         * int16 mv[16*n_mb][2] in b4 raster (frame layout, stride i_b4_stride); int16 mvr0[n_mb][2]. */
        int n_mb = h->mb.i_mb_count;
        int32_t hd[3] = { h->i_frame, g_pass, n_mb };
        rec_begin( "SLCE", (uint32_t)(sizeof(hd) + n_mb + 4*n_mb + 64*n_mb + 4*n_mb) );
        fwrite( hd, 1, sizeof(hd), g_dump );
        fwrite( h->mb.type, 1, n_mb, g_dump );
        fwrite( h->mb.ref[0], 1, 4*n_mb, g_dump );
        fwrite( h->mb.mv[0], 1, 64*n_mb, g_dump );
        fwrite( h->mb.mvr[0][0], 1, 4*n_mb, g_dump );
    }
}

/* ---- per-MB analyse timer ---------------------------------------------------------------------- */
void pcamv_hook_analyse_begin( x264_t *h )
{
    if( h->sh.i_type == SLICE_TYPE_P )
    {
        clock_gettime( CLOCK_MONOTONIC, &g_ts_an );
        g_in_analyse_p = 1;
    }
}

void pcamv_hook_analyse_end( x264_t *h )
{
    g_in_analyse_p = 0;
    if( h->sh.i_type == SLICE_TYPE_P )
    {
        struct timespec t; clock_gettime( CLOCK_MONOTONIC, &t );
        g_t_analyse_p += (t.tv_sec - g_ts_an.tv_sec) + 1e-9*(t.tv_nsec - g_ts_an.tv_nsec);
        g_cnt_mb_p++;
        if( dump_on( h ) )
        {
            /* 'MBAN': the decision x264_macroblock_analyse left in h->mb for this MB (either pass).
             * int32 frame, pass, mb_xy, type, partition, sub[4], b_skip_mc, qp; int16 mv[16][2] in
             * block_idx order (x264_scan8), int8 ref[4] per 8x8, int16 pskip_mv[2], int16 mvp16[2]. */
            int32_t hd[11] = { h->i_frame, g_pass, h->mb.i_mb_xy, h->mb.i_type, h->mb.i_partition,
                               h->mb.i_sub_partition[0], h->mb.i_sub_partition[1], h->mb.i_sub_partition[2],
                               h->mb.i_sub_partition[3], h->mb.b_skip_mc, h->mb.i_qp };
            int16_t mv[16][2]; int8_t ref[4]; int i;
            for( i = 0; i < 16; i++ )
            {
                mv[i][0] = h->mb.cache.mv[0][x264_scan8[i]][0];
                mv[i][1] = h->mb.cache.mv[0][x264_scan8[i]][1];
            }
            for( i = 0; i < 4; i++ )
                ref[i] = h->mb.cache.ref[0][x264_scan8[4*i]];
            rec_begin( "MBAN", sizeof(hd) + sizeof(mv) + sizeof(ref) + 4 );
            fwrite( hd, 1, sizeof(hd), g_dump );
            fwrite( mv, 1, sizeof(mv), g_dump );
            fwrite( ref, 1, sizeof(ref), g_dump );
            fwrite( h->mb.cache.pskip_mv, 1, 4, g_dump );
        }
    }
}

/* ---- RD mode decision (x264_dump_rd only: tools/reftree.py::rd_hook puts the call into x264_rd_cost_mb, encoder/rdo.c:139-172) ----
 * 'RDMB': one record per x264_rd_cost_mb call that sized the macroblock with CAVLC: what x264_macroblock_size_cavlc saw and what it
 * returned - the checker of csrc/pcamv_cavlc.cuh (tests/emu/emu_cavlc_check.cpp).  'VLC0' (once): the bit LENGTHS of the encoder's own
 * CAVLC tables (common/vlc.c), the way 'CMV0' carries its lambda*bits tables: a product host would upload them the same way. */
static int g_vlc_dumped;
void pcamv_hook_rd_mb( x264_t *h, int i_ssd, int i_bits_encoded, int i_lambda2 )
{
    int i, j, n_mvd = 0;
    int16_t mvd[16][2];
    int16_t mvp[2];
    if( !dump_on( h ) || h->sh.i_type != SLICE_TYPE_P )
        return;
    if( !g_vlc_dumped )
    {
        uint8_t z[5 + 5*64 + 15*16 + 3*4 + 7*16], *q = z;
        for( i = 0; i < 5; i++ ) *q++ = x264_coeff0_token[i].i_size;
        for( i = 0; i < 5; i++ ) for( j = 0; j < 64; j++ ) *q++ = x264_coeff_token[i][j].i_size;
        for( i = 0; i < 15; i++ ) for( j = 0; j < 16; j++ ) *q++ = x264_total_zeros[i][j].i_size;
        for( i = 0; i < 3; i++ ) for( j = 0; j < 4; j++ ) *q++ = x264_total_zeros_dc[i][j].i_size;
        for( i = 0; i < 7; i++ ) for( j = 0; j < 16; j++ ) *q++ = x264_run_before[i][j].i_size;
        rec_begin( "VLC0", sizeof(z) );
        fwrite( z, 1, sizeof(z), g_dump );
        g_vlc_dumped = 1;
    }
    memset( mvd, 0, sizeof(mvd) );
#define RD_MVD( idx, width ) do { x264_mb_predict_mv( h, 0, idx, width, mvp ); \
        mvd[n_mvd][0] = h->mb.cache.mv[0][x264_scan8[idx]][0] - mvp[0]; mvd[n_mvd][1] = h->mb.cache.mv[0][x264_scan8[idx]][1] - mvp[1]; n_mvd++; } while( 0 )
    if( h->mb.i_type == P_L0 )
    {
        if( h->mb.i_partition == D_16x16 ) RD_MVD( 0, 4 );
        else if( h->mb.i_partition == D_16x8 ) { RD_MVD( 0, 4 ); RD_MVD( 8, 4 ); }
        else if( h->mb.i_partition == D_8x16 ) { RD_MVD( 0, 2 ); RD_MVD( 4, 2 ); }
    }
    else if( h->mb.i_type == P_8x8 )
        for( i = 0; i < 4; i++ )
            switch( h->mb.i_sub_partition[i] )
            {
                case D_L0_8x8: RD_MVD( 4*i, 2 ); break;
                case D_L0_8x4: RD_MVD( 4*i, 2 ); RD_MVD( 4*i + 2, 2 ); break;
                case D_L0_4x8: RD_MVD( 4*i, 1 ); RD_MVD( 4*i + 1, 1 ); break;
                case D_L0_4x4: RD_MVD( 4*i, 1 ); RD_MVD( 4*i + 1, 1 ); RD_MVD( 4*i + 2, 1 ); RD_MVD( 4*i + 3, 1 ); break;
            }
#undef RD_MVD
    {
        int32_t hd[20] = { h->i_frame, g_pass, h->mb.i_mb_xy, h->mb.i_type, h->mb.i_partition,
                           h->mb.i_sub_partition[0], h->mb.i_sub_partition[1], h->mb.i_sub_partition[2], h->mb.i_sub_partition[3],
                           h->mb.pic.i_fref[0], !!( h->param.analyse.inter & X264_ANALYSE_PSUB8x8 ), h->mb.i_cbp_luma, h->mb.i_cbp_chroma,
                           h->mb.i_qp - h->mb.i_last_qp, i_ssd, i_bits_encoded, i_lambda2, n_mvd, h->mb.i_psy_rd, x264_lambda_tab[h->mb.i_qp] };
        int8_t ref[4];
        int16_t mv[16][2];
        for( i = 0; i < 4; i++ ) ref[i] = h->mb.cache.ref[0][x264_scan8[4*i]];
        for( i = 0; i < 16; i++ ) { mv[i][0] = h->mb.cache.mv[0][x264_scan8[i]][0]; mv[i][1] = h->mb.cache.mv[0][x264_scan8[i]][1]; }
        rec_begin( "RDMB", sizeof(hd) + sizeof(ref) + sizeof(mvd) + 48 + 24*16*2 + 2*4*2 + sizeof(mv) + 4 );
        fwrite( hd, 1, sizeof(hd), g_dump );
        fwrite( ref, 1, sizeof(ref), g_dump );
        fwrite( mvd, 1, sizeof(mvd), g_dump );
        fwrite( h->mb.cache.non_zero_count, 1, 48, g_dump );
        fwrite( h->dct.luma4x4, 1, 24*16*2, g_dump );
        fwrite( h->dct.chroma_dc, 1, 2*4*2, g_dump );
        fwrite( mv, 1, sizeof(mv), g_dump );              /* the candidate's vectors, block_idx order (for the distortion half) */
        {
            /* quirk q1: with b_skip_mc left set x264_macroblock_encode does not motion-compensate (encoder/macroblock.c:611-612) and
             * the "candidate" is coded against whatever the last analysis left in fdec */
            int32_t skip_mc = h->mb.b_skip_mc;
            fwrite( &skip_mc, 4, 1, g_dump );
        }
    }
}

/* ---- intra analysis (x264_dump_rd only: tools/reftree.py::intra_hook wraps x264_mb_analyse_intra, encoder/analyse.c:628-879) ----------
 * 'INTR': one record per call: what the 16x16 / chroma mode analysis saw (source macroblock, the reconstructed border pixels of the
 * neighbours, which neighbours exist, lambda) and what it decided - the checker of csrc/pcamv_intra.cuh (tests/emu/emu_intra_check.cpp). */
void pcamv_hook_intra( x264_t *h, int lambda, int i_satd_inter, int satd16, int pred16, const int *dir16, int satd_c, int pred_c, int satd4, const int *pred4,
                       int i_qp, int i_mbrd, int b_fast_intra, int i_satd_i8x8 )
{
    int i, pl;
    if( !dump_on( h ) )
        return;
    {
        int32_t hd[20] = { h->i_frame, g_pass, h->mb.i_mb_xy, h->sh.i_type, !!( h->mb.i_neighbour & MB_LEFT ), !!( h->mb.i_neighbour & MB_TOP ),
                           !!( h->mb.i_neighbour & MB_TOPLEFT ), lambda, i_satd_inter, satd16, pred16, dir16[0], dir16[1], dir16[2], dir16[3],
                           satd_c, pred_c, satd4, h->mb.b_chroma_me, 0 };
        uint8_t pix[256 + 2*64], border[33 + 2*17];
        int32_t p4[16];
        for( i = 0; i < 16; i++ ) memcpy( pix + 16*i, h->mb.pic.p_fenc[0] + i*FENC_STRIDE, 16 );
        for( pl = 0; pl < 2; pl++ )
            for( i = 0; i < 8; i++ ) memcpy( pix + 256 + 64*pl + 8*i, h->mb.pic.p_fenc[1 + pl] + i*FENC_STRIDE, 8 );
        /* luma: topleft, top[16], left[16]; chroma: topleft, top[8], left[8] per plane */
        border[0] = h->mb.pic.p_fdec[0][-FDEC_STRIDE - 1];
        for( i = 0; i < 16; i++ ) { border[1 + i] = h->mb.pic.p_fdec[0][-FDEC_STRIDE + i]; border[17 + i] = h->mb.pic.p_fdec[0][i*FDEC_STRIDE - 1]; }
        for( pl = 0; pl < 2; pl++ )
        {
            uint8_t *b = border + 33 + 17*pl, *f = h->mb.pic.p_fdec[1 + pl];
            b[0] = f[-FDEC_STRIDE - 1];
            for( i = 0; i < 8; i++ ) { b[1 + i] = f[-FDEC_STRIDE + i]; b[9 + i] = f[i*FDEC_STRIDE - 1]; }
        }
        for( i = 0; i < 16; i++ ) p4[i] = pred4[i];
        {
            /* for the 4x4 modes: qp and the analysis switches, which neighbours every 4x4 block has, the cached prediction modes of the
             * blocks left of / above the macroblock, the four pixels right of the top row, the intra quantiser of this qp */
            int32_t x4[4] = { i_qp, i_mbrd, b_fast_intra, i_satd_i8x8 };
            uint8_t nb4[16], tr[4];
            int8_t lm[4], tm[4];
            static const int left_idx[4] = { 0, 2, 8, 10 }, top_idx[4] = { 0, 1, 4, 5 };
            for( i = 0; i < 16; i++ ) nb4[i] = (uint8_t)h->mb.i_neighbour4[i];
            for( i = 0; i < 4; i++ )
            {
                lm[i] = h->mb.cache.intra4x4_pred_mode[x264_scan8[left_idx[i]] - 1];
                tm[i] = h->mb.cache.intra4x4_pred_mode[x264_scan8[top_idx[i]] - 8];
                tr[i] = h->mb.pic.p_fdec[0][-FDEC_STRIDE + 16 + i];
            }
            rec_begin( "INTR", sizeof(hd) + sizeof(pix) + sizeof(border) + sizeof(p4) + sizeof(x4) + 16 + 4 + 4 + 4 + 32 + 32 + 6*16*4 );
            fwrite( hd, 1, sizeof(hd), g_dump );
            fwrite( pix, 1, sizeof(pix), g_dump );
            fwrite( border, 1, sizeof(border), g_dump );
            fwrite( p4, 1, sizeof(p4), g_dump );
            fwrite( x4, 1, sizeof(x4), g_dump );
            fwrite( nb4, 1, 16, g_dump ); fwrite( lm, 1, 4, g_dump ); fwrite( tm, 1, 4, g_dump ); fwrite( tr, 1, 4, g_dump );
            fwrite( h->quant4_mf[CQM_4IY][i_qp], 2, 16, g_dump );
            fwrite( h->quant4_bias[CQM_4IY][i_qp], 2, 16, g_dump );
            fwrite( h->dequant4_mf[CQM_4IY], 4, 6*16, g_dump );
        }
    }
}

/* ---- embed stage -------------------------------------------------------------------------------- */
void pcamv_hook_embed( x264_t *h, int an )
{
    if( !dump_on( h ) )
        return;
    {
        /* 'EMBD': int32 frame, n_mb, length, an, num_filp; then per MB a packed record
         * { int32 type, qp, partition; uint8 used, sub[4], pad[3]; int16 mv_stego[16][2];
         *   int32 inter_stego_cost[16]; int8 ref[16]; int16 mv[16][2]; int16 pskip_mv[2] } = 232 B;
         * then uint8 cover[length], float rho_final[length], uint8 message[an], uint8 stego[length],
         * int8 filp[length]. */
        int n_mb = h->sh.i_last_mb, i, len = h->info.length;
        int32_t hd[5] = { h->i_frame, n_mb, len, an, (int32_t)h->info.num_filp };
        uint32_t n = sizeof(hd) + 232u*n_mb + len + 4u*len + (an > 0 ? an : 0) + len + len;
        rec_begin( "EMBD", n );
        fwrite( hd, 1, sizeof(hd), g_dump );
        for( i = 0; i < n_mb; i++ )
        {
            int32_t a[3] = { h->info.cache[i].i_type, h->info.cache[i].i_qp, h->info.cache[i].i_partition };
            uint8_t b[8] = { h->info.cache[i].used, h->info.cache[i].i_sub_partition[0], h->info.cache[i].i_sub_partition[1],
                             h->info.cache[i].i_sub_partition[2], h->info.cache[i].i_sub_partition[3], 0, 0, 0 };
            fwrite( a, 1, sizeof(a), g_dump );
            fwrite( b, 1, sizeof(b), g_dump );
            fwrite( h->info.cache[i].mv_stego, 1, 64, g_dump );
            fwrite( h->info.cache[i].inter_stego_cost, 1, 64, g_dump );
            fwrite( h->info.cache[i].ref, 1, 16, g_dump );
            fwrite( h->info.cache[i].mv, 1, 64, g_dump );
            fwrite( h->info.cache[i].pskip_mv_, 1, 4, g_dump );
        }
        fwrite( h->info.cover, 1, len, g_dump );
        fwrite( h->info.rho_final, 4, len, g_dump );
        if( an > 0 ) fwrite( h->info.message, 1, an, g_dump );
        fwrite( h->info.stego, 1, len, g_dump );
        fwrite( h->info.filp, 1, len, g_dump );
    }
}

/* ---- search wrappers ------------------------------------------------------------------------------ */
typedef struct
{
    int32_t frame, pass, mb_xy, mb_x, mb_y;
    int32_t i_pixel, i_ref, xoff, yoff, i_ref_cost;
    int32_t me_method, me_range, subme, b_chroma_me, qp;
    int32_t mv_min_fpel[2], mv_max_fpel[2], mv_min_spel[2], mv_max_spel[2];
    int32_t i_mvc, has_thresh, thresh_in, thresh_out;
    int16_t mvp[2];
    int16_t mvc[10][2];
    /* in (refine only) */
    int16_t mv_in[2]; int32_t cost_in, cost_mv_in;
    /* out */
    int16_t mv[2]; int32_t cost, cost_mv;
    /* accounting: block-distortion evaluations made by this call (PCAMV_COUNT=1) and its wall time */
    int32_t n_cand, t_ns, pix_sad, pix_satd;
} pcamv_call_rec_t;

static void fill_common( x264_t *h, x264_me_t *m, pcamv_call_rec_t *r )
{
    int off = (int)(m->p_fenc[0] - h->mb.pic.p_fenc[0]);
    memset( r, 0, sizeof(*r) );
    r->frame = h->i_frame; r->pass = g_pass; r->mb_xy = h->mb.i_mb_xy; r->mb_x = h->mb.i_mb_x; r->mb_y = h->mb.i_mb_y;
    r->i_pixel = m->i_pixel; r->i_ref = m->i_ref; r->xoff = off % FENC_STRIDE; r->yoff = off / FENC_STRIDE;
    {
        /* sub-8x8 searches leave m->i_ref unset (reference encoder/analyse.c:1569-1693): recover the
         * reference index from the plane pointer instead */
        int k;
        for( k = 0; k < h->mb.pic.i_fref[0]; k++ )
            if( m->p_fref[0] == &h->mb.pic.p_fref[0][k][0][r->xoff + r->yoff*m->i_stride[0]] )
                r->i_ref = k;
    }
    r->i_ref_cost = m->i_ref_cost;
    r->me_method = h->mb.i_me_method; r->me_range = h->param.analyse.i_me_range; r->subme = h->mb.i_subpel_refine;
    r->b_chroma_me = h->mb.b_chroma_me; r->qp = h->mb.i_qp;
    r->mv_min_fpel[0] = h->mb.mv_min_fpel[0]; r->mv_min_fpel[1] = h->mb.mv_min_fpel[1];
    r->mv_max_fpel[0] = h->mb.mv_max_fpel[0]; r->mv_max_fpel[1] = h->mb.mv_max_fpel[1];
    r->mv_min_spel[0] = h->mb.mv_min_spel[0]; r->mv_min_spel[1] = h->mb.mv_min_spel[1];
    r->mv_max_spel[0] = h->mb.mv_max_spel[0]; r->mv_max_spel[1] = h->mb.mv_max_spel[1];
    r->mvp[0] = m->mvp[0]; r->mvp[1] = m->mvp[1];
}

void x264_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh )
{
    pcamv_call_rec_t r;
    int rec = g_dump_calls && dump_on( h ) && h->sh.i_type == SLICE_TYPE_P;
    struct timespec t0, t1;
    if( rec )
    {
        int i;
        fill_common( h, m, &r );
        r.i_mvc = i_mvc;
        for( i = 0; i < i_mvc && i < 10; i++ ) { r.mvc[i][0] = mvc[i][0]; r.mvc[i][1] = mvc[i][1]; }
        r.has_thresh = p_halfpel_thresh != NULL;
        r.thresh_in = p_halfpel_thresh ? *p_halfpel_thresh : 0;
    }
    uint64_t c0 = g_cnt_sad + g_cnt_satd, ps0 = g_pix_sad, pt0 = g_pix_satd;
    clock_gettime( CLOCK_MONOTONIC, &t0 );
    g_in_me++;
    x264_me_search_ref_real( h, m, mvc, i_mvc, p_halfpel_thresh );
    g_in_me--;
    clock_gettime( CLOCK_MONOTONIC, &t1 );
    g_t_me += (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
    g_cnt_search++;
    if( rec )
    {
        r.n_cand = (int32_t)(g_cnt_sad + g_cnt_satd - c0);
        r.pix_sad = (int32_t)(g_pix_sad - ps0); r.pix_satd = (int32_t)(g_pix_satd - pt0);
        r.t_ns = (int32_t)((t1.tv_sec - t0.tv_sec)*1000000000LL + (t1.tv_nsec - t0.tv_nsec));
        r.thresh_out = p_halfpel_thresh ? *p_halfpel_thresh : 0;
        r.mv[0] = m->mv[0]; r.mv[1] = m->mv[1]; r.cost = m->cost; r.cost_mv = m->cost_mv;
        rec_begin( "MESR", sizeof(r) );
        fwrite( &r, 1, sizeof(r), g_dump );
    }
}

void x264_me_refine_qpel( x264_t *h, x264_me_t *m )
{
    pcamv_call_rec_t r;
    int rec = g_dump_calls && dump_on( h ) && h->sh.i_type == SLICE_TYPE_P;
    struct timespec t0, t1;
    if( rec )
    {
        fill_common( h, m, &r );
        r.mv_in[0] = m->mv[0]; r.mv_in[1] = m->mv[1]; r.cost_in = m->cost; r.cost_mv_in = m->cost_mv;
    }
    uint64_t c0 = g_cnt_sad + g_cnt_satd, ps0 = g_pix_sad, pt0 = g_pix_satd;
    clock_gettime( CLOCK_MONOTONIC, &t0 );
    g_in_me++;
    x264_me_refine_qpel_real( h, m );
    g_in_me--;
    clock_gettime( CLOCK_MONOTONIC, &t1 );
    g_t_me += (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
    g_cnt_refine++;
    if( rec )
    {
        r.n_cand = (int32_t)(g_cnt_sad + g_cnt_satd - c0);
        r.pix_sad = (int32_t)(g_pix_sad - ps0); r.pix_satd = (int32_t)(g_pix_satd - pt0);
        r.t_ns = (int32_t)((t1.tv_sec - t0.tv_sec)*1000000000LL + (t1.tv_nsec - t0.tv_nsec));
        r.mv[0] = m->mv[0]; r.mv[1] = m->mv[1]; r.cost = m->cost; r.cost_mv = m->cost_mv;
        rec_begin( "MERQ", sizeof(r) );
        fwrite( &r, 1, sizeof(r), g_dump );
    }
}

void pcamv_hook_ih_begin( void ) { clock_gettime( CLOCK_MONOTONIC, &g_ts_ih ); g_cnt_ih_calls++; g_in_ih = 1; }
void pcamv_hook_ih_end( void )
{
    struct timespec t;
    g_in_ih = 0; clock_gettime( CLOCK_MONOTONIC, &t );
    g_t_ih += (t.tv_sec - g_ts_ih.tv_sec) + 1e-9*(t.tv_nsec - g_ts_ih.tv_nsec);
}

/* one MV_SATD_FDEC_IH evaluation (encoder/analyse.c:2364-2385): +1 luma SATD, +2 chroma when chroma ME */
/* right after x264_macroblock_encode (encoder/encoder.c:1881): nothing to record here */
void pcamv_hook_encoded( x264_t *h ) { (void)h; }

void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me )
{
    g_cnt_ih_luma++;
    g_pix_satd += g_pw[i_pixel]*g_ph[i_pixel];
    g_ih_pix_satd += g_pw[i_pixel]*g_ph[i_pixel];
    if( b_chroma_me )
    {
        g_cnt_ih_chroma += 2;
        g_pix_satd += g_pw[i_pixel]*g_ph[i_pixel]/2;
        g_ih_pix_satd += g_pw[i_pixel]*g_ph[i_pixel]/2;
    }
}
