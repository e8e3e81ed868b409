/* pcamv_bitstream.c - the decoder side of the PCAMV payload channel: motion vectors and payload from the .264 alone.
 *
 * SURVEY.md 8(f) row 4.  The reference is an encoder only (x264 build 66 has no decoder, its extractor include is commented
 * out, encoder/analyse.c:43).  This file is the mirror of what the reference WRITES for a P slice, read back:
 *
 *   NAL / SPS / PPS / slice header   encoder/set.c:215-355 (x264_sps_write), :429-470 (x264_pps_write),
 *                                    encoder/encoder.c:174-310 (x264_slice_header_write), common/common.c:658-695 (x264_nal_encode)
 *   CABAC macroblock layer           encoder/cabac.c:64-130 (mb_type), :233-330 (cbp, qp_delta, skip, sub partition),
 *                                    :375-505 (ref, mvd), :508-680 (coded_block_flag contexts, residual), :781-1030 (macroblock)
 *   CAVLC macroblock layer           encoder/cavlc.c:61-198 (residual), :200-262 (qp_delta, mvd), :285-600 (macroblock)
 *   arithmetic decoder               the inverse of common/cabac.c:861-925 (encode_decision / bypass / terminal / flush)
 *   motion vector prediction         common/macroblock.c:28-163 (x264_mb_predict_mv, _pskip) restated on whole-picture
 *                                    4x4 arrays: a neighbour is unavailable when it lies outside the picture or has not
 *                                    been decoded yet, which is what the encoder's cache encodes with ref = -2
 *
 * None of the reference's data tables is restated here: the context initialisation (460 x 4 models), the LPS ranges, the state
 * transitions and the CAVLC code tables are the encoder's own objects (x264_cabac_context_init, x264_cabac_range_lps,
 * x264_cabac_transition, x264_coeff_token, x264_total_zeros, x264_run_before - this file is linked with the reference's objects
 * like the rest of the bound host), read in the opposite direction.  What IS written out below are the syntax-level constants of
 * H.264 itself, which encoder and decoder necessarily share: the ctxIdxOffsets of the residual elements (Table 9-34: 85 / 105 /
 * 166 / 227 + the per-category offsets), the bin-to-context maps of mvd and coeff_abs_level_minus1 (9.3.3.1.1.7, 9.3.3.1.3), the
 * coded_block_pattern mapping of Table 9-4 and the CAVLC level arithmetic of 9.2.2.1.
 *
 * Scope = what the reference can emit on the PCAMV path: one slice per picture, frame macroblocks, P slices
 * whose macroblocks are P_L0 (16x16 / 16x8 / 8x16), P_8x8 (8x8 / 8x4 / 4x8 / 4x4) or P_SKIP, 4x4 transform.  I slices carry
 * no vectors and are stepped over.  Anything else (intra macroblocks in P slices, B slices, 8x8 transform, several slices,
 * interlace, weighted prediction) is refused with a message - there is no silent guess.
 *
 * On top of the parser: the cover / stego vector of a P picture in the embedder's order (encoder/encoder.c:1561-1655: one
 * element per vector-carrying partition, LSB of mv_x + mv_y) and the payload (syndrome under the embedder's sub-matrices,
 * host/pcamv_stc_extract.c).  CLI (host/build_host.py):
 *
 *   x264_pcamv --parse-mv IN.264 -o MV.bin [--csv]              per P picture: int32 picture, n_mb, mb_w, mb_h, qp, cabac + n_mb records
 *                                                               (--csv: one text line per macroblock instead)
 *   x264_pcamv --extract-264 IN.264 --emrate R -o MESSAGE.bin [--stego STEGO.bin]
 *
 * What bitstream-side extraction can and cannot give for the reference's own output is a property of the reference, measured
 * in tests/test_bitstream.py: the vector pass 2 writes for a partition is its flipped OWN vector when the trellis flips it and
 * the vector of the partition the unsequenced copy picked when it does not, so the stego bit read back is
 * wrong exactly for flipped carriers whose own vector and copied vector differ in parity.  With the copy made straight
 * (PCAMV_STRAIGHT_MV_COPY=1 in the bound host; off by default because it changes the bitstream) the round trip is exact. */
#include "common/common.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int pcamv_stc_extract( const uint8_t *stego, int n, uint8_t *message, int an, int matrixheight );
extern const uint8_t x264_cabac_range_lps[128][4];     /* common/cabac.c:669 (defined there, declared nowhere) */

/* ---- output record ------------------------------------------------------------------------------------------------------ */
typedef struct
{
    int32_t type;           /* P_L0 = 4, P_8x8 = 5, P_SKIP = 6 (common/macroblock.h:64-70) */
    int32_t partition;      /* D_16x8 = 14, D_8x16 = 15, D_16x16 = 16; D_8x8 = 13 for P_8x8 */
    uint8_t sub[4];         /* D_L0_4x4 = 0, D_L0_8x4 = 1, D_L0_4x8 = 2, D_L0_8x8 = 3 */
    int8_t  ref[4];         /* per 8x8 block */
    int16_t mv[16][2];      /* per 4x4 block, block_idx order (x264_scan8) */
} pcamv_mvrec;

typedef struct
{
    int picture;            /* index of the picture in the stream (coding order = display order: no B frames) */
    int n_mb, mb_w, mb_h, qp, cabac, is_p;
    pcamv_mvrec *mb;
} pcamv_picture;

/* ---- bit reader over an unescaped RBSP ---------------------------------------------------------------------------------- */
typedef struct { const uint8_t *buf; int64_t pos, size, stop; int err; } bsr_t;

static inline int br_bit( bsr_t *b )
{
    int v;
    if( b->pos >= b->size ) { b->err = 1; b->pos++; return 0; }
    v = ( b->buf[b->pos >> 3] >> ( 7 - ( b->pos & 7 ) ) ) & 1;
    b->pos++;
    return v;
}
static uint32_t br_u( bsr_t *b, int n ) { uint32_t v = 0; while( n-- > 0 ) v = ( v << 1 ) | br_bit( b ); return v; }
static uint32_t br_peek( bsr_t *b, int n )
{
    const int64_t pos = b->pos; const int err = b->err;
    const uint32_t v = br_u( b, n );
    b->pos = pos; b->err = err;
    return v;
}
static uint32_t br_ue( bsr_t *b )
{
    int z = 0;
    while( !br_bit( b ) && z < 32 && !b->err ) z++;
    return z >= 32 ? 0xffffffffu : ( ( 1u << z ) - 1 ) + br_u( b, z );
}
static int32_t br_se( bsr_t *b ) { const uint32_t k = br_ue( b ); return ( k & 1 ) ? (int32_t)( ( k + 1 ) >> 1 ) : -(int32_t)( k >> 1 ); }
static int br_te( bsr_t *b, int max ) { return max == 1 ? !br_bit( b ) : (int)br_ue( b ); }
static int br_more_rbsp_data( const bsr_t *b ) { return b->pos < b->stop; }

/* ---- parameter sets ------------------------------------------------------------------------------------------------------ */
typedef struct { int valid, profile, log2_max_frame_num, poc_type, log2_max_poc_lsb, delta_always_zero, mb_w, mb_h, frame_mbs_only; } sps_t;
typedef struct { int valid, sps_id, cabac, pic_order, num_ref_l0, weighted_pred, init_qp, deblock_control, redundant, transform8x8; } pps_t;

/* ---- decoder state -------------------------------------------------------------------------------------------------------- */
typedef struct
{
    sps_t sps[32];
    pps_t pps[256];
    char err[256];
    /* picture geometry and the whole-picture arrays the contexts and the vector prediction read */
    int mb_w, mb_h, n_mb, w4, h4;
    int8_t *ref4;              /* [h4*w4]  reference of the 4x4 block */
    int16_t (*mv4)[2];         /* [h4*w4] */
    int16_t (*mvd4)[2];        /* [h4*w4]  vector differences (CABAC contexts, encoder/cabac.c:397-402) */
    uint8_t *nnz;              /* [h4*w4]  luma: coded_block_flag (CABAC) / total_coeff (CAVLC) */
    uint8_t *nnzc[2];          /* [2*mb_h][2*mb_w] chroma AC, per plane */
    int32_t *cbp;              /* per macroblock: luma | chroma << 4 | dc_u << 9 | dc_v << 10, as the encoder keeps it (h->mb.cbp) */
    uint8_t *skip;
    pcamv_mvrec *rec;
    /* current macroblock */
    int cur_mb, mb_x, mb_y;
    uint8_t filled[16];        /* raster 4x4 blocks of the current macroblock whose vector is decoded */
    int num_ref, last_dqp;
    /* arithmetic decoder */
    x264_cabac_t cb;           /* only .state is used: filled by the encoder's own x264_cabac_context_init */
    uint32_t range, offset;
    bsr_t *br;
} dec_t;

static int fail( dec_t *d, const char *msg ) { if( !d->err[0] ) snprintf( d->err, sizeof(d->err), "%s", msg ); return -1; }

static const uint8_t blk_x[16] = { 0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3 };    /* block_idx -> 4x4 column / row */
static const uint8_t blk_y[16] = { 0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3 };

/* ---- CABAC: the inverse of common/cabac.c:861-925 ------------------------------------------------------------------------- */
static void cd_init( dec_t *d )
{
    d->range = 510;                              /* i_range = 0x01FE (common/cabac.c:807-815) */
    d->offset = br_u( d->br, 9 );
}
static inline void cd_renorm( dec_t *d )
{
    while( d->range < 256 ) { d->range <<= 1; d->offset = ( d->offset << 1 ) | br_bit( d->br ); }
}
static int cd_decision( dec_t *d, int ctx )
{
    const int s = d->cb.state[ctx];
    const uint32_t lps = x264_cabac_range_lps[s][( d->range >> 6 ) & 3];
    int bin = s >> 6;                            /* states 64..127 have MPS 1 (common/cabac.c:866) */
    d->range -= lps;
    if( d->offset >= d->range )
    {
        d->offset -= d->range;
        d->range = lps;
        bin ^= 1;
    }
    d->cb.state[ctx] = x264_cabac_transition[s][bin];
    cd_renorm( d );
    return bin;
}
static int cd_bypass( dec_t *d )
{
    d->offset = ( d->offset << 1 ) | br_bit( d->br );
    if( d->offset >= d->range ) { d->offset -= d->range; return 1; }
    return 0;
}
static int cd_terminate( dec_t *d )
{
    d->range -= 2;
    if( d->offset >= d->range ) return 1;
    cd_renorm( d );
    return 0;
}
static int cd_ue_bypass( dec_t *d, int k )       /* x264_cabac_encode_ue_bypass read back: (k' - k) ones, a zero, k' bits */
{
    int v = 0;
    while( cd_bypass( d ) ) { v += 1 << k; if( ++k > 24 ) { d->br->err = 1; return 0; } }
    while( k-- ) v += cd_bypass( d ) << k;
    return v;
}

/* ---- motion vector prediction (common/macroblock.c:28-163) --------------------------------------------------------------- */
static void mv_neighbour( const dec_t *d, int x4, int y4, int *ref, int16_t mv[2] )
{
    int mb;
    *ref = -2; mv[0] = mv[1] = 0;
    if( x4 < 0 || y4 < 0 || x4 >= d->w4 || y4 >= d->h4 ) return;
    mb = ( y4 >> 2 ) * d->mb_w + ( x4 >> 2 );
    if( mb > d->cur_mb || ( mb == d->cur_mb && !d->filled[( y4 & 3 ) * 4 + ( x4 & 3 )] ) ) return;
    *ref = d->ref4[y4 * d->w4 + x4];
    mv[0] = d->mv4[y4 * d->w4 + x4][0]; mv[1] = d->mv4[y4 * d->w4 + x4][1];
}
static inline int median3( int a, int b, int c )
{
    const int mn = a < b ? a : b, mx = a < b ? b : a;
    return c < mn ? mn : c > mx ? mx : c;
}
/* hint: 0 none, 1 / 2 upper / lower 16x8, 3 / 4 left / right 8x16 */
static void predict_mv( const dec_t *d, int x4, int y4, int w, int ref, int hint, int16_t mvp[2] )
{
    int ra, rb, rc, n;
    int16_t a[2], b[2], c[2];
    mv_neighbour( d, x4 - 1, y4, &ra, a );
    mv_neighbour( d, x4, y4 - 1, &rb, b );
    mv_neighbour( d, x4 + w, y4 - 1, &rc, c );
    if( rc == -2 ) mv_neighbour( d, x4 - 1, y4 - 1, &rc, c );
    if( hint == 1 && rb == ref ) { mvp[0] = b[0]; mvp[1] = b[1]; return; }
    if( hint == 2 && ra == ref ) { mvp[0] = a[0]; mvp[1] = a[1]; return; }
    if( hint == 3 && ra == ref ) { mvp[0] = a[0]; mvp[1] = a[1]; return; }
    if( hint == 4 && rc == ref ) { mvp[0] = c[0]; mvp[1] = c[1]; return; }
    n = ( ra == ref ) + ( rb == ref ) + ( rc == ref );
    if( n == 1 )
    {
        const int16_t *s = ra == ref ? a : rb == ref ? b : c;
        mvp[0] = s[0]; mvp[1] = s[1];
    }
    else if( n == 0 && rb == -2 && rc == -2 && ra != -2 ) { mvp[0] = a[0]; mvp[1] = a[1]; }
    else { mvp[0] = median3( a[0], b[0], c[0] ); mvp[1] = median3( a[1], b[1], c[1] ); }
}
static void predict_mv_pskip( const dec_t *d, int16_t mv[2] )
{
    const int x4 = 4 * d->mb_x, y4 = 4 * d->mb_y;
    int ra, rb;
    int16_t a[2], b[2];
    mv_neighbour( d, x4 - 1, y4, &ra, a );
    mv_neighbour( d, x4, y4 - 1, &rb, b );
    if( ra == -2 || rb == -2 || ( ra == 0 && !a[0] && !a[1] ) || ( rb == 0 && !b[0] && !b[1] ) ) { mv[0] = mv[1] = 0; return; }
    predict_mv( d, x4, y4, 4, 0, 0, mv );
}
/* a decoded partition: x, y, w, h in 4x4 units inside the macroblock */
static void store_mv( dec_t *d, int x, int y, int w, int h, const int16_t mv[2], const int16_t mvd[2] )
{
    int i, j;
    for( j = y; j < y + h; j++ )
        for( i = x; i < x + w; i++ )
        {
            const int p = ( 4 * d->mb_y + j ) * d->w4 + 4 * d->mb_x + i;
            d->mv4[p][0] = mv[0]; d->mv4[p][1] = mv[1];
            d->mvd4[p][0] = mvd[0]; d->mvd4[p][1] = mvd[1];
            d->filled[4 * j + i] = 1;
        }
}
static void store_ref( dec_t *d, int x, int y, int w, int h, int ref )
{
    int i, j;
    for( j = y; j < y + h; j++ )
        for( i = x; i < x + w; i++ )
            d->ref4[( 4 * d->mb_y + j ) * d->w4 + 4 * d->mb_x + i] = (int8_t)ref;
}
static int ref_at( const dec_t *d, int x4, int y4 )       /* for the ref_idx contexts: outside the picture = unavailable */
{
    if( x4 < 0 || y4 < 0 || x4 >= d->w4 || y4 >= d->h4 ) return -2;
    return d->ref4[y4 * d->w4 + x4];
}
static int amvd_at( const dec_t *d, int x4, int y4, int l )
{
    if( x4 < 0 || y4 < 0 || x4 >= d->w4 || y4 >= d->h4 ) return 0;
    return abs( d->mvd4[y4 * d->w4 + x4][l] );
}

/* ---- syntax elements, both entropy coders ---------------------------------------------------------------------------------- */
static int read_ref( dec_t *d, int cabac, int x, int y )            /* encoder/cabac.c:375-395, encoder/cavlc.c bs_write_te */
{
    int ctx, ref = 0;
    if( !cabac ) return br_te( d->br, d->num_ref - 1 );
    ctx = ( ref_at( d, 4 * d->mb_x + x - 1, 4 * d->mb_y + y ) > 0 ) + 2 * ( ref_at( d, 4 * d->mb_x + x, 4 * d->mb_y + y - 1 ) > 0 );
    while( cd_decision( d, 54 + ctx ) )
    {
        ctx = ( ctx >> 2 ) + 4;
        if( ++ref > 31 ) { d->br->err = 1; break; }
    }
    return ref;
}
static int read_mvd_cpn( dec_t *d, int cabac, int x4, int y4, int l )   /* encoder/cabac.c:397-445 */
{
    static const uint8_t ctxes[9] = { 0, 3, 4, 5, 6, 6, 6, 6, 6 };
    int amvd, base, i, v;
    if( !cabac ) return br_se( d->br );
    amvd = amvd_at( d, x4 - 1, y4, l ) + amvd_at( d, x4, y4 - 1, l );
    base = l ? 47 : 40;
    if( !cd_decision( d, base + ( amvd > 2 ) + ( amvd > 32 ) ) ) return 0;
    for( i = 1; i < 9 && cd_decision( d, base + ctxes[i] ); i++ ) ;
    v = i < 9 ? i : 9 + cd_ue_bypass( d, 3 );
    return cd_bypass( d ) ? -v : v;
}
/* one partition: prediction, difference, stores (x, y, w, h in 4x4 units inside the macroblock) */
static void read_partition_mv( dec_t *d, int cabac, int x, int y, int w, int h, int ref, int hint )
{
    const int x4 = 4 * d->mb_x + x, y4 = 4 * d->mb_y + y;
    int16_t mvp[2], mvd[2], mv[2];
    predict_mv( d, x4, y4, w, ref, hint, mvp );
    mvd[0] = (int16_t)read_mvd_cpn( d, cabac, x4, y4, 0 );
    mvd[1] = (int16_t)read_mvd_cpn( d, cabac, x4, y4, 1 );
    mv[0] = (int16_t)( mvp[0] + mvd[0] ); mv[1] = (int16_t)( mvp[1] + mvd[1] );
    store_mv( d, x, y, w, h, mv, mvd );
}

/* ---- residual: nothing is reconstructed, but every bin / bit has to be consumed and the neighbour state kept -------------- */
enum { CAT_LUMA_4x4 = 2, CAT_CHROMA_DC = 3, CAT_CHROMA_AC = 4 };

static int residual_cabac( dec_t *d, int cat, int n_coeff, int nza, int nzb )      /* encoder/cabac.c:584-680; returns coded_block_flag */
{
    static const uint16_t sig_off[5] = { 105, 120, 134, 149, 152 }, last_off[5] = { 166, 181, 195, 210, 213 };
    static const uint16_t lvl_off[5] = { 227, 237, 247, 257, 266 };
    static const uint8_t lvl1_ctx[8] = { 1, 2, 3, 4, 0, 0, 0, 0 }, lvlgt1_ctx[8] = { 5, 5, 5, 5, 6, 7, 8, 9 };
    static const uint8_t trans[2][8] = { { 1, 2, 3, 3, 4, 5, 6, 7 }, { 4, 4, 4, 4, 5, 6, 7, 7 } };
    int i, n = 0, node = 0, last_seen = 0;
    if( !cd_decision( d, 85 + 4 * cat + 2 * !!nzb + !!nza ) ) return 0;
    for( i = 0; i < n_coeff - 1; i++ )
        if( cd_decision( d, sig_off[cat] + i ) )
        {
            n++;
            if( cd_decision( d, last_off[cat] + i ) ) { last_seen = 1; break; }
        }
    if( !last_seen ) n++;                        /* the last position is significant without a flag */
    while( n-- > 0 )
    {
        if( cd_decision( d, lvl_off[cat] + lvl1_ctx[node] ) )
        {
            const int ctx = lvl_off[cat] + lvlgt1_ctx[node];
            int prefix = 1;
            while( prefix < 14 && cd_decision( d, ctx ) ) prefix++;
            if( prefix >= 14 ) cd_ue_bypass( d, 0 );
            node = trans[1][node];
        }
        else
            node = trans[0][node];
        cd_bypass( d );                          /* sign */
        if( d->br->err ) break;
    }
    return 1;
}

static int vlc_match( bsr_t *b, const vlc_t *tab, int n )
{
    const uint32_t pk = br_peek( b, 16 );
    int i;
    for( i = 0; i < n; i++ )
        if( tab[i].i_size && ( pk >> ( 16 - tab[i].i_size ) ) == tab[i].i_bits )
        {
            br_u( b, tab[i].i_size );
            return i;
        }
    b->err = 1;
    return -1;
}
static int residual_cavlc( dec_t *d, int chroma_dc, int n_coeff, int nC )            /* encoder/cavlc.c:112-198; returns total_coeff */
{
    static const uint8_t ct_index[17] = { 0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3 };
    const int t = chroma_dc ? 4 : ct_index[nC > 16 ? 16 : nC];
    bsr_t *b = d->br;
    int total, trailing, suffix_length, i, zeros_left, k;
    {
        /* coeff_token: x264_coeff0_token[t] for an empty block, x264_coeff_token[t][4 * (total - 1) + trailing] otherwise */
        const uint32_t pk = br_peek( b, 16 );
        const vlc_t z = x264_coeff0_token[t];
        if( z.i_size && ( pk >> ( 16 - z.i_size ) ) == z.i_bits ) { br_u( b, z.i_size ); return 0; }
        k = vlc_match( b, x264_coeff_token[t], 16 * 4 );
        if( k < 0 ) return 0;
        total = ( k >> 2 ) + 1; trailing = k & 3;
        if( total > n_coeff || trailing > total ) { b->err = 1; return 0; }
    }
    br_u( b, trailing );                         /* signs of the trailing ones */
    suffix_length = total > 10 && trailing < 3;
    for( i = trailing; i < total; i++ )
    {
        int prefix = 0, code, size, level;
        while( !br_bit( b ) ) if( ++prefix > 31 || b->err ) { b->err = 1; return total; }
        code = ( prefix < 15 ? prefix : 15 ) << suffix_length;
        if( suffix_length > 0 || prefix >= 14 )
        {
            size = ( prefix == 14 && suffix_length == 0 ) ? 4 : prefix >= 15 ? prefix - 3 : suffix_length;
            code += br_u( b, size );
        }
        if( prefix >= 15 && suffix_length == 0 ) code += 15;
        if( prefix >= 16 ) code += ( 1 << ( prefix - 3 ) ) - 4096;
        if( i == trailing && trailing < 3 ) code += 2;
        level = ( code & 1 ) ? ( -code - 1 ) >> 1 : ( code + 2 ) >> 1;
        if( suffix_length == 0 ) suffix_length = 1;
        if( abs( level ) > ( 3 << ( suffix_length - 1 ) ) && suffix_length < 6 ) suffix_length++;
    }
    if( total < n_coeff )
    {
        zeros_left = chroma_dc ? vlc_match( b, x264_total_zeros_dc[total - 1], 4 ) : vlc_match( b, x264_total_zeros[total - 1], 16 );
        if( zeros_left < 0 ) return total;
    }
    else
        zeros_left = 0;
    for( i = 0; i < total - 1 && zeros_left > 0; i++ )
    {
        const int run = vlc_match( b, x264_run_before[zeros_left - 1 < 6 ? zeros_left - 1 : 6], 16 );
        if( run < 0 || run > zeros_left ) { b->err = 1; return total; }
        zeros_left -= run;
    }
    return total;
}

/* neighbouring coefficient state of a luma 4x4 block / chroma AC block; outside the picture: `outside` */
static int nnz_luma( const dec_t *d, int x4, int y4, int outside )
{
    if( x4 < 0 || y4 < 0 ) return outside;
    return d->nnz[y4 * d->w4 + x4];
}
static int nnz_chroma( const dec_t *d, int pl, int x2, int y2, int outside )
{
    if( x2 < 0 || y2 < 0 ) return outside;
    return d->nnzc[pl][y2 * 2 * d->mb_w + x2];
}
static int cavlc_nc( int na, int nb )            /* x264_mb_predict_non_zero_code (common/macroblock.h:445-457); -1 = not available */
{
    if( na >= 0 && nb >= 0 ) return ( na + nb + 1 ) >> 1;
    return na >= 0 ? na : nb >= 0 ? nb : 0;
}

static void read_residual( dec_t *d, int cabac, int cbp_luma, int cbp_chroma, int32_t *cbp_out )
{
    int i, pl;
    for( i = 0; i < 16; i++ )
    {
        const int x4 = 4 * d->mb_x + blk_x[i], y4 = 4 * d->mb_y + blk_y[i];
        int v = 0;
        if( cbp_luma & ( 1 << ( i >> 2 ) ) )
            v = cabac ? residual_cabac( d, CAT_LUMA_4x4, 16, nnz_luma( d, x4 - 1, y4, 0 ), nnz_luma( d, x4, y4 - 1, 0 ) )
                      : residual_cavlc( d, 0, 16, cavlc_nc( nnz_luma( d, x4 - 1, y4, -1 ), nnz_luma( d, x4, y4 - 1, -1 ) ) );
        d->nnz[y4 * d->w4 + x4] = (uint8_t)v;
    }
    if( cbp_chroma )
        for( pl = 0; pl < 2; pl++ )
        {
            if( cabac )
            {
                /* coded_block_flag context of a chroma DC block: the neighbours' DC flags, 0 outside the picture for an inter
                 * macroblock (encoder/cabac.c:527-534) */
                const int l = d->mb_x > 0 ? ( d->cbp[d->cur_mb - 1] >> ( 9 + pl ) ) & 1 : 0;
                const int t = d->mb_y > 0 ? ( d->cbp[d->cur_mb - d->mb_w] >> ( 9 + pl ) ) & 1 : 0;
                if( residual_cabac( d, CAT_CHROMA_DC, 4, l, t ) ) *cbp_out |= 0x200 << pl;
            }
            else
                residual_cavlc( d, 1, 4, 0 );
        }
    for( pl = 0; pl < 2; pl++ )
        for( i = 0; i < 4; i++ )
        {
            const int x2 = 2 * d->mb_x + ( i & 1 ), y2 = 2 * d->mb_y + ( i >> 1 );
            int v = 0;
            if( cbp_chroma & 2 )
                v = cabac ? residual_cabac( d, CAT_CHROMA_AC, 15, nnz_chroma( d, pl, x2 - 1, y2, 0 ), nnz_chroma( d, pl, x2, y2 - 1, 0 ) )
                          : residual_cavlc( d, 0, 15, cavlc_nc( nnz_chroma( d, pl, x2 - 1, y2, -1 ), nnz_chroma( d, pl, x2, y2 - 1, -1 ) ) );
            d->nnzc[pl][y2 * 2 * d->mb_w + x2] = (uint8_t)v;
        }
}

/* ---- macroblock layer ------------------------------------------------------------------------------------------------------- */
static void begin_mb( dec_t *d, int mb )
{
    d->cur_mb = mb; d->mb_x = mb % d->mb_w; d->mb_y = mb / d->mb_w;
    memset( d->filled, 0, sizeof(d->filled) );
}
static void finish_rec( dec_t *d, int type, int partition, const int sub[4] )
{
    pcamv_mvrec *r = &d->rec[d->cur_mb];
    int i;
    r->type = type; r->partition = partition;
    for( i = 0; i < 4; i++ )
    {
        r->sub[i] = (uint8_t)sub[i];
        r->ref[i] = d->ref4[( 4 * d->mb_y + 2 * ( i >> 1 ) ) * d->w4 + 4 * d->mb_x + 2 * ( i & 1 )];
    }
    for( i = 0; i < 16; i++ )
    {
        const int p = ( 4 * d->mb_y + blk_y[i] ) * d->w4 + 4 * d->mb_x + blk_x[i];
        r->mv[i][0] = d->mv4[p][0]; r->mv[i][1] = d->mv4[p][1];
    }
}
static void skip_mb( dec_t *d )
{
    static const int16_t zero[2] = { 0, 0 };
    static const int sub[4] = { D_L0_8x8, D_L0_8x8, D_L0_8x8, D_L0_8x8 };
    int16_t mv[2];
    int i, pl;
    predict_mv_pskip( d, mv );
    store_ref( d, 0, 0, 4, 4, 0 );
    store_mv( d, 0, 0, 4, 4, mv, zero );
    for( i = 0; i < 16; i++ ) d->nnz[( 4 * d->mb_y + ( i >> 2 ) ) * d->w4 + 4 * d->mb_x + ( i & 3 )] = 0;
    for( pl = 0; pl < 2; pl++ )
        for( i = 0; i < 4; i++ ) d->nnzc[pl][( 2 * d->mb_y + ( i >> 1 ) ) * 2 * d->mb_w + 2 * d->mb_x + ( i & 1 )] = 0;
    d->cbp[d->cur_mb] = 0;
    d->skip[d->cur_mb] = 1;
    d->last_dqp = 0;
    finish_rec( d, P_SKIP, D_16x16, sub );
}
static int macroblock_layer( dec_t *d, int cabac )
{
    static const int8_t golomb_to_sub[4] = { D_L0_8x8, D_L0_8x4, D_L0_4x8, D_L0_4x4 };    /* inverse of sub_mb_type_p_to_golomb (encoder/cavlc.c:55-58) */
    int sub[4] = { D_L0_8x8, D_L0_8x8, D_L0_8x8, D_L0_8x8 };
    int type, partition, ref0_only = 0, i, cbp_luma, cbp_chroma;
    int32_t cbp;
    d->skip[d->cur_mb] = 0;
    /* mb_type (encoder/cabac.c:84-118, encoder/cavlc.c:371-440) */
    if( cabac )
    {
        if( cd_decision( d, 14 ) ) return fail( d, "intra macroblock in a P slice (the PCAMV encoder never writes one)" );
        if( !cd_decision( d, 15 ) )
        {
            if( !cd_decision( d, 16 ) ) { type = P_L0; partition = D_16x16; }
            else { type = P_8x8; partition = D_8x8; }
        }
        else
        {
            type = P_L0;
            partition = cd_decision( d, 17 ) ? D_16x8 : D_8x16;
        }
    }
    else
    {
        const uint32_t t = br_ue( d->br );
        if( t > 4 ) return fail( d, "intra macroblock in a P slice (the PCAMV encoder never writes one)" );
        type = t < 3 ? P_L0 : P_8x8;
        partition = t == 0 ? D_16x16 : t == 1 ? D_16x8 : t == 2 ? D_8x16 : D_8x8;
        ref0_only = t == 4;                       /* P_8x8ref0 */
    }
    if( type == P_L0 )
    {
        const int n = partition == D_16x16 ? 1 : 2;
        int ref[2] = { 0, 0 };
        for( i = 0; i < n; i++ )
        {
            const int x = partition == D_8x16 ? 2 * i : 0, y = partition == D_16x8 ? 2 * i : 0;
            const int w = partition == D_8x16 ? 2 : 4, h = partition == D_16x8 ? 2 : 4;
            if( d->num_ref > 1 ) ref[i] = read_ref( d, cabac, x, y );
            if( ref[i] >= d->num_ref ) return fail( d, "ref_idx out of range" );
            store_ref( d, x, y, n == 1 ? 4 : w, n == 1 ? 4 : h, ref[i] );
        }
        if( partition == D_16x16 ) read_partition_mv( d, cabac, 0, 0, 4, 4, ref[0], 0 );
        else if( partition == D_16x8 ) { read_partition_mv( d, cabac, 0, 0, 4, 2, ref[0], 1 ); read_partition_mv( d, cabac, 0, 2, 4, 2, ref[1], 2 ); }
        else { read_partition_mv( d, cabac, 0, 0, 2, 4, ref[0], 3 ); read_partition_mv( d, cabac, 2, 0, 2, 4, ref[1], 4 ); }
    }
    else
    {
        int ref[4] = { 0, 0, 0, 0 };
        for( i = 0; i < 4; i++ )
        {
            if( cabac )                           /* encoder/cabac.c:309-330 */
                sub[i] = cd_decision( d, 21 ) ? D_L0_8x8 : !cd_decision( d, 22 ) ? D_L0_8x4 : cd_decision( d, 23 ) ? D_L0_4x8 : D_L0_4x4;
            else
            {
                const uint32_t t = br_ue( d->br );
                if( t > 3 ) return fail( d, "sub_mb_type out of range" );
                sub[i] = golomb_to_sub[t];
            }
        }
        for( i = 0; i < 4; i++ )
        {
            if( d->num_ref > 1 && !ref0_only ) ref[i] = read_ref( d, cabac, 2 * ( i & 1 ), 2 * ( i >> 1 ) );
            if( ref[i] >= d->num_ref ) return fail( d, "ref_idx out of range" );
            store_ref( d, 2 * ( i & 1 ), 2 * ( i >> 1 ), 2, 2, ref[i] );
        }
        for( i = 0; i < 4; i++ )
        {
            const int x = 2 * ( i & 1 ), y = 2 * ( i >> 1 );
            if( sub[i] == D_L0_8x8 ) read_partition_mv( d, cabac, x, y, 2, 2, ref[i], 0 );
            else if( sub[i] == D_L0_8x4 ) { read_partition_mv( d, cabac, x, y, 2, 1, ref[i], 0 ); read_partition_mv( d, cabac, x, y + 1, 2, 1, ref[i], 0 ); }
            else if( sub[i] == D_L0_4x8 ) { read_partition_mv( d, cabac, x, y, 1, 2, ref[i], 0 ); read_partition_mv( d, cabac, x + 1, y, 1, 2, ref[i], 0 ); }
            else
            {
                read_partition_mv( d, cabac, x, y, 1, 1, ref[i], 0 ); read_partition_mv( d, cabac, x + 1, y, 1, 1, ref[i], 0 );
                read_partition_mv( d, cabac, x, y + 1, 1, 1, ref[i], 0 ); read_partition_mv( d, cabac, x + 1, y + 1, 1, 1, ref[i], 0 );
            }
        }
    }
    /* coded_block_pattern (encoder/cabac.c:233-263, encoder/cavlc.c:42-47,572) */
    if( cabac )
    {
        const int cbp_l = d->mb_x > 0 ? d->cbp[d->cur_mb - 1] : -1, cbp_t = d->mb_y > 0 ? d->cbp[d->cur_mb - d->mb_w] : -1;
        int ctx;
        cbp_luma = 0;
        cbp_luma |= cd_decision( d, 76 - ( ( cbp_l >> 1 ) & 1 ) - ( ( cbp_t >> 1 ) & 2 ) );
        cbp_luma |= cd_decision( d, 76 - ( ( cbp_luma >> 0 ) & 1 ) - ( ( cbp_t >> 2 ) & 2 ) ) << 1;
        cbp_luma |= cd_decision( d, 76 - ( ( cbp_l >> 3 ) & 1 ) - ( ( cbp_luma << 1 ) & 2 ) ) << 2;
        cbp_luma |= cd_decision( d, 76 - ( ( cbp_luma >> 2 ) & 1 ) - ( ( cbp_luma >> 0 ) & 2 ) ) << 3;
        ctx = ( ( cbp_l & 0x30 ) && cbp_l != -1 ) + 2 * ( ( cbp_t & 0x30 ) && cbp_t != -1 );
        cbp_chroma = 0;
        if( cd_decision( d, 77 + ctx ) )
        {
            ctx = 4 + ( ( cbp_l & 0x30 ) == 0x20 ) + 2 * ( ( cbp_t & 0x30 ) == 0x20 );
            cbp_chroma = 1 + cd_decision( d, 77 + ctx );
        }
    }
    else
    {
        /* codeNum -> coded_block_pattern, inter column of H.264 table 9-4 (the encoder holds the opposite direction in a static
         * of encoder/cavlc.c:42-47, inter_cbp_to_golomb) */
        static const uint8_t inter_cbp_by_code[48] = {
            0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44,
            33, 34, 36, 40, 39, 43, 45, 46, 17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41 };
        const uint32_t t = br_ue( d->br );
        if( t > 47 ) return fail( d, "coded_block_pattern out of range" );
        cbp_luma = inter_cbp_by_code[t] & 15; cbp_chroma = inter_cbp_by_code[t] >> 4;
    }
    cbp = cbp_luma | ( cbp_chroma << 4 );
    d->cbp[d->cur_mb] = cbp;                      /* the chroma DC flags join below; the neighbours' entries are what the contexts read */
    if( cbp_luma || cbp_chroma )
    {
        /* mb_qp_delta (encoder/cabac.c:265-296, encoder/cavlc.c:200-222) */
        int dqp = 0;
        if( cabac )
        {
            int ctx = d->last_dqp != 0, v = 0;
            while( cd_decision( d, 60 + ctx ) ) { ctx = 2 + ( ctx >> 1 ); if( ++v > 104 ) { d->br->err = 1; break; } }
            dqp = ( v & 1 ) ? ( v + 1 ) >> 1 : -( v >> 1 );
        }
        else
            dqp = br_se( d->br );
        d->last_dqp = dqp;
        read_residual( d, cabac, cbp_luma, cbp_chroma, &cbp );
        d->cbp[d->cur_mb] = cbp;
    }
    else
    {
        d->last_dqp = 0;
        read_residual( d, cabac, 0, 0, &cbp );    /* clears the macroblock's coefficient state */
    }
    finish_rec( d, type, partition, sub );
    return d->br->err ? fail( d, "bitstream ends inside a macroblock" ) : 0;
}

/* ---- NAL units ---------------------------------------------------------------------------------------------------------------- */
static void parse_sps( dec_t *d, bsr_t *b )
{
    sps_t s;
    uint32_t id;
    memset( &s, 0, sizeof(s) );
    s.profile = br_u( b, 8 ); br_u( b, 8 ); br_u( b, 8 );
    id = br_ue( b );
    if( id > 31 ) return;
    if( s.profile >= 100 )
    {
        if( br_ue( b ) != 1 ) return;             /* chroma_format_idc: 4:2:0 only */
        br_ue( b ); br_ue( b ); br_u( b, 1 );
        if( br_u( b, 1 ) ) return;                /* scaling matrices: not written for the flat matrix (encoder/set.c:234) */
    }
    s.log2_max_frame_num = br_ue( b ) + 4;
    s.poc_type = br_ue( b );
    if( s.poc_type == 0 ) s.log2_max_poc_lsb = br_ue( b ) + 4;
    if( (unsigned)s.log2_max_frame_num > 16 || (unsigned)s.log2_max_poc_lsb > 16 ) return;
    else if( s.poc_type == 1 )
    {
        uint32_t n, i;
        s.delta_always_zero = br_u( b, 1 ); br_se( b ); br_se( b );
        n = br_ue( b );
        for( i = 0; i < n && i < 256; i++ ) br_se( b );
    }
    br_ue( b ); br_u( b, 1 );
    s.mb_w = br_ue( b ) + 1;
    s.mb_h = br_ue( b ) + 1;
    s.frame_mbs_only = br_u( b, 1 );
    s.valid = !b->err && s.mb_w >= 1 && s.mb_w <= 1024 && s.mb_h >= 1 && s.mb_h <= 1024;     /* an untrusted stream sizes the arrays */
    d->sps[id] = s;
}
static void parse_pps( dec_t *d, bsr_t *b )
{
    pps_t p;
    uint32_t id;
    memset( &p, 0, sizeof(p) );
    id = br_ue( b );
    if( id > 255 ) return;
    p.sps_id = br_ue( b );
    p.cabac = br_u( b, 1 );
    p.pic_order = br_u( b, 1 );
    if( br_ue( b ) != 0 ) return;                 /* slice groups */
    p.num_ref_l0 = (int)( br_ue( b ) & 0xff ) + 1; br_ue( b );
    p.weighted_pred = br_u( b, 1 ); br_u( b, 2 );
    p.init_qp = 26 + br_se( b ); br_se( b ); br_se( b );
    p.deblock_control = br_u( b, 1 ); br_u( b, 1 );
    p.redundant = br_u( b, 1 );
    if( br_more_rbsp_data( b ) ) p.transform8x8 = br_u( b, 1 );
    p.valid = !b->err && (unsigned)p.sps_id < 32;
    d->pps[id] = p;
}
static void free_picture_arrays( dec_t *d )
{
    free( d->ref4 ); free( d->mv4 ); free( d->mvd4 ); free( d->nnz ); free( d->nnzc[0] ); free( d->nnzc[1] ); free( d->cbp ); free( d->skip );
    d->ref4 = NULL; d->mv4 = NULL; d->mvd4 = NULL; d->nnz = NULL; d->nnzc[0] = d->nnzc[1] = NULL; d->cbp = NULL; d->skip = NULL;
}
static int size_picture_arrays( dec_t *d, const sps_t *s )
{
    if( d->ref4 && d->mb_w == s->mb_w && d->mb_h == s->mb_h ) return 0;
    free_picture_arrays( d );
    d->mb_w = s->mb_w; d->mb_h = s->mb_h; d->n_mb = s->mb_w * s->mb_h; d->w4 = 4 * s->mb_w; d->h4 = 4 * s->mb_h;
    d->ref4 = malloc( d->w4 * d->h4 );
    d->mv4 = malloc( (size_t)d->w4 * d->h4 * 4 );
    d->mvd4 = malloc( (size_t)d->w4 * d->h4 * 4 );
    d->nnz = malloc( d->w4 * d->h4 );
    d->nnzc[0] = malloc( 4 * d->n_mb ); d->nnzc[1] = malloc( 4 * d->n_mb );
    d->cbp = malloc( d->n_mb * sizeof(int32_t) );
    d->skip = malloc( d->n_mb );
    return d->ref4 && d->mv4 && d->mvd4 && d->nnz && d->nnzc[0] && d->nnzc[1] && d->cbp && d->skip ? 0 : fail( d, "out of memory" );
}

/* one slice NAL; returns 1 with *pic filled (pic->mb malloc'ed for a P picture), 0 for NALs that are no picture, -1 on error */
static int parse_slice( dec_t *d, bsr_t *b, int nal_type, int nal_ref_idc, pcamv_picture *pic )
{
    const sps_t *s;
    const pps_t *p;
    uint32_t first_mb, slice_type, pps_id;
    int qp, cabac_init_idc = 0, mb, i;
    first_mb = br_ue( b );
    slice_type = br_ue( b );
    pps_id = br_ue( b );
    if( pps_id > 255 || !d->pps[pps_id].valid || !d->sps[d->pps[pps_id].sps_id].valid ) return fail( d, "slice refers to a missing parameter set" );
    p = &d->pps[pps_id]; s = &d->sps[p->sps_id];
    if( first_mb != 0 ) return fail( d, "several slices per picture are not supported" );
    if( !s->frame_mbs_only ) return fail( d, "interlaced streams are not supported" );
    if( slice_type >= 5 ) slice_type -= 5;
    memset( pic, 0, sizeof(*pic) );
    pic->mb_w = s->mb_w; pic->mb_h = s->mb_h; pic->n_mb = s->mb_w * s->mb_h; pic->cabac = p->cabac;
    if( slice_type == 2 ) return 1;               /* I slice: a picture without vectors */
    if( slice_type != 0 ) return fail( d, "only I and P slices are supported" );
    if( p->transform8x8 ) return fail( d, "8x8 transform is not supported" );
    if( p->weighted_pred ) return fail( d, "weighted prediction is not supported" );
    br_u( b, s->log2_max_frame_num );
    if( nal_type == 5 ) br_ue( b );
    if( s->poc_type == 0 )
    {
        br_u( b, s->log2_max_poc_lsb );
        if( p->pic_order ) br_se( b );
    }
    else if( s->poc_type == 1 && !s->delta_always_zero )
    {
        br_se( b );
        if( p->pic_order ) br_se( b );
    }
    if( p->redundant ) br_ue( b );
    d->num_ref = p->num_ref_l0;
    if( br_u( b, 1 ) ) d->num_ref = (int)( br_ue( b ) & 0xff ) + 1;
    if( d->num_ref < 1 || d->num_ref > 32 ) return fail( d, "num_ref_idx_active out of range" );
    if( br_u( b, 1 ) )                            /* ref_pic_list_reordering: only the order of the reference list, not its length */
        for( i = 0; i < 66; i++ )
        {
            if( br_ue( b ) == 3 ) break;
            br_ue( b );
        }
    if( nal_ref_idc )
    {
        if( nal_type == 5 ) br_u( b, 2 );
        else if( br_u( b, 1 ) ) return fail( d, "adaptive reference marking is not supported" );
    }
    if( p->cabac ) cabac_init_idc = br_ue( b );
    if( cabac_init_idc > 2 ) return fail( d, "cabac_init_idc out of range" );
    qp = p->init_qp + br_se( b );
    if( qp < 0 || qp > 51 ) return fail( d, "slice QP out of range" );
    if( p->deblock_control && br_ue( b ) != 1 ) { br_se( b ); br_se( b ); }
    if( b->err ) return fail( d, "bitstream ends inside a slice header" );
    if( size_picture_arrays( d, s ) ) return -1;
    pic->is_p = 1; pic->qp = qp;
    pic->mb = d->rec = calloc( d->n_mb, sizeof(pcamv_mvrec) );
    if( !pic->mb ) return fail( d, "out of memory" );
    memset( d->nnz, 0, d->w4 * d->h4 ); memset( d->nnzc[0], 0, 4 * d->n_mb ); memset( d->nnzc[1], 0, 4 * d->n_mb );
    memset( d->mvd4, 0, (size_t)d->w4 * d->h4 * 4 ); memset( d->skip, 0, d->n_mb );
    d->last_dqp = 0; d->br = b;
    mb = 0;
    if( p->cabac )
    {
        while( b->pos & 7 ) br_bit( b );          /* cabac_alignment_one_bit */
        x264_cabac_context_init( &d->cb, SLICE_TYPE_P, qp, cabac_init_idc );
        cd_init( d );
        for( ;; mb++ )
        {
            int ctx;
            if( mb >= d->n_mb ) return fail( d, "slice data runs past the last macroblock" );
            begin_mb( d, mb );
            ctx = 11 + ( d->mb_x > 0 && !d->skip[mb - 1] ) + ( d->mb_y > 0 && !d->skip[mb - d->mb_w] );      /* encoder/cabac.c:300-306 */
            if( cd_decision( d, ctx ) ) skip_mb( d );
            else if( macroblock_layer( d, 1 ) ) return -1;
            if( b->err ) return fail( d, "bitstream ends inside the slice data" );
            if( cd_terminate( d ) ) break;
        }
        mb++;
    }
    else
    {
        int more = 1;
        while( more )
        {
            uint32_t run = br_ue( b );
            if( run > (uint32_t)( d->n_mb - mb ) ) return fail( d, "mb_skip_run runs past the last macroblock" );
            for( ; run > 0; run--, mb++ ) { begin_mb( d, mb ); skip_mb( d ); more = br_more_rbsp_data( b ); }
            if( more )
            {
                if( mb >= d->n_mb ) return fail( d, "slice data runs past the last macroblock" );
                begin_mb( d, mb );
                if( macroblock_layer( d, 0 ) ) return -1;
                mb++;
                more = br_more_rbsp_data( b );
            }
            if( b->err ) return fail( d, "bitstream ends inside the slice data" );
        }
    }
    if( mb != d->n_mb ) return fail( d, "slice does not cover the picture" );
    return 1;
}

/* ---- stream walker ------------------------------------------------------------------------------------------------------------ */
typedef int (*pcamv_picture_fn)( void *user, const pcamv_picture *pic );

/* calls `fn` for every picture of the Annex-B stream; returns the number of pictures, or -1 (message in err) */
static int pcamv_bitstream_walk( const uint8_t *data, size_t size, pcamv_picture_fn fn, void *user, char *err, size_t err_size )
{
    dec_t *d = calloc( 1, sizeof(*d) );
    uint8_t *rbsp = malloc( size + 8 );
    size_t p = 0;
    int pictures = 0, rc = 0;
    if( !d || !rbsp ) { snprintf( err, err_size, "out of memory" ); free( d ); free( rbsp ); return -1; }
    while( p + 3 < size && rc >= 0 )
    {
        size_t q, n = 0, i;
        int nal_type, nal_ref_idc, zeros = 0;
        bsr_t b;
        if( !( data[p] == 0 && data[p + 1] == 0 && data[p + 2] == 1 ) ) { p++; continue; }
        p += 3;
        for( q = p; q + 2 < size && !( data[q] == 0 && data[q + 1] == 0 && data[q + 2] <= 1 ); q++ ) ;
        if( q + 2 >= size ) q = size;
        if( q <= p ) continue;
        nal_ref_idc = ( data[p] >> 5 ) & 3; nal_type = data[p] & 31;
        for( i = p + 1; i < q; i++ )               /* drop the emulation prevention bytes (common/common.c:679-691) */
        {
            if( zeros >= 2 && data[i] == 3 ) { zeros = 0; continue; }
            zeros = data[i] == 0 ? zeros + 1 : 0;
            rbsp[n++] = data[i];
        }
        while( n > 0 && rbsp[n - 1] == 0 ) n--;     /* cabac_zero_words / trailing zero bytes */
        b.buf = rbsp; b.pos = 0; b.size = 8 * (int64_t)n; b.err = 0; b.stop = b.size;
        if( n > 0 )
        {
            int k = 0;
            while( !( ( rbsp[n - 1] >> k ) & 1 ) ) k++;
            b.stop = 8 * (int64_t)n - 1 - k;       /* position of the rbsp_stop_one_bit */
        }
        if( nal_type == 7 ) parse_sps( d, &b );
        else if( nal_type == 8 ) parse_pps( d, &b );
        else if( nal_type == 1 || nal_type == 5 )
        {
            pcamv_picture pic;
            memset( &pic, 0, sizeof(pic) );
            rc = parse_slice( d, &b, nal_type, nal_ref_idc, &pic );
            if( rc == 1 )
            {
                pic.picture = pictures++;
                if( fn && fn( user, &pic ) ) { rc = fail( d, "output failed" ); }
                free( pic.mb );
            }
            else if( rc < 0 )
            {
                char where[64];
                free( pic.mb );
                snprintf( where, sizeof(where), " (picture %d, macroblock %d)", pictures, d->cur_mb );
                strncat( d->err, where, sizeof(d->err) - strlen( d->err ) - 1 );
            }
        }
        p = q;
    }
    if( rc < 0 ) snprintf( err, err_size, "%s", d->err );
    free_picture_arrays( d ); free( d ); free( rbsp );
    return rc < 0 ? -1 : pictures;
}

/* ---- cover order (encoder/encoder.c:1566-1655) ----------------------------------------------------------------------------------
 * One element per vector-carrying partition, macroblocks in raster order; inside a macroblock 16x16: 1; 8x16 / 16x8: 2;
 * P_8x8: per 8x8 block 1 (8x8), 2 (4x8: left, right / 8x4: top, bottom) or 4 (4x4).  `blocks` receives the block_idx of the
 * partition's own top-left 4x4 block.  (The reference reads the bit from info.cache[].mv[slot], which its unsequenced copy fills
 * from another block for some slots - SURVEY fact 3; what pass 2 then WRITES for the partition is what can be read back.) */
static int carrier_blocks( const pcamv_mvrec *r, int blocks[16] )
{
    int n = 0, i;
    if( r->type == P_L0 )
    {
        blocks[n++] = 0;
        if( r->partition == D_8x16 ) blocks[n++] = 4;
        else if( r->partition == D_16x8 ) blocks[n++] = 8;
    }
    else if( r->type == P_8x8 )
        for( i = 0; i < 4; i++ )
        {
            blocks[n++] = 4 * i;
            if( r->sub[i] == D_L0_4x8 ) blocks[n++] = 4 * i + 1;
            else if( r->sub[i] == D_L0_8x4 ) blocks[n++] = 4 * i + 2;
            else if( r->sub[i] == D_L0_4x4 ) { blocks[n++] = 4 * i + 1; blocks[n++] = 4 * i + 2; blocks[n++] = 4 * i + 3; }
        }
    return n;
}
static int picture_stego( const pcamv_picture *pic, uint8_t *stego )
{
    int n = 0, mb, k;
    for( mb = 0; mb < pic->n_mb; mb++ )
    {
        int blocks[16];
        const int c = carrier_blocks( &pic->mb[mb], blocks );
        for( k = 0; k < c; k++ )
            stego[n++] = ( pic->mb[mb].mv[blocks[k]][0] + pic->mb[mb].mv[blocks[k]][1] ) & 1;
    }
    return n;
}

/* ---- CLI ------------------------------------------------------------------------------------------------------------------------ */
typedef struct { FILE *mv, *msg, *stego; float rate; int frames, bits, skipped, carriers, p_pictures, csv; } cli_t;

static int on_picture_mv( void *user, const pcamv_picture *pic )
{
    cli_t *c = user;
    int32_t hd[6] = { pic->picture, pic->n_mb, pic->mb_w, pic->mb_h, pic->qp, pic->cabac };
    if( !pic->is_p ) return 0;
    c->p_pictures++;
    if( c->csv )
    {
        /* one line per macroblock: picture, mb_x, mb_y, type, partition, the four sub-partition types, the four references, then the
         * sixteen vectors (quarter-pel x, y) in block_idx order - the motion field as steganalysis tools want it */
        int mb, i;
        for( mb = 0; mb < pic->n_mb; mb++ )
        {
            const pcamv_mvrec *r = &pic->mb[mb];
            fprintf( c->mv, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d", pic->picture, mb % pic->mb_w, mb / pic->mb_w, r->type, r->partition,
                     r->sub[0], r->sub[1], r->sub[2], r->sub[3], r->ref[0], r->ref[1], r->ref[2], r->ref[3] );
            for( i = 0; i < 16; i++ ) fprintf( c->mv, ",%d,%d", r->mv[i][0], r->mv[i][1] );
            fputc( '\n', c->mv );
        }
        return ferror( c->mv );
    }
    return fwrite( hd, 4, 6, c->mv ) != 6 || (int)fwrite( pic->mb, sizeof(pcamv_mvrec), pic->n_mb, c->mv ) != pic->n_mb;
}
static int on_picture_extract( void *user, const pcamv_picture *pic )
{
    cli_t *c = user;
    uint8_t *stego, *msg;
    int n, an, zeros;
    int32_t oh[2], sh[3];
    if( !pic->is_p ) return 0;
    c->p_pictures++;
    stego = malloc( 16 * pic->n_mb + 1 );
    n = picture_stego( pic, stego );
    /* the message length of a frame as the embedder computes it (encoder/encoder.c:1828-1836): float rate, bits per frame
     * above 1, bits per carrier vector otherwise */
    an = c->rate > 1 ? (int)c->rate : (int)( c->rate * n );
    msg = malloc( ( an > 0 ? an : 0 ) + 1 );
    /* a frame the embedder gave up on (message longer than the cover, syndrome outside the code's range) is recognisable in the
     * stream: its stego vector stays zeroed (embed.h:349-356) and pass 2 still flips every carrier whose cover bit is 1
     * (encoder/encoder.c:1848-1855), so EVERY carrier of the written picture has LSB 0.  Such a picture carries nothing. */
    for( zeros = 0; zeros < n && !stego[zeros]; zeros++ ) ;
    if( an >= 1 && n >= an && n >= 16 && zeros == n )
    {
        fprintf( stderr, "x264 [warning]: picture %d: all %d carriers are even - the embedder gave up on this frame, nothing to extract\n", pic->picture, n );
        an = 0; c->skipped++;
    }
    else if( an < 1 || n < an || pcamv_stc_extract( stego, n, msg, an, 10 ) < 0 )
    {
        if( an >= 1 ) fprintf( stderr, "x264 [warning]: picture %d: cannot extract %d bits from %d carriers, skipped\n", pic->picture, an, n );
        an = 0; c->skipped++;
    }
    oh[0] = pic->picture; oh[1] = an;
    fwrite( oh, 4, 2, c->msg ); fwrite( msg, 1, an, c->msg );
    if( c->stego )
    {
        sh[0] = pic->picture; sh[1] = n; sh[2] = an;
        fwrite( sh, 4, 3, c->stego ); fwrite( stego, 1, n, c->stego );
    }
    c->frames++; c->bits += an; c->carriers += n;
    free( stego ); free( msg );
    return 0;
}
static uint8_t *read_file( const char *path, size_t *size )
{
    FILE *f = fopen( path, "rb" );
    uint8_t *buf;
    long n;
    if( !f ) return NULL;
    fseek( f, 0, SEEK_END ); n = ftell( f ); fseek( f, 0, SEEK_SET );
    buf = malloc( n + 1 );
    if( buf && (long)fread( buf, 1, n, f ) != n ) { free( buf ); buf = NULL; }
    fclose( f );
    *size = n;
    return buf;
}
/* `x264_pcamv --parse-mv IN.264 -o MV.bin [--csv]` and `x264_pcamv --extract-264 IN.264 --emrate R -o MESSAGE.bin [--stego STEGO.bin]` */
int pcamv_bitstream_main( int argc, char **argv )
{
    const int extract = !strcmp( argv[1], "--extract-264" );
    const char *in = argv[2], *out = NULL, *stego = NULL;
    cli_t c;
    uint8_t *data;
    size_t size = 0;
    char err[320];
    int i, n;
    memset( &c, 0, sizeof(c) );
    c.rate = -1;
    for( i = 3; i < argc - 1; i++ )
    {
        if( !strcmp( argv[i], "-o" ) || !strcmp( argv[i], "--output" ) ) out = argv[i + 1];
        else if( !strcmp( argv[i], "--emrate" ) ) c.rate = atof( argv[i + 1] );      /* x264.c:522 */
        else if( !strcmp( argv[i], "--stego" ) ) stego = argv[i + 1];
    }
    for( i = 3; i < argc; i++ )
        if( !strcmp( argv[i], "--csv" ) ) c.csv = 1;
    if( !out ) { fprintf( stderr, "x264 [error]: %s needs -o\n", argv[1] ); return -1; }
    if( extract && !( c.rate > 0 ) ) { fprintf( stderr, "x264 [error]: --extract-264 needs the embedder's --emrate\n" ); return -1; }
    data = read_file( in, &size );
    if( !data ) { fprintf( stderr, "x264 [error]: cannot read %s\n", in ); return -1; }
    if( extract ) { c.msg = fopen( out, "wb" ); if( stego ) c.stego = fopen( stego, "wb" ); }
    else c.mv = fopen( out, "wb" );
    if( ( extract ? !c.msg : !c.mv ) || ( stego && !c.stego ) ) { fprintf( stderr, "x264 [error]: cannot open the output\n" ); return -1; }
    n = pcamv_bitstream_walk( data, size, extract ? on_picture_extract : on_picture_mv, &c, err, sizeof(err) );
    free( data );
    if( c.mv ) fclose( c.mv );
    if( c.msg ) fclose( c.msg );
    if( c.stego ) fclose( c.stego );
    if( n < 0 ) { fprintf( stderr, "x264 [error]: %s: %s\n", in, err ); return -1; }
    if( extract )
        fprintf( stderr, "x264 [info]: %d pictures, %d P pictures, %d carriers: extracted %d payload bits from %d frames (%d skipped)\n",
                 n, c.p_pictures, c.carriers, c.bits, c.frames, c.skipped );
    else
        fprintf( stderr, "x264 [info]: %d pictures, %d P pictures parsed\n", n, c.p_pictures );
    return 0;
}
