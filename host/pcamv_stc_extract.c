/* pcamv_stc_extract.c - the extraction side of the PCAMV payload channel.
 *
 * The reference has an embedder (stc_embed, embed.h:309-548) but no extractor: its include of "stc_extract_c.h" is
 * commented out (encoder/analyse.c:43).  This file supplies it.  It is appended to the scratch copy of
 * encoder/encoder.c by host/build_host.py, because the parity-check sub-matrices come from the reference's own
 * `getMatrix` (embed.h:276-306), a `static` function of that translation unit (built-in tables for widths 2..20, an LCG
 * with persistent state beyond) - the extractor must draw the very same matrices, in the same order, as the embedder.
 *
 * Algorithm: syndrome of the stego vector under the banded parity-check matrix the embedder's trellis walks
 * (Filler, Judas, Fridrich, "Minimizing additive distortion in steganography using syndrome-trellis codes", 2011):
 * cover element `index` of block `b` uses column k of the sub-matrix chosen for that block; bit l of the column feeds
 * message bit b + l.  Block widths follow the embedder's schedule (embed.h:376-392: floor / ceil of n / an, the wider one
 * whenever `worm + longer <= (b+1) * n/an + 0.5`), and bits past the end of the message are dropped, which is what the
 * embedder's shrinking `colmask` does (embed.h:482-483).
 *
 * Call order contract: one call per embedded frame, in coding order, from a fresh process - the same sequence of
 * getMatrix(shorter), getMatrix(longer) calls the encoder made. */
int pcamv_stc_extract( const uint8_t *stego, int n, uint8_t *message, int an, int matrixheight )
{
    uint32_t *columns[2];
    double invalpha;
    int shorter, longer, worm = 0, index = 0, b, k, l;
    if( an <= 0 || n < an || matrixheight < 1 || matrixheight > 31 )
        return -1;
    invalpha = (double)n / an;
    shorter = (int)floor( invalpha );
    longer = (int)ceil( invalpha );
    if( !( columns[0] = getMatrix( shorter, matrixheight ) ) )
        return -1;
    if( !( columns[1] = getMatrix( longer, matrixheight ) ) )
    {
        free( columns[0] );
        return -1;
    }
    memset( message, 0, an );
    for( b = 0; b < an; b++ )
    {
        const int wide = worm + longer <= ( b + 1 ) * invalpha + 0.5;
        const int width = wide ? longer : shorter;
        const uint32_t *cols = columns[wide];
        worm += width;
        for( k = 0; k < width; k++, index++ )
            if( stego[index] )
                for( l = 0; l < matrixheight && b + l < an; l++ )
                    message[b + l] ^= ( cols[k] >> l ) & 1;
    }
    free( columns[0] );
    free( columns[1] );
    return index;           /* cover elements consumed (= n when the schedule covers the vector exactly) */
}

/* `x264_pcamv --stc-columns W H`: the sub-matrix getMatrix( W, H ) hands out (tests fetch the embedder's matrices this way) */
int pcamv_stc_columns( int width, int height, uint32_t *out )
{
    uint32_t *c = getMatrix( width, height );
    if( !c ) return -1;
    memcpy( out, c, width * sizeof(uint32_t) );
    free( c );
    return 0;
}

/* The embed stage's stc_embed call (encoder/encoder.c:1843) on the GPU: same argument checks, same order of getMatrix
 * draws (shorter, then longer), same effect on h->info.stego - including none at all when the message does not fit.
 *
 * Default: the whole stage is device-resident (pcamv_embed_prepare / pcamv_embed_stc, csrc/pcamv_embed.cu).  The GPU has built
 * cover and rho_final itself from the pass-1 results it still holds, runs the trellis on those, derives the flips and the
 * decisions pass 2 forces, and keeps them in HBM; this function only hands it the message and brings the stego bits back for
 * the host's own bookkeeping (the host's loops around this call are unchanged, so h->info.* evolves as in the reference).
 * The host's cover length must equal the GPU's (fatal otherwise); PCAMV_CHECK_EMBED=1 also compares cover and rho_final bit
 * for bit.  PCAMV_HOST_EMBED=1 keeps the round-1 flow (host vectors up, stego down, pass-1 records + flips up for pass 2). */
#include "pcamv.h"
pcamv_ctx *pcamv_glue_ctx( void );
void pcamv_glue_set_device_forced( int on );
void pcamv_glue_set_embed_status( int embedded );
static int glue_stc_block_total( const float *rho, int n, int an, double *total )
{
    /* the reference's `total`: rho of the elements its block schedule covers, in order, in double (embed.h:376-392,405-470) */
    const double invalpha = (double)n / an;
    const int shorter = (int)floor( invalpha ), longer = (int)ceil( invalpha );
    int worm = 0, used = 0, b, i;
    for( b = 0; b < an; b++ )
    {
        const int wide = worm + longer <= ( b + 1 ) * invalpha + 0.5;
        worm += wide ? longer : shorter;
        used += wide ? longer : shorter;
    }
    if( used > n ) used = n;
    *total = 0;
    for( i = 0; i < used; i++ ) *total += rho[i];
    return used;
}
static void glue_stc_embed( x264_t *h, int an, int prepared )
{
    const int n = h->info.length;
    pcamv_ctx *ctx = pcamv_glue_ctx();
    const char *s;
    const int on_device = !( ( s = getenv( "PCAMV_HOST_EMBED" ) ) && atoi( s ) );
    uint32_t *cols[2] = { NULL, NULL };
    double invalpha, total = 0;
    int shorter = 0, longer = 0, rc, embed = 1;
    if( on_device && !prepared )
    {
        int dev_n = -1;
        if( pcamv_embed_prepare( ctx, &dev_n ) )
        {
            fprintf( stderr, "x264 [pcamv]: pcamv_embed_prepare: %s\n", pcamv_last_error( ctx ) );
            fflush( stderr ); _exit( 3 );
        }
        if( dev_n != n )
        {
            fprintf( stderr, "x264 [pcamv]: frame %d: the host counts %d carriers, the GPU %d\n", h->i_frame, n, dev_n );
            fflush( stderr ); _exit( 4 );
        }
        if( ( s = getenv( "PCAMV_CHECK_EMBED" ) ) && atoi( s ) && n > 0 )
        {
            uint8_t *c = malloc( n ); float *r = malloc( n * sizeof(float) );
            if( pcamv_embed_download( ctx, c, r, NULL, NULL, NULL ) || memcmp( c, h->info.cover, n ) || memcmp( r, h->info.rho_final, n * sizeof(float) ) )
            {
                fprintf( stderr, "x264 [pcamv]: frame %d: cover / rho_final built on the GPU differ from the host's\n", h->i_frame );
                fflush( stderr ); _exit( 4 );
            }
            free( c ); free( r );
        }
    }
    if( an < 1 || n < an )
    {
        if( an >= 1 ) fprintf( stderr, "The message cannot be longer than the cover object.\n" );
        embed = 0;                  /* stc_embed gives up before touching stego (embed.h:349-356; an == 0: no matrix can be drawn) */
    }
    if( embed )
    {
        invalpha = (double)n / an;
        shorter = (int)floor( invalpha );
        longer = (int)ceil( invalpha );
        if( !( cols[0] = getMatrix( shorter, 10 ) ) ) embed = 0;
        else if( !( cols[1] = getMatrix( longer, 10 ) ) ) { free( cols[0] ); cols[0] = NULL; embed = 0; }
    }
    if( on_device )
    {
        if( embed ) glue_stc_block_total( h->info.rho_final, n, an, &total );
        rc = embed ? pcamv_embed_stc( ctx, h->info.message, an, 10, cols[0], shorter, cols[1], longer, total, h->info.stego )
                   : pcamv_embed_stc( ctx, NULL, 0, 10, NULL, 0, NULL, 0, -1.0, NULL );
        pcamv_glue_set_device_forced( rc >= 0 );
    }
    else
    {
        if( !embed ) { pcamv_glue_set_embed_status( 0 ); return; }
        rc = pcamv_stc_embed( ctx, h->info.cover, n, h->info.message, an, h->info.rho_final, h->info.stego, 10,
                              cols[0], shorter, cols[1], longer );
    }
    free( cols[0] ); free( cols[1] );
    pcamv_glue_set_embed_status( embed && rc == 0 );
    if( rc < 0 )
    {
        fprintf( stderr, "x264 [pcamv]: stc embed: %s\n", pcamv_last_error( ctx ) );
        fflush( stderr ); _exit( 3 );                  /* no CPU fallback */
    }
    if( rc == 1 )
        fprintf( stderr, "The syndrome is not in the range of the syndrome matrix.\n" );
}
void pcamv_glue_stc_embed( x264_t *h, int an ) { glue_stc_embed( h, an, 0 ); }
/* pass 1 without the host (pcamv_hook_pass1_on_device): cover and rho_final in h->info are the device's own, already assembled */
void pcamv_glue_stc_embed_prepared( x264_t *h, int an ) { glue_stc_embed( h, an, 1 ); }
