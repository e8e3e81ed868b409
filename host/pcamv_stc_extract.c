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
 * draws (shorter, then longer), same effect on h->info.stego - including none at all when the message does not fit. */
#include "pcamv.h"
pcamv_ctx *pcamv_glue_ctx( void );
void pcamv_glue_stc_embed( x264_t *h, int an )
{
    const int n = h->info.length;
    uint32_t *cols[2];
    double invalpha;
    int shorter, longer, rc;
    if( an < 1 || n < an )
    {
        if( an >= 1 ) fprintf( stderr, "The message cannot be longer than the cover object.\n" );
        return;                     /* stc_embed gives up before touching stego (embed.h:349-356; an == 0: no matrix can be drawn) */
    }
    invalpha = (double)n / an;
    shorter = (int)floor( invalpha );
    longer = (int)ceil( invalpha );
    if( !( cols[0] = getMatrix( shorter, 10 ) ) ) return;
    if( !( cols[1] = getMatrix( longer, 10 ) ) ) { free( cols[0] ); return; }
    rc = pcamv_stc_embed( pcamv_glue_ctx(), h->info.cover, n, h->info.message, an, h->info.rho_final, h->info.stego, 10,
                          cols[0], shorter, cols[1], longer );
    free( cols[0] ); free( cols[1] );
    if( rc < 0 )
    {
        fprintf( stderr, "x264 [pcamv]: pcamv_stc_embed: %s\n", pcamv_last_error( pcamv_glue_ctx() ) );
        exit( 3 );                  /* no CPU fallback */
    }
    if( rc == 1 )
        fprintf( stderr, "The syndrome is not in the range of the syndrome matrix.\n" );
}
