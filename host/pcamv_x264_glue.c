/* pcamv_x264_glue.c — the reference's C host bound to libpcamv_cuda.so (include/pcamv.h).
 *
 * This is the binding INTEGRATION.md describes, compiled into the reference encoder by host/build_host.py (which
 * inserts the hook calls at anchored lines of a scratch copy of the reference; no reference source lives in this
 * repository).  The encoder keeps its x264_encoder_encode / --emrate interface and every line of its own control
 * flow; what changes is where the motion-estimation results come from:
 *
 *   x264_encoder_open   (encoder/encoder.c:766)   -> pcamv_open
 *   x264_slice_write    (encoder/encoder.c:1176)  -> per P-slice pass: pcamv_put_fenc, pcamv_put_ref for references not
 *                                                    yet on the GPU, pcamv_set_qp_tables, ONE pcamv_analyse_p call
 *   x264_me_search_ref  (encoder/me.c:158)        -> pops the next entry of the macroblock's result log
 *   x264_me_refine_qpel (encoder/me.c:669)        -> same
 *   x264_ih_get_mv_cost (encoder/analyse.c:2391)  -> same (replacement delta + embedding cost)
 *   x264_encoder_close  (encoder/encoder.c:2670)  -> pcamv_close
 *
 * The host still walks x264_macroblock_analyse for every macroblock (MV prediction, candidate lists, P_SKIP probe,
 * mode decision, pass-2 forcing, the unsequenced MV copy of SURVEY.md fact 3 ...), so its state evolves exactly as
 * in the reference; each replayed call is checked against the log entry's kind / block size / reference, and any
 * disagreement between the host's call sequence and the GPU's is fatal.  There is no CPU fallback: if the library
 * cannot be opened or a call fails, the encoder exits.
 */
#include "common/common.h"
#include "encoder/me.h"
#include "encoder/macroblock.h"
#include "pcamv.h"
#include <time.h>
#include <pthread.h>
#include <unistd.h>

void pcamv_glue_load_costs( x264_t *h, int qp );      /* added to encoder/analyse.c by host/build_host.py */

extern int16_t *g_cost_mv[52];            /* reference encoder/analyse.c:189 (malloc base, centre at +2*4*2048) */
extern uint16_t x264_cost_ref[52][3][33]; /* reference encoder/analyse.c:188 */
extern const int x264_lambda_tab[52];
extern const int x264_lambda2_tab[52];

typedef struct
{
    x264_t *h;
    pcamv_ctx *ctx;
    int n_mb, log_stride;
    pcamv_mb_out *mbs;              /* results of the running slice pass */
    pcamv_log_entry *log;
    pcamv_pass1_mb *pass1;          /* staging of h->info.cache[] for pass 2 */
    int active;                     /* a P-slice pass with GPU results is being replayed */
    int cur_mb, cur_pos;            /* replay cursor */
    int qp_loaded;
    int elide, pass, pinned;                /* pass-2 elision on; pass of the slice being replayed */
    int host_hpel;                  /* PCAMV_HOST_HPEL=1: the host filters its own half-pel planes even for frames the GPU rebuilds */
    int skip_hpel_frame;            /* i_frame of the frame whose half-pel planes come back from the GPU instead of x264_frame_filter (-1: none) */
    int host_intra;                 /* PCAMV_HOST_INTRA=1: intra analysis of every P macroblock, as the reference does for its statistics */
    int stream_rows, rows_ready, mb_h;      /* the replayed pass is consumed row by row while the kernel runs (PCAMV_NO_ROW_STREAM=1: wait for its end) */
    double t_row_wait, t_pass1, t_embed;
    int direct;                     /* pass 1 of an embedding frame never reaches the host's macroblock loop (pcamv_hook_pass1_on_device) */
    int16_t last_mv[16][2]; int have_last_mv; long stale_mismatch;
    int device_forced;              /* the embed stage of this frame ran on the device: pass 2 takes its forced decisions from HBM */
    /* reference frames built on the device (default; PCAMV_DEVICE_RECON=0 uploads them instead): after the final pass of a P frame the GPU reconstructs and
     * deblocks the frame into a free slot, and the next frame finds it resident instead of uploading it */
    int recon_on, recon_check, cur_ref_slots[PCAMV_MAX_REFS], cur_n_ref;
    pcamv_recon_patch *patches; int n_patches, cap_patches;
    long recon_frames, recon_mismatch, recon_patched_mbs, hpel_frames;
    int fenc_frame;                 /* h->fenc->i_frame currently on the GPU */
    /* reference slots: which frame each GPU slot holds */
    struct { x264_frame_t *fr; int i_frame, i_poc, age, device_built; } slot[PCAMV_MAX_REFS + 2];
    int n_slots, tick;
    /* accounting */
    double t_gpu, t_total0, t_open, t_ref_upload, t_recon;
    long n_passes, n_replayed, n_direct;
} glue_t;

/* one encoder instance lives on one thread from x264_encoder_open to x264_encoder_close (x264.c Encode), so the glue
 * state is thread-local: several encoders (GOP shards) run side by side in one process (--shards, see host/build_host.py) */
static __thread glue_t g;

/* shards of one process share the GPU through an encoder group (include/pcamv.h): one multi-context launch per step */
/* PCAMV_GROUPS=G (default 2) splits the shards into G groups that rendezvous independently (shard i joins group i % G):
 * while one group's frames are on the GPU the other groups' encoders do their host work, so neither side idles */
#define PCAMV_MAX_GROUPS 64
static pcamv_group *g_groups[PCAMV_MAX_GROUPS];
static int g_n_groups;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static __thread pcamv_group *g_group;       /* this encoder thread's group, NULL outside --shards */
static __thread int g_in_group;
static __thread int g_shard = -1;           /* this encoder thread's shard index, -1 outside --shards */

void pcamv_glue_set_shards( int n )
{
    const char *s = getenv( "PCAMV_GROUPS" );
    int i, ng = s ? atoi( s ) : 2;
    if( n <= 1 )
        return;
    if( ng < 1 ) ng = 1;
    if( ng > PCAMV_MAX_GROUPS ) ng = PCAMV_MAX_GROUPS;
    if( ng > n ) ng = n;
    for( i = 0; i < ng; i++ )
        if( pcamv_group_create( &g_groups[i], ( n - i + ng - 1 ) / ng ) )     /* members with index = i mod ng */
        {
            fprintf( stderr, "x264 [pcamv]: pcamv_group_create failed\n" );
            exit( 3 );
        }
    g_n_groups = ng;
}

/* this encoder thread's context (the embed-stage hook in the encoder.c translation unit needs it) */
pcamv_ctx *pcamv_glue_ctx( void ) { return g.ctx; }
void pcamv_glue_set_device_forced( int on ) { g.device_forced = on; }
/* did the trellis of the frame being encoded actually embed its message?  (no: fewer carriers than the message needs, or
 * "not in the range of the syndrome matrix" - the reference carries on with stego = 0 and flips every carrier whose cover
 * bit is 1, encoder/encoder.c:1826,1843-1855; nothing can be extracted from such a frame) */
static __thread int g_embedded;
void pcamv_glue_set_embed_status( int embedded ) { g_embedded = embedded; }

/* called first thing by a shard's thread: i = its index in this process (decides the group), gop = the GOP it encodes
 * (names its side files).  Group membership starts HERE, not at x264_encoder_open: a shard whose Encode() fails before it
 * ever opens an encoder must still leave its group, or its siblings would wait for it forever. */
void pcamv_glue_set_shard_index( int i, int gop )
{
    g_group = g_n_groups ? g_groups[i % g_n_groups] : NULL;
    g_in_group = g_group != NULL;
    g_shard = gop;
}

/* side files (PCAMV_PAYLOAD, PCAMV_STEGO): one per shard ("<name>.<shard>") when several encoders share the process */
static FILE *open_side_file( const char *name )
{
    char path[1200];
    if( g_shard >= 0 ) snprintf( path, sizeof(path), "%s.%d", name, g_shard );
    else snprintf( path, sizeof(path), "%s", name );
    return fopen( path, "ab" );
}

/* called by a shard's thread when its encoder is gone, however it ended */
void pcamv_glue_shard_done( void )
{
    if( g_group && g_in_group )
    {
        pcamv_group_leave( g_group );
        g_in_group = 0;
    }
}

/* The payload is rand() & 1 with no srand anywhere (encoder/encoder.c:1838-1840): the seed-1 stream of glibc rand().  A
 * shard is an independent encoder run, so it needs that stream from its start — per thread, not shared.  random_r with
 * initstate_r( 1, 128 bytes ) is the generator behind rand() (checked value for value in tests). */
int pcamv_tls_rand( void )
{
    static __thread struct random_data rd;
    static __thread char st[128];
    static __thread int init;
    int32_t r;
    if( !init )
    {
        rd.state = NULL;
        initstate_r( 1, st, sizeof(st), &rd );
        init = 1;
    }
    random_r( &rd, &r );
    return r;
}

static double now_s( void )
{
    struct timespec ts; clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + 1e-9*ts.tv_nsec;
}

/* fatal errors end the process at once (_exit): with several encoder threads in one process, exit() would run the
 * teardown of the CUDA runtime while sibling shards sit inside CUDA calls or wait on their group, which can hang */
static void die( const char *what )
{
    fprintf( stderr, "x264 [pcamv]: %s: %s\n", what, g.ctx ? pcamv_last_error( g.ctx ) : pcamv_last_error( NULL ) );
    fflush( stderr );
    _exit( 3 );
}
static void die_msg( const char *msg )
{
    fprintf( stderr, "x264 [pcamv]: %s\n", msg );
    fflush( stderr );
    _exit( 3 );
}

/* ---- open / close ---------------------------------------------------------------------------- */
/* PCAMV_CONFORMANT=1 (off by default: it changes the bitstream): the three statements of the reference's pass 2 that make its
 * embedding streams unreadable for a standard decoder are corrected - on the host by tools/reftree.py::conformance_switch (the
 * edits call this function), on the device by pcamv_set_conformant.  The parity reference of this mode is
 * oracle/_ref/x264_dump_conformant; what it buys is extraction from the .264 alone (host/pcamv_bitstream.c). */
int pcamv_conformant( void )
{
    static int on = -1;
    if( on < 0 ) { const char *s = getenv( "PCAMV_CONFORMANT" ); on = s && atoi( s ); }
    return on;
}

void pcamv_hook_open( x264_t *h )
{
    pcamv_cfg cfg;
    const char *s;
    memset( &g, 0, sizeof(g) );
    g.h = h;
    g.t_total0 = now_s();
    g.fenc_frame = -1;
    g.skip_hpel_frame = -1;
    if( h->param.i_threads > 1 )
        die_msg( "frame threads are not supported (the reference itself crashes with embedding on)" );
    if( h->param.rc.i_rc_method != X264_RC_CQP || h->param.rc.i_aq_mode )
        die_msg( "the GPU path needs constant QP (--qp N): one QP per slice" );
    if( h->param.i_bframe )
        die_msg( "B frames are not supported" );
    if( h->param.analyse.i_subpel_refine > 5 )
        die_msg( "--subme 6 and above (RD mode decision on live CABAC state) is raster-serial; use --subme 1..5" );
    if( h->param.analyse.b_mixed_references )
        die_msg( "--mixed-refs is not supported" );
    memset( &cfg, 0, sizeof(cfg) );
    cfg.abi_version = PCAMV_ABI_VERSION;
    cfg.device = (s = getenv( "PCAMV_DEVICE" )) ? atoi( s ) : 0;
    cfg.width = 16 * h->sps->i_mb_width;
    cfg.height = 16 * h->sps->i_mb_height;
    cfg.me_method = h->param.analyse.i_me_method;
    cfg.me_range = h->param.analyse.i_me_range;
    cfg.subpel_refine = h->param.analyse.i_subpel_refine;
    cfg.chroma_me = h->param.analyse.b_chroma_me;
    cfg.max_refs = h->param.i_frame_reference;
    cfg.mv_range = h->param.analyse.i_mv_range;
    cfg.b_cabac = h->param.b_cabac;
    cfg.b_fast_pskip = h->param.analyse.b_fast_pskip;
    cfg.b_dct_decimate = h->param.analyse.b_dct_decimate;
    cfg.analyse_inter = h->param.analyse.inter;
    cfg.chroma_qp_offset = h->param.analyse.i_chroma_qp_offset;
    cfg.no_deblock = !h->param.b_deblocking_filter;
    cfg.deblock_alpha_c0_offset = h->param.i_deblocking_filter_alphac0 << 1;    /* sh.i_alpha_c0_offset, encoder/encoder.c:170-171 */
    cfg.deblock_beta_offset = h->param.i_deblocking_filter_beta << 1;
    cfg.rows_per_cta = (s = getenv( "PCAMV_ROWS_PER_CTA" )) ? atoi( s ) : 1;
    /* pass 2 of a forced macroblock: only the 16x16 search is live in the reference (see include/pcamv.h); PCAMV_PASS2_FULL=1
     * makes the GPU execute and log the dead searches as well */
    g.elide = !((s = getenv( "PCAMV_PASS2_FULL" )) && atoi( s ));
    /* with sub-8x8 partitions a forced P_8x8 macroblock keeps the h->mb.i_partition its own pass-2 analysis decided
     * (analyse.c:2872-2890), and the MVD prediction of the bitstream writer reads it: those searches are not dead */
    if( h->param.analyse.inter & X264_ANALYSE_PSUB8x8 )
        g.elide = 0;
    cfg.pass2_elide = g.elide;
    pthread_mutex_lock( &g_mu );
    if( pcamv_open( &g.ctx, &cfg ) )
        die( "pcamv_open" );
    if( pcamv_conformant() && pcamv_set_conformant( g.ctx, 1 ) )
        die( "pcamv_set_conformant" );
    /* the reference builds its lambda*bits tables lazily into function-static storage (analyse.c:193-229): do it here, once,
     * under the lock, so that concurrent encoders only ever read them */
    pcamv_glue_load_costs( h, h->param.rc.i_qp_constant );
    pthread_mutex_unlock( &g_mu );
    g.n_mb = h->sps->i_mb_width * h->sps->i_mb_height;
    g.n_slots = cfg.max_refs + 2;
    g.pinned = !((s = getenv( "PCAMV_NO_PINNED" )) && atoi( s ));
    g.mbs = g.pinned ? pcamv_host_alloc( g.n_mb * sizeof(*g.mbs) ) : calloc( g.n_mb, sizeof(*g.mbs) );      /* page-locked: results arrive without a staging copy */
    g.log_stride = pcamv_log_stride( g.ctx );
    g.log = g.pinned ? pcamv_host_alloc( (size_t)g.n_mb * g.log_stride * sizeof(*g.log) ) : calloc( (size_t)g.n_mb * g.log_stride, sizeof(*g.log) );
    g.pass1 = calloc( g.n_mb, sizeof(*g.pass1) );
    if( !g.mbs || !g.log || !g.pass1 )
        die_msg( "out of memory" );
    g.qp_loaded = -1;
    /* pass 1 on the device alone unless the host's own pass 1 is asked for — or needed, because something is to be checked against it */
    g.direct = !( ( (s = getenv( "PCAMV_HOST_PASS1" )) && atoi( s ) ) || ( (s = getenv( "PCAMV_HOST_EMBED" )) && atoi( s ) ) ||
                  ( (s = getenv( "PCAMV_CHECK_EMBED" )) && atoi( s ) ) );
    g.host_intra = (s = getenv( "PCAMV_HOST_INTRA" )) && atoi( s );
    g.host_hpel = (s = getenv( "PCAMV_HOST_HPEL" )) && atoi( s );
    g.mb_h = h->sps->i_mb_height;
    g.stream_rows = g.pinned && !( (s = getenv( "PCAMV_NO_ROW_STREAM" )) && atoi( s ) );
    g.recon_on = !( (s = getenv( "PCAMV_DEVICE_RECON" )) && !atoi( s ) );       /* default on; PCAMV_DEVICE_RECON=0: every reference is uploaded */
    g.recon_check = (s = getenv( "PCAMV_CHECK_RECON" )) && atoi( s );
    g.t_open = now_s() - g.t_total0;
}

void pcamv_hook_close( x264_t *h )
{
    const char *s = getenv( "PCAMV_STATS" );
    (void)h;
    if( s && *s )
    {
        char path[1200];
        FILE *f;
        if( g_shard >= 0 ) snprintf( path, sizeof(path), "%s.%d", s, g_shard );     /* one file per shard, like the side files */
        else snprintf( path, sizeof(path), "%s", s );
        f = fopen( path, "w" );
        if( f )
        {
            fprintf( f, "{\"p_passes\": %ld, \"replayed_calls\": %ld, \"gpu_launches\": %lld, \"t_gpu_calls\": %.6f, \"t_open\": %.6f, \"t_total\": %.6f, "
                        "\"recon_frames\": %ld, \"recon_mismatch\": %ld, \"recon_patched_mbs\": %ld, \"t_ref_upload\": %.6f, \"t_recon\": %.6f, \"direct_pass1\": %ld, \"stale_mismatch\": %ld, \"t_row_wait\": %.6f, \"t_pass1\": %.6f, \"t_embed\": %.6f, \"hpel_frames\": %ld}\n",
                     g.n_passes, g.n_replayed, g.ctx ? pcamv_launch_count( g.ctx ) : 0LL, g.t_gpu, g.t_open, now_s() - g.t_total0,
                     g.recon_frames, g.recon_mismatch, g.recon_patched_mbs, g.t_ref_upload, g.t_recon, g.n_direct, g.stale_mismatch, g.t_row_wait, g.t_pass1, g.t_embed, g.hpel_frames );
            fclose( f );
        }
    }
    pcamv_glue_shard_done();
    if( getenv( "PCAMV_VERBOSE" ) )
        fprintf( stderr, "x264 [pcamv]: %ld P passes, open %.3f s, GPU calls %.3f s (waiting for the group included), encoder lifetime %.3f s\n",
                 g.n_passes, g.t_open, g.t_gpu, now_s() - g.t_total0 );
    if( g.ctx ) pcamv_close( g.ctx );
    if( g.pinned ) { pcamv_host_free( g.mbs ); pcamv_host_free( g.log ); } else { free( g.mbs ); free( g.log ); }
    free( g.pass1 ); free( g.patches );
    memset( &g, 0, sizeof(g) );
}

/* ---- slice begin: upload what is new, analyse the whole P slice on the GPU ---------------------- */
/* returns 0 when the GPU's planes of slot `s` equal the host's planes of `fr`, else 1 + the first differing plane */
static int check_device_ref( int s, x264_frame_t *fr )
{
    /* picture + borders, row by row: the luma strides of the two sides are equal by construction, the chroma strides are not
     * for widths that are 16 mod 32 (the host halves its luma stride, common/frame.c:52-60; the GPU aligns to 16), and the
     * alignment slack behind a row's right border belongs to nobody */
    int k, bad = 0;
    for( k = 0; k < 6 && !bad; k++ )
    {
        const size_t n = pcamv_plane_bytes( g.ctx, k );
        const int stride = pcamv_plane_stride( g.ctx, k ), padv = k < 4 ? 32 : 16, padh = k < 4 ? 32 : 16;
        const int hstride = fr->i_stride[k < 4 ? 0 : 1];
        const int width = ( 16 * g.h->sps->i_mb_width >> ( k >= 4 ) ) + 2 * padh, rows = (int)( n / stride );
        uint8_t *dev = malloc( n );
        const uint8_t *host = ( k < 4 ? fr->filtered[k] : fr->plane[k - 3] ) - (size_t)padv * hstride - padh;
        size_t count = 0;
        int y, x, fy = -1, fx = -1;
        if( !dev || pcamv_get_ref_plane( g.ctx, s, k, dev ) )
            die( "pcamv_get_ref_plane" );
        for( y = 0; y < rows; y++ )
            if( memcmp( dev + (size_t)y * stride, host + (size_t)y * hstride, width ) )
                for( x = 0; x < width; x++ )
                    if( dev[(size_t)y * stride + x] != host[(size_t)y * hstride + x] )
                    {
                        if( !count ) { fy = y; fx = x; }
                        count++;
                    }
        if( count )
        {
            bad = k + 1;
            fprintf( stderr, "x264 [pcamv]: plane %d: %zu bytes differ, the first at line %d column %d of the picture (GPU %d, host %d)\n", k, count,
                     fy - padv, fx - padh, dev[(size_t)fy * stride + fx], host[(size_t)fy * hstride + fx] );
        }
        free( dev );
    }
    return bad;
}

static int slot_of( x264_frame_t *fr, const int *in_use, int n_in_use )
{
    int i, k, best = -1;
    for( i = 0; i < g.n_slots; i++ )
        if( g.slot[i].fr == fr && g.slot[i].i_frame == fr->i_frame && g.slot[i].i_poc == fr->i_poc )
        {
            g.slot[i].age = ++g.tick;
            if( g.slot[i].device_built == 1 && g.recon_check )
            {
                /* check mode: the planes the GPU built for this picture against the host's own (integer + the three half-pel
                 * planes + chroma, whole padded buffers: same layout on both sides); on a mismatch the host's are uploaded */
                int bad = check_device_ref( i, fr );
                g.slot[i].device_built = 2;
                if( bad )
                {
                    g.recon_mismatch++;
                    fprintf( stderr, "x264 [pcamv]: frame %d: reference planes built on the GPU differ from the host's (plane %d)\n", fr->i_frame, bad - 1 );
                    if( pcamv_put_ref( g.ctx, i, fr->i_poc, fr->plane[0], fr->plane[1], fr->plane[2], fr->i_stride[0], fr->i_stride[1] ) )
                        die( "pcamv_put_ref" );
                }
            }
            return i;
        }
    /* not resident: take the least recently used slot that no reference of this slice occupies */
    for( i = 0; i < g.n_slots; i++ )
    {
        int busy = 0;
        for( k = 0; k < n_in_use; k++ ) busy |= in_use[k] == i;
        if( !busy && ( best < 0 || g.slot[i].age < g.slot[best].age ) )
            best = i;
    }
    if( best < 0 )
        die_msg( "no free reference slot" );
    /* the integer plane as the host holds it after deblocking (filtered[0] == plane[0]); the GPU rebuilds borders,
     * half-pel planes and integral image itself and must arrive at the host's own planes bit for bit */
    {
        double t0 = now_s();
        if( pcamv_put_ref( g.ctx, best, fr->i_poc, fr->plane[0], fr->plane[1], fr->plane[2], fr->i_stride[0], fr->i_stride[1] ) )
            die( "pcamv_put_ref" );
        g.t_ref_upload += now_s() - t0;
    }
    g.slot[best].fr = fr; g.slot[best].i_frame = fr->i_frame; g.slot[best].i_poc = fr->i_poc; g.slot[best].age = ++g.tick;
    g.slot[best].device_built = 0;
    return best;
}

void pcamv_hook_slice_begin( x264_t *h )
{
    pcamv_frame_in in;
    int qp = h->sh.i_qp, pass, i, n;
    double t0;
    g.active = 0;
    if( h->sh.i_type != SLICE_TYPE_P )
        return;
    t0 = now_s();
    pass = !h->info.embed_flag ? 0 : h->info.firstTime ? 1 : 2;
    /* a frame the GPU rebuilds as a reference (PCAMV_DEVICE_RECON) gets its half-pel planes from there too: the host skips
     * x264_frame_filter for it (not in check mode, which compares the host's own planes with the GPU's) */
    g.skip_hpel_frame = ( g.recon_on && !g.recon_check && !g.host_hpel && h->fdec->b_kept_as_ref ) ? h->fdec->i_frame : -1;
    if( qp != g.qp_loaded )
    {
        pcamv_qp_tables t;
        int qpc = h->chroma_qp_table[qp];
        pthread_mutex_lock( &g_mu );
        pcamv_glue_load_costs( h, qp );       /* the reference builds the table lazily at the first macroblock (analyse.c:198) */
        pthread_mutex_unlock( &g_mu );
        memset( &t, 0, sizeof(t) );
        t.qp = qp; t.lambda = x264_lambda_tab[qp]; t.lambda2_chroma = x264_lambda2_tab[qpc]; t.chroma_qp = qpc;
        t.cost_mv = g_cost_mv[qp];
        t.cost_ref = &x264_cost_ref[qp][0][0];
        t.quant4_mf[0] = (const uint16_t *)h->quant4_mf[CQM_4PY][qp];   t.quant4_bias[0] = (const uint16_t *)h->quant4_bias[CQM_4PY][qp];
        t.quant4_mf[1] = (const uint16_t *)h->quant4_mf[CQM_4PC][qpc];  t.quant4_bias[1] = (const uint16_t *)h->quant4_bias[CQM_4PC][qpc];
        t.dequant4_mf[0] = (const int32_t *)h->dequant4_mf[CQM_4PY];
        t.dequant4_mf[1] = (const int32_t *)h->dequant4_mf[CQM_4PC];
        if( pcamv_set_qp_tables( g.ctx, &t ) )
            die( "pcamv_set_qp_tables" );
        g.qp_loaded = qp;
    }
    if( g.fenc_frame != h->fenc->i_frame )
    {
        if( pcamv_put_fenc( g.ctx, h->fenc->plane[0], h->fenc->plane[1], h->fenc->plane[2], h->fenc->i_stride[0], h->fenc->i_stride[1] ) )
            die( "pcamv_put_fenc" );
        g.fenc_frame = h->fenc->i_frame;
    }
    memset( &in, 0, sizeof(in) );
    in.pass = pass;
    in.n_ref = h->i_ref0;
    for( i = 0; i < h->i_ref0; i++ )
    {
        in.ref_slot[i] = slot_of( h->fref0[i], in.ref_slot, i );
        in.ref_poc[i] = h->fref0[i]->i_poc;
    }
    in.cur_poc = h->fdec->i_poc;
    g.cur_n_ref = h->i_ref0;
    for( i = 0; i < h->i_ref0; i++ ) g.cur_ref_slots[i] = in.ref_slot[i];
    g.n_patches = 0;
    {
        x264_frame_t *l0 = h->fref0[0];
        in.col_n_ref = l0->i_ref[0];
        for( i = 0; i < 16; i++ ) in.col_inv_ref_poc[i] = l0->inv_ref_poc[i];
        in.col_ref8 = l0->ref[0];
        in.col_mv4 = (const int16_t *)l0->mv[0];
    }
    in.cost_table = pass == 1;
    if( pass == 1 )
        g.device_forced = 0;
    if( pass == 2 && g.device_forced )
        in.device_forced = 1;       /* cover, flips and forced decisions of this frame are already in HBM (pcamv_glue_stc_embed) */
    else if( pass == 2 )
    {
        n = 0;
        for( i = 0; i < g.n_mb; i++ )
        {
            pcamv_pass1_mb *p = &g.pass1[i];
            p->type = h->info.cache[i].i_type;
            p->partition = h->info.cache[i].i_partition;
            p->used = h->info.cache[i].used;
            memcpy( p->sub, h->info.cache[i].i_sub_partition, 4 );
            memcpy( p->ref, h->info.cache[i].ref, 16 );
            memcpy( p->mv, h->info.cache[i].mv, 64 );
            memcpy( p->mv_stego, h->info.cache[i].mv_stego, 64 );
            if( p->used && p->type == P_8x8 )
            {
                int k;      /* carriers of a split 8x8 block: 4x4 -> 4, 8x4 / 4x8 -> 2, 8x8 -> 1 (encoder/encoder.c:1571-1620) */
                for( k = 0; k < 4; k++ )
                    n += p->sub[k] == D_L0_4x4 ? 4 : p->sub[k] == D_L0_8x8 ? 1 : 2;
            }
            else if( p->used )
                n += p->partition == D_16x16 ? 1 : 2;
        }
        in.pass1 = g.pass1;
        in.filp = h->info.filp;
        in.n_filp = n;
    }
    if( pass == 2 && g.have_last_mv )
    {
        /* the MV cache between the passes is what pass 1 left for its last macroblock: with pass 1 on the device it is put
         * there from the GPU's record, with the host's own pass 1 the two are compared (stale_mismatch in PCAMV_STATS) */
        for( i = 0; i < 16; i++ )
            if( g.direct )
            {
                h->mb.cache.mv[0][x264_scan8[i]][0] = g.last_mv[i][0];
                h->mb.cache.mv[0][x264_scan8[i]][1] = g.last_mv[i][1];
            }
            else if( h->mb.cache.mv[0][x264_scan8[i]][0] != g.last_mv[i][0] || h->mb.cache.mv[0][x264_scan8[i]][1] != g.last_mv[i][1] )
            {
                g.stale_mismatch++;
                break;
            }
    }
    g.have_last_mv = 0;
    for( i = 0; i < 16; i++ )
    {
        /* what the MV cache holds before macroblock 0 (the pass-2 "forced skip without cache update" quirk reads it) */
        in.stale_mv[i][0] = h->mb.cache.mv[0][x264_scan8[i]][0];
        in.stale_mv[i][1] = h->mb.cache.mv[0][x264_scan8[i]][1];
    }
    /* the pass the host replays (0 or 2): return as soon as the launch is in flight and follow the wavefront row by row
     * (pcamv_hook_analyse_begin); pass 1 is consumed whole, by the embed stage */
    g.rows_ready = g.mb_h;
    if( g.stream_rows && pass != 1 )
    {
        if( g_group ? pcamv_group_analyse_p_begin( g_group, g.ctx, &in, g.mbs, g.log ) : pcamv_analyse_p_begin( g.ctx, &in, g.mbs, g.log ) )
            die( "pcamv_analyse_p_begin" );
        g.rows_ready = 0;
    }
    else if( g_group ? pcamv_group_analyse_p( g_group, g.ctx, &in, g.mbs, g.log ) : pcamv_analyse_p( g.ctx, &in, g.mbs, g.log ) )
        die( "pcamv_analyse_p" );
    g.active = 1;
    g.pass = pass;
    g.cur_mb = -1; g.cur_pos = 0;
    g.n_passes++;
    if( pass == 1 )
    {
        memcpy( g.last_mv, g.mbs[g.n_mb - 1].mv, sizeof(g.last_mv) );
        g.have_last_mv = 1;
        g.t_pass1 += now_s() - t0;
    }
    g.t_gpu += now_s() - t0;
}

/* x264_fdec_filter_row: does the frame being coded get its half-pel planes from the GPU?  (tools/reftree.py) */
int pcamv_hook_skip_hpel( x264_t *h )
{
    return g.skip_hpel_frame >= 0 && h->fdec->i_frame == g.skip_hpel_frame && h->sh.i_type == SLICE_TYPE_P;
}

/* P slices: is the intra analysis of this macroblock more than a statistic?  (host/build_host.py stats_only_intra) */
int pcamv_hook_want_intra( x264_t *h )
{
    return !g.active || g.host_intra || g.mbs[h->mb.i_mb_xy].early_skip == 2;
}

/* Pass 1 of an embedding P frame without the host.  Everything the reference's first slice pass produces for the second one
 * comes out of the GPU: the decisions and the candidate table (pcamv_analyse_p above), cover / rho_final (pcamv_embed_prepare),
 * the stego vector and the flips (pcamv_embed_stc).  What is left for the host is what encoder/encoder.c:1826-1855 does around
 * the trellis — message length, the message bits from rand(), filp[] / num_filp — and h->info.cache[] as analyse.c:3518-3680
 * would have left it, which the host's pass-2 analysis reads (analyse.c:2658-3105).  Returns 1: x264_slice_write returns
 * without entering its macroblock loop (no entropy coding, reconstruction or filtering of a pass whose bits are thrown away,
 * encoder/encoder.c:2380-2390). */
void pcamv_glue_stc_embed_prepared( x264_t *h, int an );
int pcamv_hook_pass1_on_device( x264_t *h )
{
    int i, n = -1, an;
    float rate = h->param.eparam.iEmRate;
    double t0;
    if( !g.active || g.pass != 1 || !g.direct )
        return 0;
    t0 = now_s();
    if( pcamv_embed_prepare( g.ctx, &n ) )
        die( "pcamv_embed_prepare" );
    h->info.length = n;
    if( n > 0 && pcamv_embed_download( g.ctx, h->info.cover, h->info.rho_final, NULL, NULL, NULL ) )
        die( "pcamv_embed_download" );
    memset( h->info.filp, 0, sizeof(h->info.filp) );
    memset( h->info.stego, 0, sizeof(h->info.stego) );
    if( rate > 1 ) an = rate;                       /* bits per frame */
    else an = (int)( rate * h->info.length );       /* bits per motion vector */
    for( i = 0; i < an; i++ )
        h->info.message[i] = pcamv_tls_rand() & 0x01;
    pcamv_glue_stc_embed_prepared( h, an );
    h->info.num_filp = 0;
    for( i = 0; i < h->info.length; i++ )
        if( h->info.cover[i] ^ h->info.stego[i] )
        {
            h->info.num_filp++;
            h->info.filp[i] = 1;
        }
    pcamv_hook_embed( h, an );
    if( pcamv_embed_download( g.ctx, NULL, NULL, NULL, NULL, g.pass1 ) )
        die( "pcamv_embed_download" );
    for( i = 0; i < g.n_mb; i++ )
    {
        const pcamv_pass1_mb *p = &g.pass1[i];
        h->info.cache[i].i_type = p->type;
        h->info.cache[i].i_partition = p->partition;
        h->info.cache[i].used = p->used;
        h->info.cache[i].i_qp = h->sh.i_qp;
        memcpy( h->info.cache[i].i_sub_partition, p->sub, 4 );
        memcpy( h->info.cache[i].ref, p->ref, 16 );
        memcpy( h->info.cache[i].mv, p->mv, 64 );
        memcpy( h->info.cache[i].mv_stego, p->mv_stego, 64 );
    }
    g.active = 0;
    g.n_direct++;
    g.t_gpu += now_s() - t0;
    g.t_embed += now_s() - t0;
    return 1;
}

/* after x264_macroblock_encode: the macroblocks the device cannot reconstruct on its own (records with early_skip == 2: the host
 * kept b_skip_mc set and coded the residual against what its intra analysis left in fdec, quirk q1) are handed over as they are */
void pcamv_hook_encoded( x264_t *h )
{
    if( g.active && g.recon_on && g.pass != 1 && g.mbs[h->mb.i_mb_xy].early_skip == 2 )
    {
        pcamv_recon_patch *p;
        int i, y;
        if( g.n_patches == g.cap_patches )
        {
            g.cap_patches = g.cap_patches ? 2 * g.cap_patches : 64;
            g.patches = realloc( g.patches, g.cap_patches * sizeof(*g.patches) );
            if( !g.patches ) die_msg( "out of memory" );
        }
        p = &g.patches[g.n_patches++];
        p->mb_xy = h->mb.i_mb_xy; p->nnz = 0; p->pad = 0;
        for( y = 0; y < 16; y++ ) memcpy( p->y + 16 * y, h->mb.pic.p_fdec[0] + y * FDEC_STRIDE, 16 );
        for( y = 0; y < 8; y++ )
        {
            memcpy( p->u + 8 * y, h->mb.pic.p_fdec[1] + y * FDEC_STRIDE, 8 );
            memcpy( p->v + 8 * y, h->mb.pic.p_fdec[2] + y * FDEC_STRIDE, 8 );
        }
        for( i = 0; i < 16; i++ )
            if( h->mb.cache.non_zero_count[x264_scan8[i]] )
                p->nnz |= 1 << ( block_idx_x[i] + 4 * block_idx_y[i] );
        g.recon_patched_mbs++;
    }
}

void pcamv_hook_slice_end( x264_t *h )
{
    if( g.active && g.rows_ready < g.mb_h && pcamv_analyse_p_rows( g.ctx, g.mb_h - 1, &g.rows_ready ) )     /* (every row is collected before the context's next call) */
        die( "pcamv_analyse_p_rows" );
    if( g.active && g.recon_on && g.pass != 1 && h->sh.i_type == SLICE_TYPE_P && h->fdec->b_kept_as_ref )
    {
        /* the frame just analysed in its final pass becomes a reference: build it on the GPU, in the least recently used slot
         * that none of its own references occupies */
        int i, k, best = -1;
        double t0 = now_s();
        for( i = 0; i < g.n_slots; i++ )
        {
            int busy = 0;
            for( k = 0; k < g.cur_n_ref; k++ ) busy |= g.cur_ref_slots[k] == i;
            if( !busy && ( best < 0 || g.slot[i].age < g.slot[best].age ) )
                best = i;
        }
        if( best < 0 )
            die_msg( "no free reference slot" );
        if( pcamv_reconstruct_ref( g.ctx, best, h->fdec->i_poc, g.pass, g.patches, g.n_patches ) )
            die( "pcamv_reconstruct_ref" );
        if( g.skip_hpel_frame == h->fdec->i_frame )
        {
            /* H, V and HV planes, borders included: the GPU's buffers have the host's layout (same stride and padding) */
            for( k = 1; k < 4; k++ )
            {
                const int stride = pcamv_plane_stride( g.ctx, k );
                if( stride != h->fdec->i_stride[0] )
                    die_msg( "plane stride of the GPU differs from the host's" );
                if( pcamv_get_ref_plane( g.ctx, best, k, h->fdec->filtered[k] - (size_t)32 * stride - 32 ) )
                    die( "pcamv_get_ref_plane" );
            }
            g.hpel_frames++;
        }
        g.slot[best].fr = h->fdec; g.slot[best].i_frame = h->fdec->i_frame; g.slot[best].i_poc = h->fdec->i_poc;
        g.slot[best].age = ++g.tick; g.slot[best].device_built = 1;
        g.recon_frames++;
        g.t_gpu += now_s() - t0;
        g.t_recon += now_s() - t0;
    }
    g.active = 0;
}

/* ---- per-macroblock replay --------------------------------------------------------------------- */
void pcamv_hook_analyse_begin( x264_t *h )
{
    if( g.active )
    {
        g.cur_mb = h->mb.i_mb_xy;
        g.cur_pos = 0;
        if( h->mb.i_mb_y >= g.rows_ready )
        {
            /* the wavefront has not been seen to finish this row yet: wait for it (rows arrive in order, behind the kernel) */
            double t0 = now_s();
            if( pcamv_analyse_p_rows( g.ctx, h->mb.i_mb_y, &g.rows_ready ) )
                die( "pcamv_analyse_p_rows" );
            g.t_gpu += now_s() - t0;
            g.t_row_wait += now_s() - t0;
        }
    }
}

void pcamv_hook_analyse_end( x264_t *h )
{
    if( g.active )
    {
        /* every GPU entry of this macroblock must have been consumed, and the host must have arrived at the GPU's decision */
        const pcamv_mb_out *r = &g.mbs[g.cur_mb];
        const int elided = g.elide && g.pass == 2 && h->info.cache[g.cur_mb].used;
        if( elided ? g.cur_pos < r->n_log : g.cur_pos != r->n_log )
        {
            fprintf( stderr, "x264 [pcamv]: frame %d mb %d: host made %d search calls, GPU logged %d\n", h->i_frame, g.cur_mb, g.cur_pos, r->n_log );
            fflush( stderr ); _exit( 4 );
        }
        if( h->mb.i_type != r->type || ( r->type != P_SKIP && h->mb.i_partition != r->partition ) )
        {
            fprintf( stderr, "x264 [pcamv]: frame %d mb %d: host decided type %d partition %d, GPU %d / %d\n",
                     h->i_frame, g.cur_mb, h->mb.i_type, h->mb.i_partition, r->type, r->partition );
            fflush( stderr ); _exit( 4 );
        }
    }
}

static const pcamv_log_entry *next_entry( x264_t *h, x264_me_t *m, int kind )
{
    const pcamv_log_entry *e;
    if( !g.active || g.cur_mb != h->mb.i_mb_xy )
        die_msg( "motion search outside a GPU-analysed P slice (no CPU fallback)" );
    if( g.cur_pos >= g.mbs[g.cur_mb].n_log && g.elide && g.pass == 2 && h->info.cache[g.cur_mb].used && kind != PCAMV_LOG_IHCOST )
    {
        /* a search of pass 2 whose result the reference overwrites at analyse.c:2868-2991: not executed on the GPU */
        g.cur_pos++;
        return NULL;
    }
    if( g.cur_pos >= g.log_stride || g.cur_pos >= g.mbs[g.cur_mb].n_log )
    {
        fprintf( stderr, "x264 [pcamv]: frame %d mb %d: host asks for call %d, GPU logged %d\n", h->i_frame, g.cur_mb, g.cur_pos, g.mbs[g.cur_mb].n_log );
        fflush( stderr ); _exit( 4 );
    }
    e = &g.log[(size_t)g.cur_mb * g.log_stride + g.cur_pos];
    if( e->kind != kind || e->i_pixel != m->i_pixel )
    {
        fprintf( stderr, "x264 [pcamv]: frame %d mb %d call %d: host wants kind %d pixel %d, GPU logged kind %d pixel %d\n",
                 h->i_frame, g.cur_mb, g.cur_pos, kind, m->i_pixel, e->kind, e->i_pixel );
        fflush( stderr ); _exit( 4 );
    }
    g.cur_pos++;
    g.n_replayed++;
    return e;
}

void x264_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh )
{
    const pcamv_log_entry *e = next_entry( h, m, PCAMV_LOG_SEARCH );
    (void)mvc; (void)i_mvc;
    /* the multi-reference half-pel threshold only ever feeds later searches, which are replayed too */
    (void)p_halfpel_thresh;
    if( !e )
    {
        /* dead value: large enough that no partition built from it is ever preferred, small enough that sums do not overflow */
        m->mv[0] = m->mv[1] = 0; m->cost = 1 << 26; m->cost_mv = 0;
        return;
    }
    m->mv[0] = e->mv[0]; m->mv[1] = e->mv[1];
    m->cost = e->cost;
    m->cost_mv = e->cost_mv;
}

void x264_me_refine_qpel( x264_t *h, x264_me_t *m )
{
    const pcamv_log_entry *e = next_entry( h, m, PCAMV_LOG_REFINE );
    if( !e )
        return;         /* dead refinement of pass 2: the vector is overwritten by the forcing block */
    m->mv[0] = e->mv[0]; m->mv[1] = e->mv[1];
    m->cost = e->cost;
    m->cost_mv = e->cost_mv;
}

/* called by the x264_ih_get_mv_cost wrapper host/build_host.py puts into encoder/analyse.c */
int pcamv_glue_ih_cost( x264_t *h, x264_me_t *m, int16_t *m_x, int16_t *m_y )
{
    const pcamv_log_entry *e = next_entry( h, m, PCAMV_LOG_IHCOST );
    *m_x = e->mv[0]; *m_y = e->mv[1];
    return e->cost;
}

/* hooks of the instrumented oracle twin that the GPU host does not need */
/* end of the embed stage (encoder/encoder.c:1855): with PCAMV_PAYLOAD=<file> append what was hidden in this frame —
 * int32 frame, length, an; uint8 message[an]; uint8 stego[length] — so that payload parity can be checked directly */
void pcamv_hook_embed( x264_t *h, int an )
{
    const char *s = getenv( "PCAMV_PAYLOAD" );
    if( s && *s )
    {
        FILE *f = open_side_file( s );
        if( f )
        {
            int32_t hd[3] = { h->i_frame, h->info.length, an };
            fwrite( hd, 4, 3, f );
            if( an > 0 ) fwrite( h->info.message, 1, an, f );
            fwrite( h->info.stego, 1, h->info.length, f );
            fclose( f );
        }
    }
    /* PCAMV_STEGO=<file>: the extractor's input (x264_pcamv --extract) — frame, length, an and the stego LSBs only; a frame whose
     * embedding failed is recorded with an = 0 ("carries nothing"), so that the extractor passes it by */
    s = getenv( "PCAMV_STEGO" );
    if( s && *s )
    {
        FILE *f = open_side_file( s );
        if( f )
        {
            int32_t hd[3] = { h->i_frame, h->info.length, g_embedded ? an : 0 };
            fwrite( hd, 4, 3, f );
            fwrite( h->info.stego, 1, h->info.length, f );
            fclose( f );
        }
    }
}
void pcamv_hook_ih_satd( int i_pixel, int b_chroma_me ) { (void)i_pixel; (void)b_chroma_me; }
