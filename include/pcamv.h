/* pcamv.h — C ABI of libpcamv_cuda.so: the B200 (sm_100a) implementation of the PCAMV encoder's
 * motion-estimation hot path.  Plain C, plain pointers and sizes; no CUDA or torch types.
 *
 * The reference (x264 build 66 + PCAMV, /root/reference) has no FFI layer; the seams this library
 * replaces are (file:line in the reference):
 *
 *   leaf seam     x264_pixel_function_t   common/pixel.h:63-103   filled by x264_pixel_init (pixel.c:565)
 *                 x264_mc_functions_t     common/mc.h:31-77       filled by x264_mc_init   (mc.c:406)
 *   search seam   x264_me_search_ref      encoder/me.h:58  (me.c:158)
 *                 x264_me_refine_qpel     encoder/me.h:62  (me.c:669)
 *                 x264_ih_get_mv_cost     encoder/analyse.c:2391 (static; the stego cost table)
 *   frame seam    x264_frame_filter       common/mc.c:453  (+ x264_frame_expand_border[_filtered],
 *                                         common/frame.c:246,275), called from x264_fdec_filter_row
 *                 P-slice body of x264_macroblock_analyse   encoder/analyse.c:2613-3172,3518-3689
 *
 * Conventions follow the reference: every entry point returns 0 on success and -1 on failure
 * (the reference's `int` + x264_log convention, common/common.h:39-47); the message is available
 * from pcamv_last_error().  The caller owns all host memory; the context owns device memory, its
 * CUDA stream and pinned staging.  A context is thread-compatible (one encoder thread per context,
 * like one x264_t), not thread-safe.  There is NO CPU fallback: without a usable CUDA device
 * pcamv_open fails.  CUDA errors are sticky: after the first failure every call returns -1.
 */
#ifndef PCAMV_H
#define PCAMV_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCAMV_ABI_VERSION 6
#define PCAMV_MAX_REFS 16
#define PCAMV_MAX_MVC 10

/* block sizes: same numbering as the reference's PIXEL_WxH enum (common/pixel.h:30-42) */
enum { PCAMV_PIXEL_16x16 = 0, PCAMV_PIXEL_16x8, PCAMV_PIXEL_8x16, PCAMV_PIXEL_8x8,
       PCAMV_PIXEL_8x4, PCAMV_PIXEL_4x8, PCAMV_PIXEL_4x4 };
/* search methods: X264_ME_* (x264.h) */
enum { PCAMV_ME_DIA = 0, PCAMV_ME_HEX, PCAMV_ME_UMH, PCAMV_ME_ESA, PCAMV_ME_TESA };
/* macroblock types / partitions as numbered by the reference (common/macroblock.h:29-60) */
enum { PCAMV_P_L0 = 4, PCAMV_P_8x8 = 5, PCAMV_P_SKIP = 6 };
enum { PCAMV_D_L0_4x4 = 0, PCAMV_D_L0_8x4 = 1, PCAMV_D_L0_4x8 = 2, PCAMV_D_L0_8x8 = 3,
       PCAMV_D_8x8 = 13, PCAMV_D_16x8 = 14, PCAMV_D_8x16 = 15, PCAMV_D_16x16 = 16 };

typedef struct pcamv_ctx pcamv_ctx;

/* Encoder-wide configuration: the x264_param_t fields the path reads (x264.h:154-311) plus geometry. */
typedef struct pcamv_cfg
{
    int abi_version;        /* PCAMV_ABI_VERSION */
    int device;             /* CUDA device ordinal */
    int width, height;      /* luma size in pixels, already rounded up to multiples of 16 */
    int me_method;          /* param.analyse.i_me_method */
    int me_range;           /* param.analyse.i_me_range (after validation) */
    int subpel_refine;      /* param.analyse.i_subpel_refine; this build supports 1..5 for frame analysis */
    int chroma_me;          /* param.analyse.b_chroma_me */
    int max_refs;           /* param.i_frame_reference */
    int mv_range;           /* param.analyse.i_mv_range in pixels (level default, encoder/encoder.c:558-561) */
    int b_cabac;            /* param.b_cabac */
    int b_fast_pskip;       /* param.analyse.b_fast_pskip */
    int b_dct_decimate;     /* param.analyse.b_dct_decimate */
    int analyse_inter;      /* param.analyse.inter flag word (X264_ANALYSE_PSUB16x16 = 0x10, PSUB8x8 = 0x20) */
    int chroma_qp_offset;   /* pps chroma_qp_index_offset */
    int rows_per_cta;       /* wavefront layout: 0/1 = one macroblock row per CTA (lowest latency, a single encoder);
                               2 or 4 = consecutive rows share a CTA (throughput, many concurrent contexts per GPU);
                               -1 = row pool for multi-context launches: rows are resumable tasks claimed by any team whose
                               next macroblock is ready, so no team sleeps on a dependency (single launches: as 1) */
    int pass2_elide;        /* 1: in pass 2, macroblocks whose decision is forced from pass 1 (info.cache[].used) only run the
                               16x16 search — the one result of that pass the reference still uses (h->mb.mvr candidates of later
                               macroblocks, early-skip detection); the 8x8 / 16x8 / 8x16 searches and the refinement, whose
                               results the reference overwrites at encoder/analyse.c:2868-2991, are not executed and do not
                               appear in the log.  Exempt (full analysis, complete log): macroblocks whose pass-2 probe
                               finds them skippable although pass 1 coded them — the host keeps b_skip_mc set there and
                               its residual depends on the intra analysis, which prunes against those "dead" costs.
                               Ignored with sub-8x8 partitions (a forced P_8x8 keeps the partition pass 2 decided).
                               0: pass 2 executes and logs everything the reference executes. */
    int no_deblock;         /* sh.i_disable_deblocking_filter_idc == 1 (--nf); the three deblocking fields matter to pcamv_reconstruct_ref only */
    int deblock_alpha_c0_offset, deblock_beta_offset;   /* sh.i_alpha_c0_offset, sh.i_beta_offset (the --deblock values, doubled) */
    int reserved[3];
} pcamv_cfg;

/* Per-QP tables.  They are built on the host because the reference builds cost_mv with float
 * log under -ffast-math (encoder/analyse.c:46,193-229); the device only ever reads them. */
typedef struct pcamv_qp_tables
{
    int qp;                          /* luma QP the tables are for */
    int lambda;                      /* x264_lambda_tab[qp]  (analyse.c:148) */
    int lambda2_chroma;              /* x264_lambda2_tab[chroma qp], for the P_SKIP chroma SSD gate */
    int chroma_qp;                   /* h->chroma_qp_table[qp] */
    const int16_t *cost_mv;          /* 32769 entries, entry 16384 <-> mv difference 0 */
    const uint16_t *cost_ref;        /* 3*33 entries: x264_cost_ref[qp] */
    const uint16_t *quant4_mf[2];    /* [0]=CQM_4PY(inter luma) [1]=CQM_4PC(inter chroma): 16 entries for this qp / chroma qp */
    const uint16_t *quant4_bias[2];  /* same indexing */
    const int32_t *dequant4_mf[2];   /* 6*16 entries each: h->dequant4_mf[CQM_4PY / CQM_4PC] */
} pcamv_qp_tables;

/* One x264_me_search_ref (mode 0) or x264_me_refine_qpel (mode 1) call: the x264_me_t inputs
 * (encoder/me.h:30-51) plus the h->mb.* state those functions read. */
typedef struct pcamv_me_call
{
    int32_t mode;                    /* 0 = search, 1 = refine_qpel */
    int32_t mb_x, mb_y;              /* macroblock position */
    int32_t xoff, yoff;              /* block offset inside the macroblock, pixels */
    int32_t i_pixel;                 /* PCAMV_PIXEL_* */
    int32_t ref_slot;                /* reference slot uploaded with pcamv_put_ref* */
    int32_t i_ref_cost;
    int32_t mv_min_fpel[2], mv_max_fpel[2], mv_min_spel[2], mv_max_spel[2];
    int32_t i_mvc, has_thresh, thresh_in;
    int16_t mvp[2];
    int16_t mvc[PCAMV_MAX_MVC][2];
    int16_t mv_in[2];                /* refine only */
    int32_t cost_in, cost_mv_in;     /* refine only */
} pcamv_me_call;

typedef struct pcamv_me_result
{
    int16_t mv[2];
    int32_t cost, cost_mv, thresh_out;
} pcamv_me_result;

/* ---- lifetime ---------------------------------------------------------------------------------- */
/* replaces: nothing in the reference; called from x264_encoder_open after mbcmp_init (encoder/encoder.c:766) */
int pcamv_open(pcamv_ctx **out, const pcamv_cfg *cfg);
/* called from x264_encoder_close (encoder/encoder.c:2670) */
void pcamv_close(pcamv_ctx *ctx);
const char *pcamv_last_error(const pcamv_ctx *ctx);   /* ctx may be NULL: error of the last failed pcamv_open */
int pcamv_abi_version(void);

/* Page-locked host memory for frame planes and result arrays: copies to / from such buffers go straight over PCIe
 * (pageable buffers work too, through the context's own pinned staging).  Zero-filled; NULL on failure. */
void *pcamv_host_alloc(size_t bytes);
void pcamv_host_free(void *p);

/* ---- tables -------------------------------------------------------------------------------------- */
/* called when x264_mb_analyse_load_costs first sees a QP (encoder/analyse.c:198) */
int pcamv_set_qp_tables(pcamv_ctx *ctx, const pcamv_qp_tables *t);

/* ---- frames -------------------------------------------------------------------------------------- */
/* Source frame after x264_frame_copy_picture + mod16 expansion (encoder/encoder.c:2163-2167).
 * y/u/v point at pixel (0,0); strides in bytes. */
int pcamv_put_fenc(pcamv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int stride_y, int stride_c);

/* Reconstructed + deblocked reference frame (end of x264_fdec_filter_row for a kept frame,
 * encoder/encoder.c:2030).  The GPU replicates the borders (common/frame.c:224-273), builds the H, V
 * and HV half-pel planes (common/mc.c:134-190,453-475), replicates their borders (frame.c:275-301) and,
 * for --me esa, the integral image (mc.c:311-345,477-511). */
int pcamv_put_ref(pcamv_ctx *ctx, int slot, int poc, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                  int stride_y, int stride_c);
/* Test/debug: upload the six planes of a reference exactly as the host holds them (padded buffers:
 * luma planes start 32 rows / 32 columns before pixel (0,0), chroma 16/16), bypassing the GPU filter. */
int pcamv_put_ref_planes(pcamv_ctx *ctx, int slot, int poc, const uint8_t *const luma_padded[4],
                         const uint8_t *u_padded, const uint8_t *v_padded);
/* Copy one device plane of a slot back (padded buffer, same layout as above).  plane: 0..3 luma
 * (integer, H, V, HV), 4 = U, 5 = V, 6 = the integral plane (--me esa / tesa contexts only: uint16 8x8 box sums, same
 * geometry as a luma plane, reference common/mc.c:311-345).  dst must hold pcamv_plane_bytes(). */
int pcamv_get_ref_plane(pcamv_ctx *ctx, int slot, int plane, uint8_t *dst);
size_t pcamv_plane_bytes(const pcamv_ctx *ctx, int plane);
int pcamv_plane_stride(const pcamv_ctx *ctx, int plane);

/* ---- search seam ---------------------------------------------------------------------------------- */
/* n independent x264_me_search_ref / x264_me_refine_qpel evaluations against the current fenc and
 * the uploaded reference slots (1:1 stand-in for encoder/me.h:58,62; used by the parity tests and
 * by the stateless throughput benchmark). */
int pcamv_me_search_batch(pcamv_ctx *ctx, const pcamv_me_call *calls, int n, pcamv_me_result *results);

/* Benchmark support: keep the batch resident on the device and time repeated launches with CUDA
 * events on the context's stream.  pcamv_me_batch_upload stages calls once; pcamv_me_batch_run
 * launches the search kernel `iters` times and returns the mean kernel time in milliseconds. */
int pcamv_me_batch_upload(pcamv_ctx *ctx, const pcamv_me_call *calls, int n);
int pcamv_me_batch_run(pcamv_ctx *ctx, int iters, float *ms_per_launch);
int pcamv_me_batch_download(pcamv_ctx *ctx, pcamv_me_result *results, int n);

/* ---- frame seam: the P-slice body of x264_macroblock_analyse ------------------------------------------------ */
#define PCAMV_LOG_MAX 112

/* One entry of a macroblock's result log.  The analysis of a macroblock calls x264_me_search_ref,
 * x264_me_refine_qpel and (pass 1) x264_ih_get_mv_cost in a data-dependent order (encoder/analyse.c:2646-2810,
 * 3518-3689); the GPU appends one entry per call, in that order, and the host encoder replays them instead of
 * searching. */
enum { PCAMV_LOG_SEARCH = 0, PCAMV_LOG_REFINE = 1, PCAMV_LOG_IHCOST = 2 };
typedef struct pcamv_log_entry
{
    int8_t kind, i_pixel, i_ref, pad;
    int16_t mv[2];      /* search / refine: m->mv;  ih-cost: the chosen replacement delta (m_x, m_y) */
    int32_t cost;       /* search / refine: m->cost;  ih-cost: cost_opt (the embedding cost) */
    int32_t cost_mv;    /* search / refine: m->cost_mv */
} pcamv_log_entry;

/* Per-macroblock outcome of the analysis (what x264_macroblock_analyse leaves in h->mb for a P macroblock). */
typedef struct pcamv_mb_out
{
    int8_t type, partition, n_part, early_skip;   /* PCAMV_P_*, PCAMV_D_*, MV-carrying partitions; early_skip: 1 = early P_SKIP exit,
                                                   * 2 = this pass's probe said skippable but the forced decision codes it (the host keeps b_skip_mc: quirk q1) */
    int8_t ref[4];                                /* reference index per 8x8 block */
    int16_t mv[16][2];                            /* h->mb.cache.mv[0][x264_scan8[i]], block_idx order */
    struct { int16_t mv[2]; int16_t mvp[2]; int8_t ref, i_pixel, xoff, yoff; } part[4];
    int32_t n_log;                                /* entries used in this macroblock's log */
    int16_t pskip_mv[2];
} pcamv_mb_out;

/* The fields of the reference's h->info.cache[mb] (common/common.h:585-603) that pass 2 reads, verbatim —
 * including how analyse.c:3526-3632 fills them (ref[] raster, mv[] through the unsequenced idx++ copy). */
typedef struct pcamv_pass1_mb
{
    int32_t type;            /* PCAMV_P_L0 / PCAMV_P_8x8 / PCAMV_P_SKIP */
    int32_t partition;       /* PCAMV_D_16x8 / 8x16 / 16x16 */
    uint8_t used;
    uint8_t sub[4];
    int8_t ref[16];
    int16_t mv[16][2];
    int16_t mv_stego[16][2];
} pcamv_pass1_mb;

typedef struct pcamv_frame_in
{
    int32_t pass;                       /* 0 = embedding off, 1 = pre-encode, 2 = final encode (info.firstTime == 0) */
    int32_t n_ref;                      /* h->i_ref0 */
    int32_t ref_slot[PCAMV_MAX_REFS];   /* slot of h->fref0[i] as uploaded with pcamv_put_ref */
    int32_t ref_poc[PCAMV_MAX_REFS];    /* h->fref0[i]->i_poc */
    int32_t cur_poc;                    /* h->fdec->i_poc */
    int32_t col_n_ref;                  /* h->fref0[0]->i_ref[0]; <= 0 disables the temporal candidates */
    int32_t col_inv_ref_poc[PCAMV_MAX_REFS];   /* h->fref0[0]->inv_ref_poc[] */
    const int8_t *col_ref8;             /* h->fref0[0]->ref[0]: [2*mb_h][2*mb_w] */
    const int16_t *col_mv4;             /* h->fref0[0]->mv[0]:  [4*mb_h][4*mb_w][2] */
    const pcamv_pass1_mb *pass1;        /* pass 2: h->info.cache[0..n_mb) */
    const int8_t *filp;                 /* pass 2: h->info.filp[0..n_filp) */
    int32_t n_filp;
    int32_t cost_table;                 /* pass 1: also build the candidate-MV cost table (emrate != 0) */
    int16_t stale_mv[16][2];            /* h->mb.cache.mv[0][x264_scan8[i]] as left by the previous slice pass */
    int32_t device_forced;              /* pass 2: the forced decisions are already in HBM (pcamv_embed_stc built them from this
                                         * context's pass 1): pass1 / filp / n_filp are ignored */
} pcamv_frame_in;

/* Switch pcamv_cfg.pass2_elide of an open context (takes effect with the next launch). */
int pcamv_set_pass2_elide(pcamv_ctx *ctx, int on);

/* Conformance switch of an open context (0 by default = the reference's behaviour, bit for bit; takes effect with the next
 * launch).  With it on, the two places where the DEVICE reproduces statements of the reference that make its embedding
 * streams unreadable for a standard decoder are corrected: a macroblock that pass 2 forces to P_SKIP gets the skip
 * predictor as its vector (the reference returns without x264_analyse_update_cache, encoder/analyse.c:2677-2680, quirk q2),
 * and pcamv_embed_prepare copies a macroblock's vectors into the info.cache[] record straight (the reference's unsequenced
 * `idx++` copy, analyse.c:3537-3543 / 3626-3632, SURVEY fact 3, hands slot k the vector of another block).  The host side
 * of the same switch is tools/reftree.py::conformance_switch (PCAMV_CONFORMANT=1 in the bound host); with both on, the
 * payload can be extracted from the .264 alone (host/pcamv_bitstream.c).  The bitstream then differs from the reference's
 * by design; the parity reference for this mode is oracle/_ref/x264_dump_conformant. */
int pcamv_set_conformant(pcamv_ctx *ctx, int on);

/* Entries per macroblock in the log arrays of this context: the most its configuration can produce (14 -> 16 for one
 * reference frame), never more than PCAMV_LOG_MAX.  Entry k of macroblock mb is log[mb * pcamv_log_stride(ctx) + k]. */
int pcamv_log_stride(const pcamv_ctx *ctx);

/* Analyse every macroblock of a P slice against the current fenc and the uploaded references.
 * mbs: n_mb records; log: n_mb * pcamv_log_stride(ctx) entries (may be NULL).  Replaces, per macroblock, the searches of
 * x264_macroblock_analyse (encoder/encoder.c:1273 -> encoder/analyse.c:2555) for P slices with subme <= 5. */
int pcamv_analyse_p(pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log);

/* Benchmark / pipelining support: the three stages of pcamv_analyse_p separately.  _upload stages the frame
 * inputs of in->pass in HBM (one set per pass is kept, so pass 1 and pass 2 of a frame can both be resident);
 * _run launches the wavefront (+ cost table) of `pass` (-1: the last uploaded) `iters` times and returns the mean
 * device time of one analysis in milliseconds (CUDA events on the context's stream) and, when ms_kernels is not
 * NULL, the mean duration of its two kernels (ms_kernels[0] wavefront analysis, [1] cost table; events recorded
 * around each launch on the same stream); _download copies the results of the last run back. */
int pcamv_frame_upload(pcamv_ctx *ctx, const pcamv_frame_in *in);
int pcamv_frame_run(pcamv_ctx *ctx, int pass, int iters, float *ms_per_frame, float *ms_kernels);
int pcamv_frame_download(pcamv_ctx *ctx, pcamv_mb_out *mbs, pcamv_log_entry *log);

/* Multi-context launches: n encoder contexts of equal geometry and search configuration on one device (GOP shards or
 * independent streams) analysed by ONE wavefront kernel (+ one cost-table kernel), so that a B200 is
 * filled by many frames' wavefronts at once without one stream per context.  ctxs[0] leads (its stream carries the
 * launch); results land in each member's own buffers.  pcamv_analyse_p_batch = upload each, launch once, download
 * each; pcamv_frame_run_batch re-launches frames already staged with pcamv_frame_upload (benchmark / pipelining),
 * returning the mean device time of one launch pair and, in ms_kernels[2], of the wavefront and cost-table kernels. */
int pcamv_analyse_p_batch(pcamv_ctx *const *ctxs, const pcamv_frame_in *const *ins, int n,
                          pcamv_mb_out *const *mbs, pcamv_log_entry *const *logs);
int pcamv_frame_run_batch(pcamv_ctx *const *ctxs, int n, int pass, int iters, float *ms_per_step, float *ms_kernels);

/* Encoder groups: n encoder threads of one process (one per GOP shard / stream, each with its own context) share the
 * GPU through multi-context launches without knowing about each other.  pcamv_group_analyse_p is pcamv_analyse_p for a
 * member: it blocks until every live member has submitted its frame, the last arriver launches for all, and each call
 * returns with its own results.  A member whose encoder has no more frames calls pcamv_group_leave (once). */
typedef struct pcamv_group pcamv_group;
int pcamv_group_create(pcamv_group **out, int n_members);
void pcamv_group_destroy(pcamv_group *g);
int pcamv_group_analyse_p(pcamv_group *g, pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log);
int pcamv_group_leave(pcamv_group *g);

/* Results while the wavefront is still running.  The host that replays a P slice (host/pcamv_x264_glue.c) walks the macroblocks
 * in raster order, as x264_slice_write does (encoder/encoder.c:1240-2010), and needs row r only when it gets there; the
 * wavefront finishes row r after (mb_w + 2 r) of its (mb_w + 2 mb_h) steps.  pcamv_analyse_p_begin / _batch_begin /
 * pcamv_group_analyse_p_begin are pcamv_analyse_p / _batch / pcamv_group_analyse_p without the wait for the end of the kernel
 * and without the download: they return once the launch is in flight.  pcamv_analyse_p_rows( ctx, row, &n ) then blocks until
 * the records and log entries of rows 0..row are in the buffers given to _begin (n = rows complete so far, >= row + 1) — a
 * second stream follows the kernel's row counters and copies finished rows out behind it.  mbs / log must be page-locked
 * (pcamv_host_alloc) and stay untouched by the caller until their rows are reported.  Every row must be collected (row =
 * mb_h - 1 at the latest) before the context's next call; meant for the pass whose results the host replays (0 or 2). */
int pcamv_analyse_p_begin(pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log);
int pcamv_analyse_p_batch_begin(pcamv_ctx *const *ctxs, const pcamv_frame_in *const *ins, int n,
                                pcamv_mb_out *const *mbs, pcamv_log_entry *const *logs);
int pcamv_group_analyse_p_begin(pcamv_group *g, pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log);
int pcamv_analyse_p_rows(pcamv_ctx *ctx, int row, int *rows_ready);

/* Profiling aid: with enable != 0 the next wavefront launches record the device globaltimer (ns) at the start and end
 * of every macroblock; out (may be NULL) receives the [n_mb][2] records of the last traced launch. */
int pcamv_frame_trace(pcamv_ctx *ctx, int enable, unsigned long long *out);

/* Measured integer-pipe issue peak of the device, in giga lane-operations/s: a microbenchmark of the
 * VABSDIFF4 / IADD3 / LOP3 mix the SAD and SATD loops consist of.  Roofline denominator for the search kernels. */
int pcamv_int_peak(pcamv_ctx *ctx, double *gops);

/* ---- embed stage: the syndrome-trellis code of the embedder (reference embed.h:309-548 stc_embed) -----
 * Replaces the call at encoder/encoder.c:1843: Viterbi over 2^matrixheight states on the GPU (forward pass: one CTA, one
 * thread per state; backward trace: one warp).  cover / message / stego are one byte per bit, rho the float embedding
 * costs, all of the host's own buffers; cols_short / cols_long are the two sub-matrices the host drew with the reference's
 * getMatrix( floor(n/an) ), getMatrix( ceil(n/an) ) (embed.h:276-306; drawing them is stateful beyond width 20, so it
 * stays with the caller).  Returns 0 = stego written (bit-identical to stc_embed's), 1 = the message is not embeddable in
 * this cover (stc_embed returns 0 and leaves stego untouched; so does this), -1 = error. */
int pcamv_stc_embed(pcamv_ctx *ctx, const uint8_t *cover, int n, const uint8_t *message, int an, const float *rho,
                    uint8_t *stego, int matrixheight, const uint32_t *cols_short, int w_short,
                    const uint32_t *cols_long, int w_long);

/* ---- the embed stage between the two passes, device-resident (reference encoder/encoder.c:1561-1855) ----------------------
 * After pcamv_analyse_p( pass 1, cost_table = 1 ) the frame's records and cost table are in HBM.  pcamv_embed_prepare builds
 * from them, on the device, what the reference's embed stage builds on the host: the h->info.cache[] entry of every macroblock
 * (encoder/analyse.c:3526-3689, including the motion vectors as the unsequenced copy loops leave them), the cover bits
 * LSB(mvx + mvy) and the float costs rho_final with the MVC penalties (x 2, x (0.7 n + 1)), in the reference's carrier order;
 * *length = h->info.length.  The host then draws its message (an = (int)(rate * length) bits of rand() & 1,
 * encoder/encoder.c:1828-1840) and calls pcamv_embed_stc: the syndrome-trellis code runs on the device-resident cover / rho
 * (same contract and return value as pcamv_stc_embed; total = the reference's sum of rho_final in double, or < 0 to have it
 * summed on the device), then filp = cover ^ stego and the decisions pass 2 forces (encoder/analyse.c:2870-3107) are built in
 * HBM; an <= 0 or message == NULL embeds nothing (stego stays zero, as when the reference's stc_embed gives up).  stego (may be
 * NULL) receives the length stego bits.  pcamv_analyse_p( pass 2 ) with in->device_forced = 1 then needs neither the pass-1
 * records nor filp from the host.  pcamv_embed_download copies out whatever the host wants to see (any pointer may be NULL). */
int pcamv_embed_prepare(pcamv_ctx *ctx, int *length);
int pcamv_embed_stc(pcamv_ctx *ctx, const uint8_t *message, int an, int matrixheight, const uint32_t *cols_short, int w_short,
                    const uint32_t *cols_long, int w_long, double total, uint8_t *stego);
int pcamv_embed_download(pcamv_ctx *ctx, uint8_t *cover, float *rho, uint8_t *stego, int8_t *filp, pcamv_pass1_mb *pass1);

/* ---- the reference frame built on the device (reference encoder/macroblock.c:605-755, common/frame.c:594-800,
 * encoder/encoder.c:1004-1052) ------------------------------------------------------------------------------------------------
 * After the FINAL pass of a P frame (pass 2, or pass 0 when nothing is embedded) pcamv_reconstruct_ref rebuilds the frame in
 * reference slot `slot` from the decisions that pass left in HBM: motion compensation + residual coding of every macroblock,
 * the in-loop deblocking filter, then borders, half-pel planes and integral planes as pcamv_put_ref builds them — byte for
 * byte the planes the host's x264_fdec_filter_row produces, so the next frame needs no pcamv_put_ref for this picture.
 * One class of macroblocks the device cannot reconstruct on its own: records with early_skip == 2 (the host keeps b_skip_mc
 * set and codes the residual against what its intra analysis left in fdec).  For those the caller passes the
 * host's own reconstruction BEFORE deblocking (h->mb.pic.p_fdec right after x264_macroblock_encode) and the surviving-coefficient
 * flags of the 16 luma blocks (bit x + 4 y = non_zero_count of raster block (x, y) != 0) as patches. */
typedef struct pcamv_recon_patch
{
    int32_t mb_xy;
    uint16_t nnz, pad;
    uint8_t y[256], u[64], v[64];       /* rows of 16 / 8 / 8 pixels */
} pcamv_recon_patch;
int pcamv_reconstruct_ref(pcamv_ctx *ctx, int slot, int poc, int pass, const pcamv_recon_patch *patches, int n_patches);

/* Number of kernel launches issued by this context so far (for bench.py's gpu_launches). */
long long pcamv_launch_count(const pcamv_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PCAMV_H */
