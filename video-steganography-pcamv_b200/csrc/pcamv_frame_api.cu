// pcamv_frame_api.cu — C-ABI of the frame seam (include/pcamv.h: pcamv_analyse_p and its three stages).
// Host-side only: argument checks, the pass-2 glue (pcamv_glue.h), staging, launches.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <vector>
#include "pcamv_ctx.h"
#include "pcamv_glue.h"
#include "pcamv_split.h"

using namespace pcamv;

namespace pcamv {
void launch_analyse_p(const DevFrameCtx &fc, const FrameParams &fp, int *row_claim, int n_rows, int rows_per_cta, void *stream);
void launch_cost_table(const DevFrameCtx &fc, const FrameParams &fp, int n_mb, void *stream);
void launch_analyse_p_batch(const BatchItem *items, int n_items, int *row_claim, int n_rows, int rows_per_cta, int max_ctas, const DevFrameCtx &fc, void *stream);
void launch_cost_table_batch(const BatchItem *items, int n_items, int n_mb, void *stream);
}

// the ABI records are the device records
static_assert(sizeof(pcamv_log_entry) == sizeof(LogEntry) && sizeof(LogEntry) == 16, "log entry layout");
static_assert(sizeof(pcamv_mb_out) == sizeof(MbResult) && sizeof(MbResult) == 128, "mb_out layout");
static_assert(offsetof(pcamv_mb_out, part) == offsetof(MbResult, part) && offsetof(pcamv_mb_out, n_log) == offsetof(MbResult, n_log), "mb_out layout");
static_assert(sizeof(pcamv_pass1_mb) == sizeof(Pass1Mb) && offsetof(pcamv_pass1_mb, mv_stego) == offsetof(Pass1Mb, mv_stego), "pass1 layout");
static_assert(sizeof(ForcedOut) == sizeof(ForcedMb), "forced layout");

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, #call, e_); } while (0)
#define GUARD() do { if (!ctx) return -1; if (ctx->failed) return -1; cudaSetDevice(ctx->cfg.device); } while (0)   /* calls may come from any host thread */

static int ensure_frame_buffers(pcamv_ctx *ctx)
{
    if (ctx->d_log) return 0;
    const DevFrameCtx &fc = ctx->fc;
    const size_t n_mb = (size_t)fc.mb_w * fc.mb_h;
    CK(cudaMalloc(&ctx->fa.type, n_mb));
    CK(cudaMalloc(&ctx->fa.ref8, 4 * n_mb));
    CK(cudaMalloc(&ctx->fa.mv4, 16 * n_mb * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->fa.mvr, (size_t)PCAMV_MAX_REFS * n_mb * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->d_col_ref8, 4 * n_mb));
    CK(cudaMalloc(&ctx->d_col_mv4, 16 * n_mb * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->d_forced, n_mb * sizeof(ForcedMb)));
    CK(cudaMalloc(&ctx->d_mb_results, n_mb * sizeof(MbResult)));
    CK(cudaMalloc(&ctx->d_progress, (2 * fc.mb_h + 2) * sizeof(int)));      // row progress | row-claim counter | row owners (row pool)
    CK(cudaMalloc(&ctx->d_trace, 2 * n_mb * sizeof(unsigned long long)));
    if (fc.analyse_inter & 0x20)
    {
        CK(cudaMalloc(&ctx->d_subparts, 16 * n_mb * sizeof(PartInfo)));
        CK(cudaMemsetAsync(ctx->d_subparts, 0, 16 * n_mb * sizeof(PartInfo), ctx->stream));
    }
    if (fc.me_method == PCAMV_ME_TESA)
    {
        // every position of the widest window can end up in a team's list (reference: h->scratch_buffer)
        ctx->mvsads_cap = (2 * fc.me_range + 4) * (2 * fc.me_range + 1);
        CK(cudaMalloc(&ctx->d_mvsads, (size_t)fc.mb_h * ctx->mvsads_cap * sizeof(unsigned long long)));
    }
    CK(cudaMemsetAsync(ctx->fa.type, 0, n_mb, ctx->stream));
    CK(cudaMemsetAsync(ctx->fa.ref8, 0, 4 * n_mb, ctx->stream));
    CK(cudaMemsetAsync(ctx->fa.mv4, 0, 16 * n_mb * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->fa.mvr, 0, (size_t)PCAMV_MAX_REFS * n_mb * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_mb_results, 0, n_mb * sizeof(MbResult), ctx->stream));
    // pinned staging: inputs (col ref/mv + forced) and outputs (results + log)
    ctx->h_frame_in_bytes = (n_mb * (4 + 64 + sizeof(ForcedMb)) + 255) & ~(size_t)255;
    ctx->h_frame_bytes = ctx->h_frame_in_bytes + n_mb * (sizeof(MbResult) + ctx->log_stride * sizeof(LogEntry)) + 1024;
    CK(cudaMallocHost(&ctx->h_frame, ctx->h_frame_bytes));
    CK(cudaMalloc(&ctx->d_log, n_mb * ctx->log_stride * sizeof(LogEntry)));
    CK(cudaMemsetAsync(ctx->d_log, 0, n_mb * ctx->log_stride * sizeof(LogEntry), ctx->stream));
    return 0;
}

// stages the frame inputs and issues the copies on the context's stream; the caller synchronises
static int frame_upload_async(pcamv_ctx *ctx, const pcamv_frame_in *in)
{
    GUARD();
    if (!in) return ctx_fail(ctx, "pcamv_frame_upload: null argument", cudaSuccess);
    const DevFrameCtx &fc = ctx->fc;
    if (!fc.tab.cost_mv || !fc.tab.quant4_mf[0]) return ctx_fail(ctx, "pcamv_frame_upload: pcamv_set_qp_tables has not been called", cudaSuccess);
    if (fc.subme < 1 || fc.subme > 5)
        return ctx_fail(ctx, "pcamv_frame_upload: frame analysis supports subpel_refine 1..5 (RD mode decision is raster-serial)", cudaSuccess);
    if (fc.me_method < PCAMV_ME_DIA || fc.me_method > PCAMV_ME_TESA)
        return ctx_fail(ctx, "pcamv_frame_upload: me_method must be dia/hex/umh/esa/tesa", cudaSuccess);
    if (in->pass < 0 || in->pass > 2 || in->n_ref < 1 || in->n_ref > ctx->cfg.max_refs)
        return ctx_fail(ctx, "pcamv_frame_upload: bad pass / n_ref", cudaSuccess);
    for (int i = 0; i < in->n_ref; i++)
        if (in->ref_slot[i] < 0 || in->ref_slot[i] >= ctx->cfg.max_refs + 2 || !fc.ref[in->ref_slot[i]].valid)
            return ctx_fail(ctx, "pcamv_frame_upload: reference slot was never uploaded", cudaSuccess);
    if (in->col_n_ref > 0 && (!in->col_ref8 || !in->col_mv4))
        return ctx_fail(ctx, "pcamv_frame_upload: co-located ref/mv arrays missing", cudaSuccess);
    if (in->pass == 2 && !in->device_forced && (!in->pass1 || (in->n_filp > 0 && !in->filp)))
        return ctx_fail(ctx, "pcamv_frame_upload: pass 2 needs the pass-1 records and filp[]", cudaSuccess);
    if (in->pass == 2 && in->device_forced && ctx->emb_state < 2)
        return ctx_fail(ctx, "pcamv_frame_upload: device_forced set, but pcamv_embed_stc has not built the forced decisions of this frame", cudaSuccess);
    if (ensure_frame_buffers(ctx)) return -1;

    const size_t n_mb = (size_t)fc.mb_w * fc.mb_h;
    ctx->frame_ready[in->pass] = false;
    FrameParams &fp = ctx->fp[in->pass];
    memset(&fp, 0, sizeof(fp));
    fp.pass = in->pass; fp.n_ref = in->n_ref; fp.cur_poc = in->cur_poc;
    for (int i = 0; i < PCAMV_MAX_REFS; i++)
    {
        fp.ref_slot[i] = in->ref_slot[i]; fp.ref_poc[i] = in->ref_poc[i];
        fp.col_inv_ref_poc[i] = in->col_inv_ref_poc[i];
    }
    fp.col_n_ref = in->col_n_ref;
    uint8_t *h = ctx->h_frame;
    if (in->col_n_ref > 0)
    {
        memcpy(h, in->col_ref8, 4 * n_mb);
        memcpy(h + 4 * n_mb, in->col_mv4, 64 * n_mb);
        CK(cudaMemcpyAsync(ctx->d_col_ref8, h, 4 * n_mb, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_col_mv4, h + 4 * n_mb, 64 * n_mb, cudaMemcpyHostToDevice, ctx->stream));
    }
    fp.col_ref8 = ctx->d_col_ref8; fp.col_mv4 = ctx->d_col_mv4;
    if (in->pass == 2 && in->device_forced)
        fp.forced = ctx->d_forced;                   // built in HBM by pcamv_embed_stc
    else if (in->pass == 2)
    {
        ForcedOut *fo = (ForcedOut *)(h + 68 * n_mb);
        // filp[] is consumed in cover order; build_forced walks the macroblocks in the same order
        std::vector<int8_t> flips((size_t)(in->n_filp > 0 ? in->n_filp : 0) + 16 * n_mb, 0);
        if (in->n_filp > 0) memcpy(flips.data(), in->filp, in->n_filp);
        const int used = build_forced((int)n_mb, (const Pass1Mb *)in->pass1, flips.data(), fo);
        if (used != in->n_filp)
        {
            char msg[160];
            snprintf(msg, sizeof(msg), "pcamv_frame_upload: pass-1 records carry %d motion vectors but n_filp is %d", used, in->n_filp);
            return ctx_fail(ctx, msg, cudaSuccess);
        }
        CK(cudaMemcpyAsync(ctx->d_forced, fo, n_mb * sizeof(ForcedMb), cudaMemcpyHostToDevice, ctx->stream));
        fp.forced = ctx->d_forced;
    }
    for (int i = 0; i < 16; i++)
        fp.stale_mv[i] = ((uint32_t)(uint16_t)in->stale_mv[i][0]) | ((uint32_t)(uint16_t)in->stale_mv[i][1] << 16);
    fp.cur = ctx->fa;
    fp.mvsads = ctx->d_mvsads; fp.mvsads_cap = ctx->mvsads_cap;
    fp.subparts = ctx->d_subparts;
    fp.log = ctx->d_log; fp.log_stride = ctx->log_stride; fp.results = ctx->d_mb_results; fp.row_progress = ctx->d_progress;
    // the row counters are cleared here already (the launch clears them again): whoever polls them between upload and launch
    // (pcamv_analyse_p_rows) must never see the previous pass's final values
    ctx->sr_active = false; ctx->sr_rows = 0;
    CK(cudaMemsetAsync(ctx->d_progress, 0, (2 * fc.mb_h + 2) * sizeof(int), ctx->stream));
    if (in->pass == 1) { ctx->frame_cost_table = in->cost_table != 0; ctx->emb_state = 0; }
    ctx->frame_last = in->pass;
    return 0;
}

extern "C" int pcamv_frame_upload(pcamv_ctx *ctx, const pcamv_frame_in *in)
{
    GUARD();
    if (frame_upload_async(ctx, in)) return -1;
    CK(cudaStreamSynchronize(ctx->stream));        // the caller may reuse its buffers; pinned staging is free again
    ctx->frame_ready[in->pass] = true;
    return 0;
}

static int launch_frame(pcamv_ctx *ctx, int pass, cudaEvent_t *ev = nullptr)
{
    const DevFrameCtx &fc = ctx->fc;
    ctx->fp[pass].trace = ctx->trace_on ? ctx->d_trace : nullptr;
    CK(cudaMemsetAsync(ctx->d_progress, 0, (2 * fc.mb_h + 2) * sizeof(int), ctx->stream));
    if (ev) CK(cudaEventRecord(ev[0], ctx->stream));
    launch_analyse_p(fc, ctx->fp[pass], ctx->d_progress + fc.mb_h, fc.mb_h, ctx->cfg.rows_per_cta, ctx->stream);
    ctx->launches += 1;
    if (ev) CK(cudaEventRecord(ev[1], ctx->stream));
    if (pass == 1 && ctx->frame_cost_table)
    {
        launch_cost_table(fc, ctx->fp[pass], fc.mb_w * fc.mb_h, ctx->stream);
        ctx->launches += 1;
    }
    if (ev) CK(cudaEventRecord(ev[2], ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int pcamv_frame_run(pcamv_ctx *ctx, int pass, int iters, float *ms_per_frame, float *ms_kernels)
{
    GUARD();
    if (pass < 0) pass = ctx->frame_last;
    if (iters <= 0 || pass < 0 || pass > 2 || !ctx->frame_ready[pass])
        return ctx_fail(ctx, "pcamv_frame_run: no frame uploaded for that pass", cudaSuccess);
    if (ms_kernels)
        while ((int)ctx->ev_pool.size() < 3 * iters)
        {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            ctx->ev_pool.push_back(e);
        }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; i++)
        if (launch_frame(ctx, pass, ms_kernels ? ctx->ev_pool.data() + 3 * i : nullptr)) return -1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_per_frame) *ms_per_frame = ms / iters;
    if (ms_kernels)
    {
        double a = 0, b = 0;
        for (int i = 0; i < iters; i++)
        {
            float t0 = 0, t1 = 0;
            CK(cudaEventElapsedTime(&t0, ctx->ev_pool[3 * i], ctx->ev_pool[3 * i + 1]));
            CK(cudaEventElapsedTime(&t1, ctx->ev_pool[3 * i + 1], ctx->ev_pool[3 * i + 2]));
            a += t0; b += t1;
        }
        ms_kernels[0] = (float)(a / iters); ms_kernels[1] = (float)(b / iters);
    }
    return 0;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// issues the result copies on the context's stream: straight into the caller's buffers when those are page-locked
// (cudaHostAlloc / cudaHostRegister), else into the context's pinned staging
static int frame_download_async(pcamv_ctx *ctx, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    if (ctx->frame_last < 0) return ctx_fail(ctx, "pcamv_frame_download: no frame uploaded", cudaSuccess);
    const size_t n_mb = (size_t)ctx->fc.mb_w * ctx->fc.mb_h;
    uint8_t *h = ctx->h_frame + ctx->h_frame_in_bytes;
    const size_t res_bytes = n_mb * sizeof(MbResult), log_bytes = n_mb * ctx->log_stride * sizeof(LogEntry);
    ctx->dl_mbs_direct = mbs && is_pinned(mbs);
    ctx->dl_log_direct = log && is_pinned(log);
    if (mbs) CK(cudaMemcpyAsync(ctx->dl_mbs_direct ? (void *)mbs : (void *)h, ctx->d_mb_results, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (log) CK(cudaMemcpyAsync(ctx->dl_log_direct ? (void *)log : (void *)(h + res_bytes), ctx->d_log, log_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

static int frame_download_finish(pcamv_ctx *ctx, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    const size_t n_mb = (size_t)ctx->fc.mb_w * ctx->fc.mb_h;
    uint8_t *h = ctx->h_frame + ctx->h_frame_in_bytes;
    const size_t res_bytes = n_mb * sizeof(MbResult), log_bytes = n_mb * ctx->log_stride * sizeof(LogEntry);
    CK(cudaStreamSynchronize(ctx->stream));
    if (mbs && !ctx->dl_mbs_direct) memcpy(mbs, h, res_bytes);
    if (log && !ctx->dl_log_direct) memcpy(log, h + res_bytes, log_bytes);
    return 0;
}

extern "C" int pcamv_frame_download(pcamv_ctx *ctx, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    GUARD();
    if (frame_download_async(ctx, mbs, log)) return -1;
    return frame_download_finish(ctx, mbs, log);
}

extern "C" int pcamv_analyse_p(pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    GUARD();
    if (!mbs) return ctx_fail(ctx, "pcamv_analyse_p: null output", cudaSuccess);
    if (pcamv_frame_upload(ctx, in)) return -1;
    if (launch_frame(ctx, in->pass)) return -1;
    return pcamv_frame_download(ctx, mbs, log);
}

extern "C" int pcamv_set_pass2_elide(pcamv_ctx *ctx, int on)
{
    GUARD();
    ctx->fc.pass2_elide = on != 0 && !(ctx->fc.analyse_inter & 0x20);
    return 0;
}

extern "C" int pcamv_set_conformant(pcamv_ctx *ctx, int on)
{
    GUARD();
    ctx->fc.conformant = on != 0;
    return 0;
}

extern "C" int pcamv_log_stride(const pcamv_ctx *ctx) { return ctx ? ctx->log_stride : 0; }

extern "C" int pcamv_frame_trace(pcamv_ctx *ctx, int enable, unsigned long long *out)
{
    GUARD();
    if (ensure_frame_buffers(ctx)) return -1;
    if (out)
    {
        const size_t n_mb = (size_t)ctx->fc.mb_w * ctx->fc.mb_h;
        CK(cudaMemcpyAsync(out, ctx->d_trace, 2 * n_mb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->trace_on = enable != 0;
    return 0;
}

// ---- multi-context launches: n encoder contexts (GOP shards / streams of equal geometry) analysed by ONE wavefront kernel ----
static int batch_check(pcamv_ctx *const *ctxs, int n, int pass)
{
    if (!ctxs || n <= 0 || !ctxs[0]) return -1;
    pcamv_ctx *ctx = ctxs[0];
    GUARD();
    if (pass < 0 || pass > 2) return ctx_fail(ctx, "batch: bad pass", cudaSuccess);
    for (int i = 0; i < n; i++)
    {
        pcamv_ctx *c = ctxs[i];
        if (!c || c->failed) return ctx_fail(ctx, "batch: a member context is null or failed", cudaSuccess);
        if (c->cfg.device != ctx->cfg.device || c->fc.mb_w != ctx->fc.mb_w || c->fc.mb_h != ctx->fc.mb_h ||
            c->fc.me_method != ctx->fc.me_method || c->fc.subme != ctx->fc.subme || c->fc.analyse_inter != ctx->fc.analyse_inter)
            return ctx_fail(ctx, "batch: member contexts must share device, geometry and search configuration", cudaSuccess);
        if (!c->frame_ready[pass]) return ctx_fail(ctx, "batch: a member context has no frame uploaded for that pass", cudaSuccess);
        for (int k = 0; k < i; k++)
            if (ctxs[k] == c) return ctx_fail(ctx, "batch: a context appears twice", cudaSuccess);
    }
    return 0;
}

static int launch_batch(pcamv_ctx *const *ctxs, int n, int pass, cudaEvent_t *ev)
{
    pcamv_ctx *ctx = ctxs[0];
    const DevFrameCtx &fc = ctx->fc;
    if (ctx->batch_items_cap < n)
    {
        cudaFree(ctx->d_batch); if (ctx->h_batch) cudaFreeHost(ctx->h_batch);
        ctx->d_batch = nullptr; ctx->h_batch = nullptr; ctx->batch_items_cap = 0;
        CK(cudaMalloc(&ctx->d_batch, n * sizeof(BatchItem)));
        CK(cudaMallocHost(&ctx->h_batch, n * sizeof(BatchItem)));
        ctx->batch_items_cap = n;
    }
    if (ctx->batch_claim_cap < n)
    {
        cudaFree(ctx->d_batch_claim); ctx->d_batch_claim = nullptr;
        CK(cudaMalloc(&ctx->d_batch_claim, n * sizeof(int)));
        ctx->batch_claim_cap = n;
    }
    if (!ctx->batch_max_ctas)
    {
        // persistent grid: what can be resident at once (the kernel's launch bounds ask for 24 warps per SM)
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, ctx->cfg.device));
        const int w = ctx->cfg.rows_per_cta >= 4 || ctx->cfg.rows_per_cta < 0 ? 4 : ctx->cfg.rows_per_cta >= 2 ? 2 : 1;
        int warps_per_sm = 24;
        if (const char *e = getenv("PCAMV_BATCH_WARPS_PER_SM"))        // experiments with other register budgets (PCAMV_BATCH_MIN_CTAS at build time)
            if (atoi(e) >= 4) warps_per_sm = atoi(e);
        ctx->batch_max_ctas = prop.multiProcessorCount * (warps_per_sm / w);
    }
    int cost_table = pass == 1;
    for (int i = 0; i < n; i++)
    {
        pcamv_ctx *c = ctxs[i];
        c->fp[pass].trace = c->trace_on ? c->d_trace : nullptr;
        ctx->h_batch[i].fc = c->fc;
        ctx->h_batch[i].fp = c->fp[pass];
        cost_table = cost_table && c->frame_cost_table;
        // everything a member uploaded on its own stream has been synchronised by pcamv_frame_upload
        CK(cudaMemsetAsync(c->d_progress, 0, (2 * fc.mb_h + 2) * sizeof(int), ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->d_batch, ctx->h_batch, n * sizeof(BatchItem), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_batch_claim, 0, n * sizeof(int), ctx->stream));
    // split wavefront (rows_per_cta = -2): searches and per-macroblock control code on different SMs (pcamv_split.cu); the
    // sub-8x8 partition searches are synchronous, those configurations keep the row groups
    const bool split = ctx->cfg.rows_per_cta == -2 && !(fc.analyse_inter & 0x20);
    SplitBufs sb = {};
    int split_ctrl = 0, split_rows = 0, n_sms = 0;
    if (split)
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, ctx->cfg.device));
        n_sms = prop.multiProcessorCount;
        const int total_warps = ctx->batch_max_ctas * 4;
        if (ctx->split_warps < total_warps)
        {
            cudaFree(ctx->d_split); ctx->d_split = nullptr; ctx->split_warps = 0;
            ctx->split_bytes = split_bytes(total_warps, nullptr);
            CK(cudaMalloc(&ctx->d_split, ctx->split_bytes));
            ctx->split_warps = total_warps;
        }
        size_t zero = 0;
        sb = split_carve(ctx->d_split, ctx->split_warps, &zero);
        CK(cudaMemsetAsync(ctx->d_split, 0, zero, ctx->stream));
        split_ctrl = n_sms * 2 / 9;                         // control SMs (PCAMV_SPLIT_CTRL_SMS), row slots per control team (PCAMV_SPLIT_ROWS)
        split_rows = 6;
        if (const char *e = getenv("PCAMV_SPLIT_CTRL_SMS")) if (atoi(e) > 0) split_ctrl = atoi(e);
        if (const char *e = getenv("PCAMV_SPLIT_ROWS")) if (atoi(e) > 0) split_rows = atoi(e);
        if (split_ctrl >= n_sms) split_ctrl = n_sms - 1;
        if (split_rows > SPLIT_MAX_ROWS) split_rows = SPLIT_MAX_ROWS;
    }
    if (ev) CK(cudaEventRecord(ev[0], ctx->stream));
    if (split)
        launch_analyse_p_split(ctx->d_batch, n, ctx->d_batch_claim, sb, ctx->batch_max_ctas, split_ctrl, n_sms, split_rows,
                               fc.me_method >= PCAMV_ME_ESA ? 1 : 0, ctx->stream);
    else
        launch_analyse_p_batch(ctx->d_batch, n, ctx->d_batch_claim, fc.mb_h, ctx->cfg.rows_per_cta == -2 ? 4 : ctx->cfg.rows_per_cta, ctx->batch_max_ctas, fc, ctx->stream);
    ctx->launches += 1;
    if (ev) CK(cudaEventRecord(ev[1], ctx->stream));
    if (cost_table)
    {
        launch_cost_table_batch(ctx->d_batch, n, fc.mb_w * fc.mb_h, ctx->stream);
        ctx->launches += 1;
    }
    if (ev) CK(cudaEventRecord(ev[2], ctx->stream));
    CK(cudaGetLastError());
    for (int i = 0; i < n; i++) ctxs[i]->frame_last = pass;
    return 0;
}

// after a split-wavefront launch has been synchronised: did its watchdog fire?
static int split_check(pcamv_ctx *ctx)
{
    if (ctx->cfg.rows_per_cta != -2 || !ctx->d_split) return 0;
    int flag = 0;
    CK(cudaMemcpy(&flag, ctx->d_split + SPH_ABORT * sizeof(int), sizeof(int), cudaMemcpyDeviceToHost));
    if (getenv("PCAMV_SPLIT_STATS"))
    {
        unsigned long long s[13];
        CK(cudaMemcpy(s, ctx->d_split + SPH_STATS * sizeof(int), sizeof(s), cudaMemcpyDeviceToHost));
        fprintf(stderr, "split stats: search teams %llu: wait %.1f ms serve %.1f ms each, %llu requests (%.1f us per request) | control teams %llu: "
                        "idle %.1f claim %.1f restore %.1f analyse %.1f park %.1f ms each of %.1f, %llu steps (%.2f us analysis per step), %llu in phase\n",
                s[3], s[3] ? s[0] / 1e6 / s[3] : 0.0, s[3] ? s[1] / 1e6 / s[3] : 0.0, s[2], s[2] ? s[1] / 1e3 / s[2] : 0.0,
                s[10], s[10] ? s[4] / 1e6 / s[10] : 0.0, s[10] ? s[5] / 1e6 / s[10] : 0.0, s[10] ? s[6] / 1e6 / s[10] : 0.0, s[10] ? s[7] / 1e6 / s[10] : 0.0,
                s[10] ? s[8] / 1e6 / s[10] : 0.0, s[10] ? s[11] / 1e6 / s[10] : 0.0, s[9], s[9] ? s[7] / 1e3 / s[9] : 0.0, s[12]);
    }
    if (flag) return ctx_fail(ctx, "split wavefront: a team waited 20 s for work that never came (watchdog); the launch was abandoned", cudaSuccess);
    return 0;
}

extern "C" int pcamv_frame_run_batch(pcamv_ctx *const *ctxs, int n, int pass, int iters, float *ms_per_step, float *ms_kernels)
{
    if (batch_check(ctxs, n, pass)) return -1;
    pcamv_ctx *ctx = ctxs[0];
    if (iters <= 0) return ctx_fail(ctx, "pcamv_frame_run_batch: iters must be positive", cudaSuccess);
    while ((int)ctx->ev_pool.size() < 3 * iters)
    {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->ev_pool.push_back(e);
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; i++)
        if (launch_batch(ctxs, n, pass, ctx->ev_pool.data() + 3 * i)) return -1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    if (split_check(ctx)) return -1;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_per_step) *ms_per_step = ms / iters;
    if (ms_kernels)
    {
        double a = 0, b = 0;
        for (int i = 0; i < iters; i++)
        {
            float t0 = 0, t1 = 0;
            CK(cudaEventElapsedTime(&t0, ctx->ev_pool[3 * i], ctx->ev_pool[3 * i + 1]));
            CK(cudaEventElapsedTime(&t1, ctx->ev_pool[3 * i + 1], ctx->ev_pool[3 * i + 2]));
            a += t0; b += t1;
        }
        ms_kernels[0] = (float)(a / iters); ms_kernels[1] = (float)(b / iters);
    }
    return 0;
}

// upload + launch of a multi-context pass; on return the kernel is in flight on ctxs[0]'s stream and every member's own stream
// waits for it
static int batch_begin(pcamv_ctx *const *ctxs, const pcamv_frame_in *const *ins, int n)
{
    pcamv_ctx *ctx = ctxs[0];
    // every member stages and copies on its own stream, all in flight together
    for (int i = 0; i < n; i++)
    {
        if (!ctxs[i] || !ins[i]) return ctx_fail(ctx, "pcamv_analyse_p_batch: null member", cudaSuccess);
        if (ctxs[i]->failed) return ctx_fail(ctx, "pcamv_analyse_p_batch: a member context has failed", cudaSuccess);
        if (ins[i]->pass != ins[0]->pass) return ctx_fail(ctx, "pcamv_analyse_p_batch: all frames of a launch must be in the same pass", cudaSuccess);
        if (frame_upload_async(ctxs[i], ins[i]))
            return ctxs[i] == ctx ? -1 : ctx_fail(ctx, pcamv_last_error(ctxs[i]), cudaSuccess);
    }
    for (int i = 0; i < n; i++)
    {
        CK(cudaStreamSynchronize(ctxs[i]->stream));
        ctxs[i]->frame_ready[ins[0]->pass] = true;
    }
    if (batch_check(ctxs, n, ins[0]->pass)) return -1;
    if (launch_batch(ctxs, n, ins[0]->pass, nullptr)) return -1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    for (int i = 0; i < n; i++)
        if (ctxs[i] != ctx) CK(cudaStreamWaitEvent(ctxs[i]->stream, ctx->ev1, 0));
    return 0;
}

extern "C" int pcamv_analyse_p_batch(pcamv_ctx *const *ctxs, const pcamv_frame_in *const *ins, int n,
                                     pcamv_mb_out *const *mbs, pcamv_log_entry *const *logs)
{
    if (!ctxs || !ins || !mbs || n <= 0 || !ctxs[0]) return -1;
    pcamv_ctx *ctx = ctxs[0];
    GUARD();
    for (int i = 0; i < n; i++)
        if (!mbs[i]) return ctx_fail(ctx, "pcamv_analyse_p_batch: null member", cudaSuccess);
    if (batch_begin(ctxs, ins, n)) return -1;
    for (int i = 0; i < n; i++)
        if (frame_download_async(ctxs[i], mbs[i], logs ? logs[i] : nullptr))
            return ctxs[i] == ctx ? -1 : ctx_fail(ctx, pcamv_last_error(ctxs[i]), cudaSuccess);
    for (int i = 0; i < n; i++)
        if (frame_download_finish(ctxs[i], mbs[i], logs ? logs[i] : nullptr))
            return ctxs[i] == ctx ? -1 : ctx_fail(ctx, pcamv_last_error(ctxs[i]), cudaSuccess);
    return split_check(ctx);
}

// ---- results while the wavefront is still running --------------------------------------------------------------------------
// The host that replays a P slice walks its macroblocks in raster order and needs row r only when it gets there, while the
// wavefront finishes row r at (mb_w + 2 r) / (mb_w + 2 mb_h) of its run: pcamv_analyse_p_begin returns once the kernel is in
// flight, pcamv_analyse_p_rows( row ) blocks until rows 0..row are in the caller's (page-locked) buffers.  A second stream reads
// the row counters the kernel publishes (release at GPU scope after a row's records are written, pcamv_frame_kernels.cu) and
// copies finished rows out behind it; records and log of a finished row are never written again.
static int stream_arm(pcamv_ctx *ctx, cudaStream_t kstream, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    if (!mbs || !is_pinned(mbs) || (log && !is_pinned(log)))
        return ctx_fail(ctx, "pcamv_analyse_p_begin: the result buffers must be page-locked (pcamv_host_alloc)", cudaSuccess);
    if (!ctx->side)
    {
        CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
        CK(cudaMallocHost(&ctx->h_progress, ctx->fc.mb_h * sizeof(int)));
    }
    ctx->sr_kstream = kstream; ctx->sr_mbs = mbs; ctx->sr_log = log; ctx->sr_rows = 0; ctx->sr_active = true;
    return 0;
}

extern "C" int pcamv_analyse_p_begin(pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    GUARD();
    if (!in || !mbs) return ctx_fail(ctx, "pcamv_analyse_p_begin: null argument", cudaSuccess);
    if (pcamv_frame_upload(ctx, in)) return -1;
    if (launch_frame(ctx, in->pass)) return -1;
    return stream_arm(ctx, ctx->stream, mbs, log);
}

extern "C" int pcamv_analyse_p_batch_begin(pcamv_ctx *const *ctxs, const pcamv_frame_in *const *ins, int n,
                                           pcamv_mb_out *const *mbs, pcamv_log_entry *const *logs)
{
    if (!ctxs || !ins || !mbs || n <= 0 || !ctxs[0]) return -1;
    pcamv_ctx *ctx = ctxs[0];
    GUARD();
    if (batch_begin(ctxs, ins, n)) return -1;
    for (int i = 0; i < n; i++)
        if (stream_arm(ctxs[i], ctx->stream, mbs[i], logs ? logs[i] : nullptr))
            return ctxs[i] == ctx ? -1 : ctx_fail(ctx, pcamv_last_error(ctxs[i]), cudaSuccess);
    return 0;
}

extern "C" int pcamv_analyse_p_rows(pcamv_ctx *ctx, int row, int *rows_ready)
{
    GUARD();
    const int mb_w = ctx->fc.mb_w, mb_h = ctx->fc.mb_h;
    if (!ctx->sr_active && ctx->sr_rows < mb_h)
        return ctx_fail(ctx, "pcamv_analyse_p_rows: no analysis started with pcamv_analyse_p_begin is in flight", cudaSuccess);
    if (row >= mb_h) row = mb_h - 1;
    long idle_ns = 20000;
    while (ctx->sr_rows <= row)
    {
        // the kernel's state BEFORE the look at the counters: once it has ended they are final
        const cudaError_t q = cudaStreamQuery(ctx->sr_kstream);
        if (q != cudaSuccess && q != cudaErrorNotReady) return ctx_fail(ctx, "pcamv_analyse_p_rows: the wavefront launch failed", q);
        CK(cudaMemcpyAsync(ctx->h_progress, ctx->d_progress, mb_h * sizeof(int), cudaMemcpyDeviceToHost, ctx->side));
        CK(cudaStreamSynchronize(ctx->side));
        int r = ctx->sr_rows;
        while (r < mb_h && ctx->h_progress[r] >= mb_w) r++;
        if (r > ctx->sr_rows)
        {
            const size_t first = (size_t)ctx->sr_rows * mb_w, cnt = (size_t)(r - ctx->sr_rows) * mb_w;
            CK(cudaMemcpyAsync(ctx->sr_mbs + first, ctx->d_mb_results + first, cnt * sizeof(MbResult), cudaMemcpyDeviceToHost, ctx->side));
            if (ctx->sr_log)
                CK(cudaMemcpyAsync(ctx->sr_log + first * ctx->log_stride, ctx->d_log + first * ctx->log_stride,
                                   cnt * ctx->log_stride * sizeof(LogEntry), cudaMemcpyDeviceToHost, ctx->side));
            CK(cudaStreamSynchronize(ctx->side));
            ctx->sr_rows = r;
        }
        else if (q == cudaSuccess)
            return ctx_fail(ctx, "pcamv_analyse_p_rows: the wavefront kernel ended without finishing the frame", cudaSuccess);
        else
        {
            // the next row is 0.3 - 1 ms away: back off (20 us .. 320 us) instead of burning the core — and the driver's lock —
            // that the other encoder threads of the process need
            struct timespec ts = { 0, idle_ns };
            nanosleep(&ts, nullptr);
            if (idle_ns < 320000) idle_ns *= 2;
            continue;
        }
        idle_ns = 20000;
    }
    if (ctx->sr_rows >= mb_h && ctx->sr_active)
    {
        ctx->sr_active = false;
        if (split_check(ctx)) return -1;
    }
    if (rows_ready) *rows_ready = ctx->sr_rows;
    return 0;
}

// ---- encoder groups: several encoder threads of one process (GOP shards / streams), one GPU launch per step --------------
// Every member calls pcamv_group_analyse_p when its encoder reaches a P-slice pass; the call blocks until every live
// member has submitted, the last arriver analyses all submitted frames with multi-context launches (one per distinct pass),
// and everybody returns with its own results.  Members that have no more frames call pcamv_group_leave.
#include <condition_variable>
#include <mutex>

struct pcamv_group
{
    std::mutex mu;
    std::condition_variable cv;
    int n_live = 0, n_arrived = 0;
    unsigned long long generation = 0;
    std::vector<pcamv_ctx *> ctxs;
    std::vector<const pcamv_frame_in *> ins;
    std::vector<pcamv_mb_out *> mbs;
    std::vector<pcamv_log_entry *> logs;
    std::vector<int *> rcs;
    std::vector<int> begin_only;       // member asked for pcamv_group_analyse_p_begin: launch, arm the row streaming, no download
    std::string err;
};

extern "C" int pcamv_group_create(pcamv_group **out, int n_members)
{
    if (!out || n_members <= 0) return -1;
    pcamv_group *g = new pcamv_group();
    g->n_live = n_members;
    *out = g;
    return 0;
}

extern "C" void pcamv_group_destroy(pcamv_group *g) { delete g; }

// runs with g->mu held by the last arriver
static void group_launch(pcamv_group *g)
{
    const int n = (int)g->ctxs.size();
    std::vector<char> done(n, 0);
    for (int i = 0; i < n; i++)
    {
        if (done[i]) continue;
        std::vector<pcamv_ctx *> c; std::vector<const pcamv_frame_in *> in; std::vector<pcamv_mb_out *> mb; std::vector<pcamv_log_entry *> lg;
        std::vector<int> idx;
        for (int k = i; k < n; k++)
            if (!done[k] && g->ins[k]->pass == g->ins[i]->pass && g->begin_only[k] == g->begin_only[i])
            {
                c.push_back(g->ctxs[k]); in.push_back(g->ins[k]); mb.push_back(g->mbs[k]); lg.push_back(g->logs[k]);
                idx.push_back(k); done[k] = 1;
            }
        const int rc = g->begin_only[i] ? pcamv_analyse_p_batch_begin(c.data(), in.data(), (int)c.size(), mb.data(), lg.data())
                                        : pcamv_analyse_p_batch(c.data(), in.data(), (int)c.size(), mb.data(), lg.data());
        if (rc) g->err = pcamv_last_error(c[0]);
        for (int k : idx) *g->rcs[k] = rc;
    }
    g->ctxs.clear(); g->ins.clear(); g->mbs.clear(); g->logs.clear(); g->rcs.clear(); g->begin_only.clear();
    g->n_arrived = 0;
    g->generation++;
}

static int group_submit(pcamv_group *g, pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log, int begin_only)
{
    if (!g || !ctx || !in || !mbs) return -1;
    int rc = 0;
    std::unique_lock<std::mutex> lk(g->mu);
    g->ctxs.push_back(ctx); g->ins.push_back(in); g->mbs.push_back(mbs); g->logs.push_back(log); g->rcs.push_back(&rc);
    g->begin_only.push_back(begin_only);
    g->n_arrived++;
    if (g->n_arrived >= g->n_live)
    {
        group_launch(g);
        g->cv.notify_all();
    }
    else
    {
        const unsigned long long gen = g->generation;
        g->cv.wait(lk, [&] { return g->generation != gen; });
    }
    return rc;
}

extern "C" int pcamv_group_analyse_p(pcamv_group *g, pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    return group_submit(g, ctx, in, mbs, log, 0);
}

// same rendezvous; returns when the group's launch is in flight, results arrive through pcamv_analyse_p_rows( ctx, ... )
extern "C" int pcamv_group_analyse_p_begin(pcamv_group *g, pcamv_ctx *ctx, const pcamv_frame_in *in, pcamv_mb_out *mbs, pcamv_log_entry *log)
{
    return group_submit(g, ctx, in, mbs, log, 1);
}

extern "C" int pcamv_group_leave(pcamv_group *g)
{
    if (!g) return -1;
    std::unique_lock<std::mutex> lk(g->mu);
    g->n_live--;
    if (g->n_live > 0 && g->n_arrived >= g->n_live)
    {
        // the members still waiting were only waiting for this one
        group_launch(g);
        g->cv.notify_all();
    }
    return 0;
}
