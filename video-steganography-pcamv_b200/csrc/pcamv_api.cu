// pcamv_api.cu — the C-ABI of libpcamv_cuda.so (include/pcamv.h): context, HBM layout, staging, launches.
// Host-side only; every compute step is a kernel in pcamv_kernels.cu / pcamv_frame.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "pcamv_ctx.h"

using namespace pcamv;

static thread_local std::string g_open_error;

int pcamv::ctx_fail(pcamv_ctx *c, const char *what, cudaError_t e)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "pcamv: %s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
    if (c) { c->err = buf; if (e != cudaSuccess) c->failed = true; }
    else g_open_error = buf;
    return -1;
}
static int fail(pcamv_ctx *c, const char *what, cudaError_t e) { return ctx_fail(c, what, e); }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, #call, e_); } while (0)
#define GUARD() do { if (!ctx) return -1; if (ctx->failed) return -1; cudaSetDevice(ctx->cfg.device); } while (0)   /* calls may come from any host thread */

static int align_up(int v, int a) { return (v + a - 1) / a * a; }

extern "C" int pcamv_abi_version(void) { return PCAMV_ABI_VERSION; }

extern "C" const char *pcamv_last_error(const pcamv_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_open_error.c_str();
}

extern "C" void *pcamv_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    memset(p, 0, bytes);
    return p;
}
extern "C" void pcamv_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" long long pcamv_launch_count(const pcamv_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int ensure_stage(pcamv_ctx *ctx, size_t bytes)
{
    if (ctx->h_stage_bytes >= bytes) return 0;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr; ctx->h_stage_bytes = 0;
    CK(cudaMallocHost(&ctx->h_stage, bytes));
    ctx->h_stage_bytes = bytes;
    return 0;
}

extern "C" int pcamv_open(pcamv_ctx **out, const pcamv_cfg *cfg)
{
    pcamv_ctx *ctx = nullptr;
    if (!out || !cfg) return fail(nullptr, "pcamv_open: null argument", cudaSuccess);
    *out = nullptr;
    if (cfg->abi_version != PCAMV_ABI_VERSION) return fail(nullptr, "pcamv_open: ABI version mismatch", cudaSuccess);
    if (cfg->width <= 0 || cfg->height <= 0 || (cfg->width & 15) || (cfg->height & 15))
        return fail(nullptr, "pcamv_open: width/height must be positive multiples of 16", cudaSuccess);
    if (cfg->max_refs < 1 || cfg->max_refs > PCAMV_MAX_REFS)
        return fail(nullptr, "pcamv_open: max_refs out of range", cudaSuccess);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, "pcamv_open: no CUDA device (this library has no CPU fallback)", e == cudaSuccess ? cudaErrorNoDevice : e);
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "pcamv_open: bad device ordinal", cudaSuccess);
    // PCAMV_BLOCKING_SYNC=1: host threads waiting for the GPU sleep instead of spinning — for processes with more encoder threads
    // than cores to spare (the bound host sets it for --shards jobs that oversubscribe the machine).  It is a property of the
    // device's primary context and must be chosen before that exists: first pcamv_open of the process only.
    {
        static bool chosen = false;
        if (!chosen)
        {
            chosen = true;
            const char *bs = getenv("PCAMV_BLOCKING_SYNC");
            if (bs && atoi(bs) && cudaInitDevice(cfg->device, cudaDeviceScheduleBlockingSync, cudaInitDeviceFlagsAreValid) != cudaSuccess)
                cudaGetLastError();             // (context already there with other flags: carry on with those)
        }
    }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, "cudaSetDevice", e);

    ctx = new pcamv_ctx();
    ctx->cfg = *cfg;
    DevFrameCtx &fc = ctx->fc;
    memset(&fc, 0, sizeof(fc));
    fc.width = cfg->width; fc.height = cfg->height; fc.mb_w = cfg->width / 16; fc.mb_h = cfg->height / 16;
    // same plane geometry as the host encoder (reference common/frame.c:46-60, align = 16 without asm)
    fc.stride_y = align_up(cfg->width + 2 * PCAMV_PADH, 16);
    fc.stride_c = align_up(fc.stride_y / 2, 16);
    fc.me_method = cfg->me_method; fc.me_range = cfg->me_range; fc.subme = cfg->subpel_refine;
    if (fc.me_method == PCAMV_ME_TESA && fc.subme <= 1)
        fc.me_method = PCAMV_ME_ESA;          // as the reference does (encoder/encoder.c:490-492)
    fc.chroma_me = cfg->chroma_me; fc.mv_range = cfg->mv_range; fc.max_refs = cfg->max_refs;
    fc.b_cabac = cfg->b_cabac; fc.b_fast_pskip = cfg->b_fast_pskip; fc.b_dct_decimate = cfg->b_dct_decimate;
    fc.analyse_inter = cfg->analyse_inter;
    // (not with sub-8x8 partitions: a forced P_8x8 keeps the partition its own pass-2 analysis decided, which the host reads)
    fc.pass2_elide = cfg->pass2_elide != 0 && !(cfg->analyse_inter & 0x20);
    {
        // most entries one macroblock can log: a 16x16 search per reference (twice in pass 2 when an early skip is
        // overridden), four 8x8, two 16x8 + two 8x16 per candidate reference (<= 2 each), two refinements, two cost-table
        // entries; rounded up to a multiple of 4 and capped by the ABI constant
        const int r = cfg->max_refs, r2 = r < 2 ? r : 2;
        int n = 2 * r + 4 + 4 * r2 + 2 + 2;
        if (cfg->analyse_inter & 0x20)
            n += 32 + 14 + 14;      // X264_ANALYSE_PSUB8x8: 8 sub-block searches per 8x8 block, up to 16 refinements and 16 cost-table entries
        n = (n + 3) & ~3;
        ctx->log_stride = n < PCAMV_LOG_MAX ? n : PCAMV_LOG_MAX;
    }
    ctx->luma_bytes = (size_t)fc.stride_y * (cfg->height + 2 * PCAMV_PADV);
    ctx->chroma_bytes = (size_t)fc.stride_c * (cfg->height / 2 + PCAMV_PADV);
    ctx->ref_bytes = 4 * ctx->luma_bytes + 2 * ctx->chroma_bytes;

#define OCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(nullptr, #call, e_); pcamv_close(ctx); return -1; } } while (0)
    OCK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    OCK(cudaEventCreate(&ctx->ev0));
    OCK(cudaEventCreate(&ctx->ev1));
    {
        const size_t fenc_bytes = (size_t)fc.stride_y * cfg->height + 2 * (size_t)fc.stride_c * (cfg->height / 2) + 256;
        OCK(cudaMalloc(&ctx->d_fenc, fenc_bytes));
        OCK(cudaMemsetAsync(ctx->d_fenc, 0, fenc_bytes, ctx->stream));
        fc.fenc_y = ctx->d_fenc;
        fc.fenc_u = ctx->d_fenc + (size_t)fc.stride_y * cfg->height;
        fc.fenc_v = fc.fenc_u + (size_t)fc.stride_c * (cfg->height / 2);
    }
    for (int s = 0; s < cfg->max_refs + 2; s++)
    {
        OCK(cudaMalloc(&ctx->d_ref[s], ctx->ref_bytes + 4096));
        OCK(cudaMemsetAsync(ctx->d_ref[s], 0, ctx->ref_bytes + 4096, ctx->stream));
        DevRef &r = fc.ref[s];
        for (int k = 0; k < 4; k++)
            r.y[k] = ctx->d_ref[s] + k * ctx->luma_bytes + (size_t)fc.stride_y * PCAMV_PADV + PCAMV_PADH;
        r.u = ctx->d_ref[s] + 4 * ctx->luma_bytes + (size_t)fc.stride_c * (PCAMV_PADV / 2) + PCAMV_PADH / 2;
        r.v = r.u + ctx->chroma_bytes;
        r.integral = nullptr; r.integral4 = nullptr; r.poc = -1; r.valid = 0;
        r.base = ctx->d_ref[s]; r.bytes = ctx->ref_bytes + 4096;
        if (fc.me_method >= PCAMV_ME_ESA)
        {
            OCK(cudaMalloc(&ctx->d_integral[s], ctx->luma_bytes * sizeof(uint16_t)));
            OCK(cudaMemsetAsync(ctx->d_integral[s], 0, ctx->luma_bytes * sizeof(uint16_t), ctx->stream));
            r.integral = ctx->d_integral[s] + (size_t)fc.stride_y * PCAMV_PADV + PCAMV_PADH;
            if (fc.analyse_inter & 0x20)
            {
                OCK(cudaMalloc(&ctx->d_integral4[s], ctx->luma_bytes * sizeof(uint16_t)));
                OCK(cudaMemsetAsync(ctx->d_integral4[s], 0, ctx->luma_bytes * sizeof(uint16_t), ctx->stream));
                r.integral4 = ctx->d_integral4[s] + (size_t)fc.stride_y * PCAMV_PADV + PCAMV_PADH;
            }
        }
    }
    OCK(cudaMalloc(&ctx->d_cost_mv, 32769 * sizeof(int16_t)));
    OCK(cudaMalloc(&ctx->d_tables, 4096));
    OCK(cudaStreamSynchronize(ctx->stream));
#undef OCK
    *out = ctx;
    return 0;
}

extern "C" void pcamv_close(pcamv_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_fenc);
    for (int s = 0; s < PCAMV_SLOTS; s++) { cudaFree(ctx->d_ref[s]); cudaFree(ctx->d_integral[s]); cudaFree(ctx->d_integral4[s]); }
    cudaFree(ctx->d_cost_mv); cudaFree(ctx->d_tables);
    cudaFree(ctx->d_calls); cudaFree(ctx->d_results);
    cudaFree(ctx->fa.type); cudaFree(ctx->fa.ref8); cudaFree(ctx->fa.mv4); cudaFree(ctx->fa.mvr);
    cudaFree(ctx->d_col_ref8); cudaFree(ctx->d_col_mv4); cudaFree(ctx->d_forced); cudaFree(ctx->d_log);
    cudaFree(ctx->d_mb_results); cudaFree(ctx->d_progress); cudaFree(ctx->d_trace); cudaFree(ctx->d_mvsads); cudaFree(ctx->d_seam_mvsads); cudaFree(ctx->d_stc); cudaFree(ctx->d_subparts); cudaFree(ctx->d_batch); cudaFree(ctx->d_batch_claim); cudaFree(ctx->d_split); cudaFree(ctx->d_emb); cudaFree(ctx->d_stc_io); cudaFree(ctx->d_recon_nnz); cudaFree(ctx->d_recon_patches);
    if (ctx->h_progress) cudaFreeHost(ctx->h_progress);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->h_batch) cudaFreeHost(ctx->h_batch);
    if (ctx->h_frame) cudaFreeHost(ctx->h_frame);
    if (ctx->h_calls) cudaFreeHost(ctx->h_calls);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int pcamv_set_qp_tables(pcamv_ctx *ctx, const pcamv_qp_tables *t)
{
    GUARD();
    if (!t || !t->cost_mv) return fail(ctx, "pcamv_set_qp_tables: null table", cudaSuccess);
    // stage everything in one pinned block: cost_mv | cost_ref | mf0 mf1 | bias0 bias1 | dq0 dq1
    const size_t need = 65540 + 3 * 33 * 2 + 4 + 4 * 16 * 2 + 2 * 96 * 4;
    if (ensure_stage(ctx, need + 64)) return -1;
    uint8_t *h = ctx->h_stage;
    memcpy(h, t->cost_mv, 32769 * 2);
    uint8_t *ht = h + 65540;        // 4-byte aligned table block behind the 32769 int16 entries
    size_t o = 0;
    auto put = [&](const void *src, size_t n) { if (src) memcpy(ht + o, src, n); else memset(ht + o, 0, n); size_t at = o; o += n; return at; };
    const size_t o_ref = put(t->cost_ref, 3 * 33 * 2);
    o = (o + 3) & ~(size_t)3;
    const size_t o_mf0 = put(t->quant4_mf[0], 32), o_mf1 = put(t->quant4_mf[1], 32);
    const size_t o_b0 = put(t->quant4_bias[0], 32), o_b1 = put(t->quant4_bias[1], 32);
    const size_t o_dq0 = put(t->dequant4_mf[0], 96 * 4), o_dq1 = put(t->dequant4_mf[1], 96 * 4);
    CK(cudaMemcpyAsync(ctx->d_cost_mv, h, 32769 * 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tables, ht, o, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    DevTables &d = ctx->fc.tab;
    d.cost_mv = ctx->d_cost_mv + 16384;
    d.cost_ref = (const uint16_t *)(ctx->d_tables + o_ref);
    d.quant4_mf[0] = (const uint16_t *)(ctx->d_tables + o_mf0); d.quant4_mf[1] = (const uint16_t *)(ctx->d_tables + o_mf1);
    d.quant4_bias[0] = (const uint16_t *)(ctx->d_tables + o_b0); d.quant4_bias[1] = (const uint16_t *)(ctx->d_tables + o_b1);
    d.dequant4_mf[0] = (const int32_t *)(ctx->d_tables + o_dq0); d.dequant4_mf[1] = (const int32_t *)(ctx->d_tables + o_dq1);
    d.qp = t->qp; d.lambda = t->lambda; d.lambda2_chroma = t->lambda2_chroma; d.chroma_qp = t->chroma_qp;
    return 0;
}

extern "C" int pcamv_put_fenc(pcamv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int stride_y, int stride_c)
{
    GUARD();
    if (!y || !u || !v) return fail(ctx, "pcamv_put_fenc: null plane", cudaSuccess);
    const DevFrameCtx &fc = ctx->fc;
    CK(cudaMemcpy2DAsync((void *)fc.fenc_y, fc.stride_y, y, stride_y, fc.width, fc.height, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync((void *)fc.fenc_u, fc.stride_c, u, stride_c, fc.width / 2, fc.height / 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync((void *)fc.fenc_v, fc.stride_c, v, stride_c, fc.width / 2, fc.height / 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));      // the caller may reuse its buffers
    return 0;
}

// build borders + half-pel planes of a slot whose integer luma / chroma interiors are already in HBM
int pcamv::filter_slot(pcamv_ctx *ctx, int slot)
{
    const DevFrameCtx &fc = ctx->fc;
    DevRef &r = ctx->fc.ref[slot];
    const int W = fc.width, H = fc.height;
    launch_expand_border(r.y[0], nullptr, nullptr, 1, fc.stride_y, 0, W, 0, H, -PCAMV_PADH, W + PCAMV_PADH, -PCAMV_PADV, H + PCAMV_PADV, ctx->stream);
    launch_expand_border(r.u, r.v, nullptr, 2, fc.stride_c, 0, W / 2, 0, H / 2, -PCAMV_PADH / 2, W / 2 + PCAMV_PADH / 2,
                         -PCAMV_PADV / 2, H / 2 + PCAMV_PADV / 2, ctx->stream);
    if (r.integral)
    {
        const size_t pad = (size_t)fc.stride_y * PCAMV_PADV + PCAMV_PADH;
        launch_box_sum(r.y[0] - pad, r.integral - pad, fc.stride_y, H + 2 * PCAMV_PADV, 8, ctx->stream);
        ctx->launches += 1;
        if (r.integral4)
        {
            launch_box_sum(r.y[0] - pad, r.integral4 - pad, fc.stride_y, H + 2 * PCAMV_PADV, 4, ctx->stream);
            ctx->launches += 1;
        }
    }
    launch_hpel_filter(r.y[0], r.y[1], r.y[2], r.y[3], fc.stride_y, W, H, ctx->stream);
    launch_expand_border(r.y[1], r.y[2], r.y[3], 3, fc.stride_y, -4, W + 4, -8, H + 8, -PCAMV_PADH, W + PCAMV_PADH,
                         -PCAMV_PADV, H + PCAMV_PADV, ctx->stream);
    ctx->launches += 4;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int pcamv_put_ref(pcamv_ctx *ctx, int slot, int poc, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                             int stride_y, int stride_c)
{
    GUARD();
    if (slot < 0 || slot >= ctx->cfg.max_refs + 2) return fail(ctx, "pcamv_put_ref: bad slot", cudaSuccess);
    if (!y || !u || !v) return fail(ctx, "pcamv_put_ref: null plane", cudaSuccess);
    const DevFrameCtx &fc = ctx->fc;
    DevRef &r = ctx->fc.ref[slot];
    CK(cudaMemcpy2DAsync(r.y[0], fc.stride_y, y, stride_y, fc.width, fc.height, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(r.u, fc.stride_c, u, stride_c, fc.width / 2, fc.height / 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(r.v, fc.stride_c, v, stride_c, fc.width / 2, fc.height / 2, cudaMemcpyHostToDevice, ctx->stream));
    if (filter_slot(ctx, slot)) return -1;
    CK(cudaStreamSynchronize(ctx->stream));
    r.poc = poc; r.valid = 1;
    return 0;
}

extern "C" int pcamv_put_ref_planes(pcamv_ctx *ctx, int slot, int poc, const uint8_t *const luma_padded[4],
                                    const uint8_t *u_padded, const uint8_t *v_padded)
{
    GUARD();
    if (slot < 0 || slot >= ctx->cfg.max_refs + 2) return fail(ctx, "pcamv_put_ref_planes: bad slot", cudaSuccess);
    if (!luma_padded || !u_padded || !v_padded) return fail(ctx, "pcamv_put_ref_planes: null plane", cudaSuccess);
    DevRef &r = ctx->fc.ref[slot];
    uint8_t *base = ctx->d_ref[slot];
    for (int k = 0; k < 4; k++)
        CK(cudaMemcpyAsync(base + k * ctx->luma_bytes, luma_padded[k], ctx->luma_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + 4 * ctx->luma_bytes, u_padded, ctx->chroma_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + 4 * ctx->luma_bytes + ctx->chroma_bytes, v_padded, ctx->chroma_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (r.integral)
    {
        launch_box_sum(base, ctx->d_integral[slot], ctx->fc.stride_y, ctx->fc.height + 2 * PCAMV_PADV, 8, ctx->stream);
        ctx->launches += 1;
        if (r.integral4)
        {
            launch_box_sum(base, ctx->d_integral4[slot], ctx->fc.stride_y, ctx->fc.height + 2 * PCAMV_PADV, 4, ctx->stream);
            ctx->launches += 1;
        }
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    r.poc = poc; r.valid = 1;
    return 0;
}

extern "C" size_t pcamv_plane_bytes(const pcamv_ctx *ctx, int plane)
{
    if (!ctx || plane < 0 || plane > 6) return 0;
    if (plane == 6) return ctx->fc.me_method >= PCAMV_ME_ESA ? ctx->luma_bytes * sizeof(uint16_t) : 0;
    return plane < 4 ? ctx->luma_bytes : ctx->chroma_bytes;
}

extern "C" int pcamv_plane_stride(const pcamv_ctx *ctx, int plane)
{
    if (!ctx || plane < 0 || plane > 6) return 0;
    if (plane == 6) return ctx->fc.stride_y * (int)sizeof(uint16_t);
    return plane < 4 ? ctx->fc.stride_y : ctx->fc.stride_c;
}

extern "C" int pcamv_get_ref_plane(pcamv_ctx *ctx, int slot, int plane, uint8_t *dst)
{
    GUARD();
    if (slot < 0 || slot >= ctx->cfg.max_refs + 2 || plane < 0 || plane > 6 || !dst || (plane == 6 && !ctx->d_integral[slot]))
        return fail(ctx, "pcamv_get_ref_plane: bad argument", cudaSuccess);
    const uint8_t *src = plane == 6 ? (const uint8_t *)ctx->d_integral[slot] : ctx->d_ref[slot] + (plane < 4 ? plane * ctx->luma_bytes : 4 * ctx->luma_bytes + (plane - 4) * ctx->chroma_bytes);
    CK(cudaMemcpyAsync(dst, src, pcamv_plane_bytes(ctx, plane), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int ensure_batch(pcamv_ctx *ctx, int n)
{
    if (n <= ctx->batch_cap) return 0;
    cudaFree(ctx->d_calls); cudaFree(ctx->d_results);
    if (ctx->h_calls) cudaFreeHost(ctx->h_calls);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    ctx->d_calls = nullptr; ctx->d_results = nullptr; ctx->h_calls = nullptr; ctx->h_results = nullptr; ctx->batch_cap = 0;
    const int cap = n + n / 4 + 1024;
    CK(cudaMalloc(&ctx->d_calls, (size_t)cap * sizeof(pcamv_me_call)));
    CK(cudaMalloc(&ctx->d_results, (size_t)cap * sizeof(pcamv_me_result)));
    CK(cudaMallocHost(&ctx->h_calls, (size_t)cap * sizeof(pcamv_me_call)));
    CK(cudaMallocHost(&ctx->h_results, (size_t)cap * sizeof(pcamv_me_result)));
    ctx->batch_cap = cap;
    return 0;
}

// --me tesa through the stateless seam: calls are launched PCAMV_SEAM_CHUNK at a time, each with its own candidate list
#define PCAMV_SEAM_CHUNK 2048
static int seam_cap(const pcamv_ctx *ctx) { return (2 * ctx->fc.me_range + 4) * (2 * ctx->fc.me_range + 1); }
static void run_search_batch(pcamv_ctx *ctx, int n)
{
    launch_search_batch(ctx->fc, ctx->d_calls, n, ctx->d_results, ctx->d_seam_mvsads, seam_cap(ctx), PCAMV_SEAM_CHUNK, ctx->stream);
}

static int check_calls(pcamv_ctx *ctx, const pcamv_me_call *calls, int n)
{
    if (!ctx->fc.tab.cost_mv) return fail(ctx, "search: pcamv_set_qp_tables has not been called", cudaSuccess);
    if (ctx->fc.me_method == PCAMV_ME_TESA && !ctx->d_seam_mvsads)
    {
        cudaError_t e = cudaMalloc(&ctx->d_seam_mvsads, (size_t)PCAMV_SEAM_CHUNK * seam_cap(ctx) * sizeof(unsigned long long));
        if (e != cudaSuccess) return fail(ctx, "search: cudaMalloc of the --me tesa candidate lists", e);
    }
    for (int i = 0; i < n; i++)
    {
        const pcamv_me_call &c = calls[i];
        if (c.ref_slot < 0 || c.ref_slot >= ctx->cfg.max_refs + 2 || !ctx->fc.ref[c.ref_slot].valid)
            return fail(ctx, "search: call names a reference slot that was never uploaded", cudaSuccess);
        if (c.mb_x < 0 || c.mb_x >= ctx->fc.mb_w || c.mb_y < 0 || c.mb_y >= ctx->fc.mb_h ||
            c.i_pixel < 0 || c.i_pixel > PCAMV_PIXEL_4x4 || c.i_mvc < 0 || c.i_mvc > PCAMV_MAX_MVC ||
            (c.xoff & 3) || (c.yoff & 3) || c.xoff < 0 || c.yoff < 0 || c.xoff > 12 || c.yoff > 12)
            return fail(ctx, "search: malformed call record", cudaSuccess);
    }
    return 0;
}

extern "C" int pcamv_me_batch_upload(pcamv_ctx *ctx, const pcamv_me_call *calls, int n)
{
    GUARD();
    if (n < 0 || (n && !calls)) return fail(ctx, "pcamv_me_batch_upload: bad argument", cudaSuccess);
    if (check_calls(ctx, calls, n)) return -1;
    if (ensure_batch(ctx, n)) return -1;
    memcpy(ctx->h_calls, calls, (size_t)n * sizeof(pcamv_me_call));
    CK(cudaMemcpyAsync(ctx->d_calls, ctx->h_calls, (size_t)n * sizeof(pcamv_me_call), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->batch_n = n;
    return 0;
}

extern "C" int pcamv_me_batch_run(pcamv_ctx *ctx, int iters, float *ms_per_launch)
{
    GUARD();
    if (iters <= 0 || ctx->batch_n <= 0) return fail(ctx, "pcamv_me_batch_run: nothing to run", cudaSuccess);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; i++)
        run_search_batch(ctx, ctx->batch_n);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(ctx->ev1));
    ctx->launches += iters;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_per_launch) *ms_per_launch = ms / iters;
    return 0;
}

extern "C" int pcamv_me_batch_download(pcamv_ctx *ctx, pcamv_me_result *results, int n)
{
    GUARD();
    if (n < 0 || n > ctx->batch_n || (n && !results)) return fail(ctx, "pcamv_me_batch_download: bad argument", cudaSuccess);
    CK(cudaMemcpyAsync(ctx->h_results, ctx->d_results, (size_t)n * sizeof(pcamv_me_result), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(results, ctx->h_results, (size_t)n * sizeof(pcamv_me_result));
    return 0;
}

extern "C" int pcamv_me_search_batch(pcamv_ctx *ctx, const pcamv_me_call *calls, int n, pcamv_me_result *results)
{
    GUARD();
    if (n < 0 || (n && (!calls || !results))) return fail(ctx, "pcamv_me_search_batch: bad argument", cudaSuccess);
    if (n == 0) return 0;
    if (pcamv_me_batch_upload(ctx, calls, n)) return -1;
    run_search_batch(ctx, n);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return pcamv_me_batch_download(ctx, results, n);
}

// Integer-pipe issue peak, in giga lane-operations per second (32 lanes x instructions retired).
extern "C" int pcamv_int_peak(pcamv_ctx *ctx, double *gops)
{
    GUARD();
    if (!gops) return fail(ctx, "pcamv_int_peak: null argument", cudaSuccess);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
    const int blocks = sms * 8, iters = 4096;
    uint32_t *d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * 256 * 4));
    launch_int_peak(d, blocks, 64, ctx->stream);                    // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        launch_int_peak(d, blocks, iters, ctx->stream);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (ms < best) best = ms;
    }
    ctx->launches += 6;
    cudaFree(d);
    // per thread and iteration: 8 chains x 5 integer instructions (VABSDIFF4.ACC, IADD3 x2, LOP3, SHF)
    const double ops = (double)blocks * 256 * iters * 8 * 5;
    *gops = ops / (best * 1e-3) / 1e9;
    return 0;
}
