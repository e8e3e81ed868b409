// pcamv_embed.cu — the embed stage between the two passes of a P frame, on the device.
//
// Behavioural contract (bit-exact), reference encoder/encoder.c:
//   :1561-1647  cover bit LSB(mvx + mvy) and rho_final = (float)inter_stego_cost of every motion vector of every carrier
//               macroblock (P_L0, P_8x8), in macroblock raster order, partition order inside a macroblock
//   :1650-1819  MVC penalties: the two vectors of a 16x8 / 8x16 / 8x4 / 4x8 split whose components differ by less than 2 in
//               total get rho * 2; the four vectors of an all-8x8 P_8x8 macroblock or of a 4x4 split get
//               rho * (0.7f * n + 1) with n = how many of the eight component differences around the ring are 0 or 1.
//               (alpha_loc = 1, alpha_com = 0: the "complexity" term of the absent S-UNIWARD library drops out.)
//   :1848-1855  filp[i] = cover[i] ^ stego[i]
//   encoder/analyse.c:3526-3689  what h->info.cache[mb] holds after pass 1, including the motion vectors as the
//               unsequenced `idx++` copy loops leave them (entry k is the vector of 4x4 block
//               B[k] = 0 0 1 3 3 4 6 6 7 9 9 10 12 12 13 15 with gcc) — cover bits, penalties and the vectors pass 2 forces
//               all read THAT array, so it is rebuilt here exactly
//   encoder/analyse.c:2870-3107  pass 2 forces type / partition / references / vectors, flipped where filp says so
//               (build_forced_mb, pcamv_glue.h: the same function the host path runs)
// The float arithmetic is two single-precision operations per penalty, each rounded on its own (__fmul_rn / __fadd_rn: no
// contraction), as the reference's x86-64 SSE code rounds them.
//
// Everything stays in HBM between pcamv_analyse_p(pass 1) and pcamv_analyse_p(pass 2): per-macroblock records, cover, rho,
// stego, flips, forced decisions.  What crosses PCIe: the cover length (4 bytes up), the message bits (down), and whatever
// the host asks to see.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "pcamv_ctx.h"
#include "pcamv_glue.h"

namespace pcamv {

enum { EPIX_8x8 = 3, EPIX_8x4 = 4, EPIX_4x8 = 5 };      // i_pixel values of PartInfo (PIX_* of pcamv_me.cuh / reference common/pixel.h:30-42)

__device__ __forceinline__ int emb_carriers(const MbResult &r) { return r.type == MB_P_SKIP ? 0 : r.n_part; }

// offsets[mb] = carriers of the macroblocks before mb (raster order); offsets[n_mb] = cover length.  One CTA.
__global__ void __launch_bounds__(1024) k_embed_scan(const MbResult *__restrict__ res, int n_mb, int *__restrict__ offsets)
{
    __shared__ int s_sum[1024];
    const int t = threadIdx.x;
    const int per = (n_mb + 1023) / 1024;
    const int lo = min(t * per, n_mb), hi = min(lo + per, n_mb);
    int sum = 0;
    for (int mb = lo; mb < hi; mb++) sum += emb_carriers(res[mb]);
    s_sum[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1)             // inclusive scan of the chunk sums
    {
        const int v = t >= d ? s_sum[t - d] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int run = s_sum[t] - sum;
    for (int mb = lo; mb < hi; mb++) { offsets[mb] = run; run += emb_carriers(res[mb]); }
    if (t == 1023) offsets[n_mb] = s_sum[1023];
}

__device__ __forceinline__ int emb_mvx(uint32_t p) { return (int)(int16_t)(p & 0xffffu); }
__device__ __forceinline__ int emb_mvy(uint32_t p) { return (int)(int16_t)(p >> 16); }
__device__ __forceinline__ int emb_is01(int d) { return d == 0 || d == 1; }
__device__ __forceinline__ int emb_iabs(int v) { return v < 0 ? -v : v; }

// one thread per macroblock: info.cache[mb] as the reference leaves it after pass 1, cover bits and rho of its carriers
__global__ void k_embed_fill(const MbResult *__restrict__ res, const LogEntry *__restrict__ log, int log_stride,
                             const PartInfo *__restrict__ subparts, const int *__restrict__ offsets, int n_mb, int straight_copy,
                             Pass1Mb *__restrict__ p1, uint8_t *__restrict__ cover, float *__restrict__ rho)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= n_mb) return;
    const MbResult r = res[mb];
    Pass1Mb p;
    memset(&p, 0, sizeof(p));
    p.type = r.type; p.partition = r.partition; p.used = r.type != MB_P_SKIP;
    if (!p.used)
    {
        p1[mb] = p;
        return;
    }
    // references of the 4x4 blocks in RASTER order; vectors through the reference's scrambled copy
    for (int row = 0; row < 4; row++)
        for (int x = 0; x < 4; x++) p.ref[4 * row + x] = r.ref[(row >> 1) * 2 + (x >> 1)];
    const int B[16] = { 0, 0, 1, 3, 3, 4, 6, 6, 7, 9, 9, 10, 12, 12, 13, 15 };
    uint32_t mvs[16];
    for (int k = 0; k < 16; k++)
    {
        mvs[k] = r.mv[straight_copy ? k : B[k]];          // pcamv_set_conformant: entry k is the vector of block k
        p.mv[k][0] = (int16_t)emb_mvx(mvs[k]); p.mv[k][1] = (int16_t)emb_mvy(mvs[k]);
    }
    // carriers in the reference's order: slot in info.cache (mv / mv_stego / inter_stego_cost index), split kind per 8x8 block
    int slots[16], n = 0;
    const int n_part = r.n_part;
    if (r.type == MB_P_L0)
    {
        if (r.partition == PART_16x16) slots[n++] = 0;
        else if (r.partition == PART_8x16) { slots[n++] = 0; slots[n++] = 4; }
        else { slots[n++] = 0; slots[n++] = 8; }
    }
    else
    {
        // P_8x8: the MV-carrying blocks are in the side array (up to 16), in the cost table's order: 8x8 block by 8x8 block
        const PartInfo *sp = subparts ? subparts + (size_t)16 * mb : r.part;
        for (int i8 = 0; i8 < 4; i8++) p.sub[i8] = 3;          // D_L0_8x8 unless the parts say otherwise
        for (int k = 0; k < n_part; k++)
        {
            const PartInfo pi = sp[k];
            const int i8 = (pi.xoff >> 3) + 2 * (pi.yoff >> 3), sx = (pi.xoff & 7) >> 2, sy = (pi.yoff & 7) >> 2;
            const int kind = pi.i_pixel == EPIX_8x8 ? 3 : pi.i_pixel == EPIX_8x4 ? 1 : pi.i_pixel == EPIX_4x8 ? 2 : 0;
            p.sub[i8] = (uint8_t)kind;
            slots[n++] = 4 * i8 + (kind == 3 ? 0 : kind == 1 ? 2 * sy : kind == 2 ? sx : sx + 2 * sy);
        }
    }
    // cost-table entries: the last n_part log entries of the macroblock, in partition order
    const LogEntry *ih = log + (size_t)mb * log_stride + (r.n_log - n_part);
    const int off = offsets[mb];
    float rh[16];
    for (int j = 0; j < n; j++)
    {
        const int s = slots[j];
        const PartInfo pi = (r.type == MB_P_8x8 && subparts) ? subparts[(size_t)16 * mb + j] : r.part[j];
        p.mv_stego[s][0] = (int16_t)(pi.mv[0] + ih[j].mv[0]);
        p.mv_stego[s][1] = (int16_t)(pi.mv[1] + ih[j].mv[1]);
        cover[off + j] = (uint8_t)((p.mv[s][0] + p.mv[s][1]) & 1);
        rh[j] = (float)ih[j].cost;
    }
    // MVC penalties (every group of carriers is contiguous in cover order)
    const float c1 = 2.0f, c2 = 0.7f;
    #define EMB_D(a, b, comp) emb_iabs(p.mv[a][comp] - p.mv[b][comp])
    if (r.type == MB_P_L0)
    {
        if (r.partition != PART_16x16)
        {
            const int b = r.partition == PART_8x16 ? 4 : 8;
            if (EMB_D(0, b, 0) + EMB_D(0, b, 1) < 2) { rh[0] = __fmul_rn(rh[0], c1); rh[1] = __fmul_rn(rh[1], c1); }
        }
    }
    else
    {
        if (p.sub[0] == 3 && p.sub[1] == 3 && p.sub[2] == 3 && p.sub[3] == 3)
        {
            const int cnt = emb_is01(EMB_D(0, 4, 0)) + emb_is01(EMB_D(4, 12, 0)) + emb_is01(EMB_D(12, 8, 0)) + emb_is01(EMB_D(8, 0, 0)) +
                            emb_is01(EMB_D(0, 4, 1)) + emb_is01(EMB_D(4, 12, 1)) + emb_is01(EMB_D(12, 8, 1)) + emb_is01(EMB_D(8, 0, 1));
            const float f = __fadd_rn(__fmul_rn(c2, (float)cnt), 1.0f);
            for (int j = 0; j < 4; j++) rh[j] = __fmul_rn(rh[j], f);
        }
        int j = 0;
        for (int i = 0; i < 4; i++)
        {
            const int kind = p.sub[i];
            if (kind == 3) { j += 1; continue; }
            if (kind == 2 || kind == 1)
            {
                const int b = kind == 2 ? 4 * i + 1 : 4 * i + 2;
                if (EMB_D(4 * i, b, 0) + EMB_D(4 * i, b, 1) < 2) { rh[j] = __fmul_rn(rh[j], c1); rh[j + 1] = __fmul_rn(rh[j + 1], c1); }
                j += 2;
                continue;
            }
            const int a = 4 * i;
            const int cnt = emb_is01(EMB_D(a, a + 1, 0)) + emb_is01(EMB_D(a + 1, a + 3, 0)) + emb_is01(EMB_D(a + 2, a + 3, 0)) + emb_is01(EMB_D(a, a + 2, 0)) +
                            emb_is01(EMB_D(a, a + 1, 1)) + emb_is01(EMB_D(a + 1, a + 3, 1)) + emb_is01(EMB_D(a + 2, a + 3, 1)) + emb_is01(EMB_D(a, a + 2, 1));
            const float f = __fadd_rn(__fmul_rn(c2, (float)cnt), 1.0f);
            for (int q = 0; q < 4; q++) rh[j + q] = __fmul_rn(rh[j + q], f);
            j += 4;
        }
    }
    #undef EMB_D
    for (int j = 0; j < n; j++) rho[off + j] = rh[j];
    p1[mb] = p;
}

// one thread per macroblock: flips of its carriers, then the decision pass 2 forces
__global__ void k_embed_forced(const Pass1Mb *__restrict__ p1, const int *__restrict__ offsets, const uint8_t *__restrict__ cover,
                               const uint8_t *__restrict__ stego, int8_t *__restrict__ filp, ForcedOut *__restrict__ forced, int n_mb)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= n_mb) return;
    const int off = offsets[mb], n = offsets[mb + 1] - off;
    int8_t f[16];
    for (int j = 0; j < n; j++)
    {
        f[j] = (int8_t)((cover[off + j] ^ stego[off + j]) & 1);
        filp[off + j] = f[j];
    }
    ForcedOut o;
    const Pass1Mb p = p1[mb];
    build_forced_mb(p, f, o);
    forced[mb] = o;
}

// sum of rho in double, fixed order (chunk per thread, then a tree): the "total" the trellis result is compared with
__global__ void __launch_bounds__(1024) k_embed_total(const float *__restrict__ rho, int n, double *__restrict__ total)
{
    __shared__ double s[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    double a = 0;
    for (int i = lo; i < hi; i++) a += (double)rho[i];
    s[t] = a;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1)
    {
        if (t < d) s[t] += s[t + d];
        __syncthreads();
    }
    if (t == 0) *total = s[0];
}

} // namespace pcamv

using namespace pcamv;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, #call, e_); } while (0)

namespace pcamv {
int stc_run_device(pcamv_ctx *ctx, const uint8_t *d_cover, const float *d_rho, int n, const uint8_t *message, int an, int matrixheight,
                   const uint32_t *cols_short, int w_short, const uint32_t *cols_long, int w_long, double total, uint8_t *d_stego);
}

static int ensure_embed_buffers(pcamv_ctx *ctx)
{
    if (ctx->d_emb) return 0;
    const size_t n_mb = (size_t)ctx->fc.mb_w * ctx->fc.mb_h, cap = 16 * n_mb;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_off = 0, o_p1 = up((n_mb + 1) * sizeof(int)), o_cover = o_p1 + up(n_mb * sizeof(Pass1Mb)), o_stego = o_cover + up(cap),
                 o_filp = o_stego + up(cap), o_rho = o_filp + up(cap), o_total = o_rho + up(cap * sizeof(float)), bytes = o_total + 256;
    CK(cudaMalloc(&ctx->d_emb, bytes));
    ctx->emb_offsets = (int *)(ctx->d_emb + o_off); ctx->emb_pass1 = (void *)(ctx->d_emb + o_p1);
    ctx->emb_cover = ctx->d_emb + o_cover; ctx->emb_stego = ctx->d_emb + o_stego; ctx->emb_filp = (int8_t *)(ctx->d_emb + o_filp);
    ctx->emb_rho = (float *)(ctx->d_emb + o_rho); ctx->emb_total = (double *)(ctx->d_emb + o_total);
    return 0;
}

// Cover / rho assembly from the pass-1 results resident in HBM.  *length = number of carriers (h->info.length).
extern "C" int pcamv_embed_prepare(pcamv_ctx *ctx, int *length)
{
    if (!ctx || ctx->failed) return -1;
    cudaSetDevice(ctx->cfg.device);
    if (!length) return ctx_fail(ctx, "pcamv_embed_prepare: null argument", cudaSuccess);
    if (!ctx->frame_ready[1] || ctx->frame_last != 1 || !ctx->frame_cost_table)
        return ctx_fail(ctx, "pcamv_embed_prepare: the last analysis of this context must be pass 1 with the cost table", cudaSuccess);
    if (ensure_embed_buffers(ctx)) return -1;
    const int n_mb = ctx->fc.mb_w * ctx->fc.mb_h;
    k_embed_scan<<<1, 1024, 0, ctx->stream>>>(ctx->d_mb_results, n_mb, ctx->emb_offsets);
    k_embed_fill<<<(n_mb + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_mb_results, ctx->d_log, ctx->log_stride, ctx->d_subparts, ctx->emb_offsets,
                                                              n_mb, ctx->fc.conformant, (Pass1Mb *)ctx->emb_pass1, ctx->emb_cover, ctx->emb_rho);
    ctx->launches += 2;
    int len = 0;
    CK(cudaMemcpyAsync(&len, ctx->emb_offsets + n_mb, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->emb_length = len;
    ctx->emb_state = 1;
    *length = len;
    return 0;
}

// The trellis on the device-resident cover / rho, then flips and the forced decisions of pass 2.  an <= 0 or message == NULL:
// nothing is embedded (stego stays zero, as the reference leaves it when stc_embed gives up: every carrier whose cover bit is 1
// then counts as "to be flipped", encoder/encoder.c:1826,1848-1855).  total < 0: the sum of rho is taken on the device.
// Returns what pcamv_stc_embed returns (0 embedded, 1 not embeddable, -1 error); stego (may be NULL) receives the stego bits.
extern "C" int pcamv_embed_stc(pcamv_ctx *ctx, const uint8_t *message, int an, int matrixheight, const uint32_t *cols_short, int w_short,
                               const uint32_t *cols_long, int w_long, double total, uint8_t *stego)
{
    if (!ctx || ctx->failed) return -1;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->emb_state < 1) return ctx_fail(ctx, "pcamv_embed_stc: pcamv_embed_prepare has not run for this frame", cudaSuccess);
    const int n = ctx->emb_length, n_mb = ctx->fc.mb_w * ctx->fc.mb_h;
    int rc = 0;
    CK(cudaMemsetAsync(ctx->emb_stego, 0, (size_t)16 * n_mb, ctx->stream));
    if (message && an > 0 && n > 0)
    {
        if (an > n) return ctx_fail(ctx, "pcamv_embed_stc: the message is longer than the cover", cudaSuccess);
        if (total < 0)
        {
            k_embed_total<<<1, 1024, 0, ctx->stream>>>(ctx->emb_rho, n, ctx->emb_total);
            ctx->launches += 1;
            CK(cudaMemcpyAsync(&total, ctx->emb_total, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        rc = stc_run_device(ctx, ctx->emb_cover, ctx->emb_rho, n, message, an, matrixheight, cols_short, w_short, cols_long, w_long, total,
                            ctx->emb_stego);
        if (rc < 0) return -1;
    }
    k_embed_forced<<<(n_mb + 127) / 128, 128, 0, ctx->stream>>>((const Pass1Mb *)ctx->emb_pass1, ctx->emb_offsets, ctx->emb_cover, ctx->emb_stego, ctx->emb_filp,
                                                                (ForcedOut *)ctx->d_forced, n_mb);
    ctx->launches += 1;
    if (stego && n > 0) CK(cudaMemcpyAsync(stego, ctx->emb_stego, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->emb_state = 2;
    return rc;
}

// Whatever the host wants to see of the embed stage (any pointer may be NULL): cover / rho / stego / filp hold `length` entries
// (pcamv_embed_prepare), pass1 holds one record per macroblock (h->info.cache[] as the reference leaves it after pass 1).
extern "C" int pcamv_embed_download(pcamv_ctx *ctx, uint8_t *cover, float *rho, uint8_t *stego, int8_t *filp, pcamv_pass1_mb *pass1)
{
    if (!ctx || ctx->failed) return -1;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->emb_state < 1) return ctx_fail(ctx, "pcamv_embed_download: pcamv_embed_prepare has not run for this frame", cudaSuccess);
    const size_t n = (size_t)ctx->emb_length, n_mb = (size_t)ctx->fc.mb_w * ctx->fc.mb_h;
    if (cover && n) CK(cudaMemcpyAsync(cover, ctx->emb_cover, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (rho && n) CK(cudaMemcpyAsync(rho, ctx->emb_rho, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if ((stego || filp) && ctx->emb_state < 2) return ctx_fail(ctx, "pcamv_embed_download: pcamv_embed_stc has not run for this frame", cudaSuccess);
    if (stego && n) CK(cudaMemcpyAsync(stego, ctx->emb_stego, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (filp && n) CK(cudaMemcpyAsync(filp, ctx->emb_filp, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (pass1) CK(cudaMemcpyAsync(pass1, ctx->emb_pass1, n_mb * sizeof(Pass1Mb), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
