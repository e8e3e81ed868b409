// pcamv_recon.cu — kernels and C-ABI entry of the device-side reference frame (device code and the
// reference lines it follows: pcamv_recon.cuh).
//
//   k_recon          one team per macroblock (independent): motion compensation of the final decision, residual transform /
//                    quantisation / decimation / reconstruction, the macroblock into the integer planes of the target slot,
//                    its surviving-coefficient flags for the filter
//   k_recon_patch    macroblocks the host reconstructed differently by construction (quirk q1): their pixels and flags as the
//                    host has them, uploaded by the caller
//   k_deblock        the in-loop filter as a macroblock wavefront: a macroblock's edges read and write the left neighbour's
//                    right columns and the rows above up to the top-right neighbour, so row y runs two macroblocks behind
//                    row y - 1 (same dependency shape as the analysis; rows are claimed in increasing order, no deadlock)
// followed by the border / half-pel / integral kernels every reference slot gets (filter_slot).
#include <cuda_runtime.h>
#include <stdint.h>
#include <new>
#include "pcamv_ctx.h"
#include "pcamv_recon.cuh"

namespace pcamv {

__device__ __forceinline__ void rc_stage_fenc(const DevFrameCtx &fc, int mb_x, int mb_y, MbWork &w)
{
    const int lane = threadIdx.x & 31;
    if (lane < 16)
        *(uint4 *)(w.fenc_y + 16 * lane) = *(const uint4 *)(fc.fenc_y + (size_t)(16 * mb_y + lane) * fc.stride_y + 16 * mb_x);
    else if (lane < 24)
        *(uint2 *)(w.fenc_u + 8 * (lane - 16)) = *(const uint2 *)(fc.fenc_u + (size_t)(8 * mb_y + lane - 16) * fc.stride_c + 8 * mb_x);
    else
        *(uint2 *)(w.fenc_v + 8 * (lane - 24)) = *(const uint2 *)(fc.fenc_v + (size_t)(8 * mb_y + lane - 24) * fc.stride_c + 8 * mb_x);
    __syncwarp();
}

#define RC_WARPS 4
__global__ void __launch_bounds__(RC_WARPS * 32) k_recon(const __grid_constant__ DevFrameCtx fc, const __grid_constant__ FrameParams fp,
                                                         const ReconPlanes rp, int n_mb)
{
    __shared__ MbWork s_work[RC_WARPS];
    __shared__ MbResult s_res[RC_WARPS];
    __shared__ __align__(16) unsigned char s_ctx[RC_WARPS][sizeof(MbCtx)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * RC_WARPS + warp;
    if (mb >= n_mb)
        return;
    {
        const uint32_t *src = (const uint32_t *)(fp.results + mb);
        uint32_t *dst = (uint32_t *)&s_res[warp];
        for (int i = lane; i < (int)(sizeof(MbResult) / 4); i += 32) dst[i] = src[i];
        __syncwarp();
    }
    MbCtx &c = *new (s_ctx[warp]) MbCtx(fc, fp, s_work[warp]);
    c.mb_x = mb % fc.mb_w; c.mb_y = mb / fc.mb_w; c.mb_xy = mb;
    rc_stage_fenc(fc, c.mb_x, c.mb_y, s_work[warp]);
    recon_mb(c, s_res[warp], rp);
}

struct ReconPatch { int32_t mb_xy; uint16_t nnz, pad; uint8_t y[256], u[64], v[64]; };
static_assert(sizeof(ReconPatch) == sizeof(pcamv_recon_patch), "patch layout");

__global__ void __launch_bounds__(96) k_recon_patch(const ReconPatch *__restrict__ patches, const ReconPlanes rp, int mb_w)
{
    const ReconPatch &p = patches[blockIdx.x];
    const int mb_x = p.mb_xy % mb_w, mb_y = p.mb_xy / mb_w, t = threadIdx.x;
    if (t < 64)
    {
        const int y = t >> 2, x = (t & 3) << 2;
        *(uint32_t *)(rp.y + (size_t)(16 * mb_y + y) * rp.stride_y + 16 * mb_x + x) = *(const uint32_t *)(p.y + 16 * y + x);
    }
    else
    {
        const int k = t - 64, pl = k >> 4, y = (k >> 1) & 7, x = (k & 1) << 2;
        *(uint32_t *)((pl ? rp.v : rp.u) + (size_t)(8 * mb_y + y) * rp.stride_c + 8 * mb_x + x) = *(const uint32_t *)((pl ? p.v : p.u) + 8 * y + x);
    }
    if (t == 0) rp.nnz[p.mb_xy] = p.nnz;
}

__device__ __forceinline__ int rc_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rc_st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// progress[row] = macroblocks of the row finished; row_claim = next row to hand out
// boundary strengths of the whole frame, one thread per 4-pixel piece (32 per macroblock): no dependence on filtered pixels
__global__ void __launch_bounds__(256) k_deblock_bs(const DbFrame f, const DeblockParams dp, int n_mb, uint8_t *__restrict__ bs)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 32 * n_mb)
        return;
    const int mb = t >> 5;
    bs[t] = (uint8_t)db_strength_piece(f, dp, mb % f.mb_w, mb / f.mb_w, t & 31);
}

__global__ void __launch_bounds__(RC_WARPS * 32) k_deblock(const ReconPlanes rp, const DeblockParams dp, int mb_w, int mb_h,
                                                           const uint8_t *__restrict__ bs, int *progress, int *row_claim)
{
    __shared__ int s_group;
    __shared__ __align__(16) uint8_t s_stage[RC_WARPS][(DB_STAGE_BYTES + 15) & ~15];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;)
    {
        __syncthreads();
        if (threadIdx.x == 0) s_group = atomicAdd(row_claim, 1);
        __syncthreads();
        if (s_group * RC_WARPS >= mb_h)
            return;
        const int row = s_group * RC_WARPS + warp;
        if (row >= mb_h)
            continue;
        for (int x = 0; x < mb_w; x++)
        {
            if (row > 0)
            {
                const int need = min(x + 2, mb_w);
                unsigned ns = 32;
                while (rc_ld_acquire(progress + row - 1) < need)
                {
                    __nanosleep(ns);
                    if (ns < 1024) ns <<= 1;
                }
            }
            deblock_mb(rp, dp, x, row, bs + 32 * (size_t)(row * mb_w + x), s_stage[warp]);
            __syncwarp();
            if (lane == 0)
            {
                __threadfence();
                rc_st_release(progress + row, x + 1);
            }
        }
    }
}

} // namespace pcamv

using namespace pcamv;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, #call, e_); } while (0)

// Build reference slot `slot` (POC `poc`) on the device from the decisions the analysis of `pass` (the frame's FINAL pass: 2, or
// 0 when nothing is embedded) left in HBM: reconstruction of every macroblock, patches for the macroblocks the host says it
// reconstructed differently (records with early_skip == 2, see pcamv_recon.cuh), deblocking, borders, half-pel planes.
// The slot must not be one of the frame's own references.  Replaces x264_macroblock_encode's reconstruction + x264_fdec_filter_row
// (encoder/encoder.c:1004-1052) as the producer of the next frame's reference planes, and pcamv_put_ref with them.
extern "C" int pcamv_reconstruct_ref(pcamv_ctx *ctx, int slot, int poc, int pass, const pcamv_recon_patch *patches, int n_patches)
{
    if (!ctx || ctx->failed) return -1;
    cudaSetDevice(ctx->cfg.device);
    if (slot < 0 || slot >= ctx->cfg.max_refs + 2) return ctx_fail(ctx, "pcamv_reconstruct_ref: bad slot", cudaSuccess);
    if (pass < 0 || pass > 2 || pass == 1 || !ctx->frame_ready[pass] || ctx->frame_last != pass)
        return ctx_fail(ctx, "pcamv_reconstruct_ref: the last analysis of this context must be the frame's final pass (0 or 2)", cudaSuccess);
    if (n_patches < 0 || (n_patches && !patches)) return ctx_fail(ctx, "pcamv_reconstruct_ref: bad patch list", cudaSuccess);
    const DevFrameCtx &fc = ctx->fc;
    const FrameParams &fp = ctx->fp[pass];
    for (int i = 0; i < fp.n_ref; i++)
        if (fp.ref_slot[i] == slot) return ctx_fail(ctx, "pcamv_reconstruct_ref: the target slot is one of the frame's references", cudaSuccess);
    const int n_mb = fc.mb_w * fc.mb_h;
    // one allocation: coefficient flags (uint16 per macroblock), then the 32 boundary strengths of every macroblock
    const size_t nnz_bytes = ((size_t)n_mb * sizeof(uint16_t) + 255) & ~(size_t)255;
    if (!ctx->d_recon_nnz) CK(cudaMalloc(&ctx->d_recon_nnz, nnz_bytes + (size_t)32 * n_mb));
    uint8_t *d_bs = (uint8_t *)ctx->d_recon_nnz + nnz_bytes;
    DevRef &r = ctx->fc.ref[slot];
    r.valid = 0;
    ReconPlanes rp;
    rp.y = r.y[0]; rp.u = r.u; rp.v = r.v; rp.stride_y = fc.stride_y; rp.stride_c = fc.stride_c; rp.nnz = ctx->d_recon_nnz;
    k_recon<<<(n_mb + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, ctx->stream>>>(fc, fp, rp, n_mb);
    ctx->launches += 1;
    if (n_patches)
    {
        for (int i = 0; i < n_patches; i++)
            if (patches[i].mb_xy < 0 || patches[i].mb_xy >= n_mb) return ctx_fail(ctx, "pcamv_reconstruct_ref: patch outside the frame", cudaSuccess);
        if (ctx->recon_patch_cap < n_patches)
        {
            cudaFree(ctx->d_recon_patches); ctx->d_recon_patches = nullptr; ctx->recon_patch_cap = 0;
            CK(cudaMalloc(&ctx->d_recon_patches, (size_t)(n_patches + 64) * sizeof(ReconPatch)));
            ctx->recon_patch_cap = n_patches + 64;
        }
        CK(cudaMemcpyAsync(ctx->d_recon_patches, patches, (size_t)n_patches * sizeof(ReconPatch), cudaMemcpyHostToDevice, ctx->stream));
        k_recon_patch<<<n_patches, 96, 0, ctx->stream>>>((const ReconPatch *)ctx->d_recon_patches, rp, fc.mb_w);
        ctx->launches += 1;
    }
    // the slice header disables the filter when it could not change anything (encoder/encoder.c:158-168, constant QP)
    const int off_min = ctx->cfg.deblock_alpha_c0_offset < ctx->cfg.deblock_beta_offset ? ctx->cfg.deblock_alpha_c0_offset : ctx->cfg.deblock_beta_offset;
    if (!ctx->cfg.no_deblock && 15 < fc.tab.qp + off_min)
    {
        DbFrame f;
        f.type = fp.cur.type; f.ref8 = fp.cur.ref8; f.mv4 = fp.cur.mv4; f.nnz = ctx->d_recon_nnz; f.mb_w = fc.mb_w;
        DeblockParams dp;
        dp.disable = 0; dp.alpha_c0_offset = ctx->cfg.deblock_alpha_c0_offset; dp.beta_offset = ctx->cfg.deblock_beta_offset;
        dp.qp = fc.tab.qp; dp.qp_chroma = fc.tab.chroma_qp; dp.chroma_qp_offset = ctx->cfg.chroma_qp_offset;
        dp.no_sub8x8_all = !(fc.analyse_inter & 0x20);
        CK(cudaMemsetAsync(ctx->d_progress, 0, (2 * fc.mb_h + 2) * sizeof(int), ctx->stream));
        const int groups = (fc.mb_h + RC_WARPS - 1) / RC_WARPS;
        k_deblock_bs<<<(32 * n_mb + 255) / 256, 256, 0, ctx->stream>>>(f, dp, n_mb, d_bs);
        k_deblock<<<groups, RC_WARPS * 32, 0, ctx->stream>>>(rp, dp, fc.mb_w, fc.mb_h, d_bs, ctx->d_progress, ctx->d_progress + fc.mb_h);
        ctx->launches += 2;
    }
    if (filter_slot(ctx, slot)) return -1;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    r.poc = poc; r.valid = 1;
    return 0;
}
