// pcamv_kernels.cu — sm_100a kernels of the PCAMV motion-estimation path.
//
//   k_expand_border   border replication of integer / filtered planes   (reference common/frame.c:224-301)
//   k_hpel_filter     6-tap half-pel planes H, V, HV in one pass          (reference common/mc.c:134-190,453-475)
//   k_search_batch    stateless x264_me_search_ref / x264_me_refine_qpel  (reference encoder/me.c:158-843)
//
// All are byte/integer kernels: no tensor cores (nothing here is a dense contraction).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pcamv_device.h"
#include "pcamv_me.cuh"

namespace pcamv {

// =====================================================================================================
// Border replication.  The padded buffer region [pad_x0,pad_x1) x [pad_y0,pad_y1) (pixel coordinates
// relative to pixel (0,0)) is filled from the kept region [keep_x0,keep_x1) x [keep_y0,keep_y1) by
// clamping the coordinates — the closed form of plane_expand_border's "left/right bands then
// upper/lower bands" (common/frame.c:224-244).  One launch handles up to 3 planes (blockIdx.z).
// =====================================================================================================
struct BorderArgs
{
    uint8_t *plane[3];
    int stride;
    int keep_x0, keep_x1, keep_y0, keep_y1;
    int pad_x0, pad_x1, pad_y0, pad_y1;
};

__global__ void __launch_bounds__(256) k_expand_border(const __grid_constant__ BorderArgs a)
{
    uint8_t *p = a.plane[blockIdx.z];
    const int y = a.pad_y0 + (int)blockIdx.y;
    const int cy = min(max(y, a.keep_y0), a.keep_y1 - 1);
    const bool row_inside = (y == cy);
    uint8_t *dst = p + (ptrdiff_t)y * a.stride;
    const uint8_t *src = p + (ptrdiff_t)cy * a.stride;
    // 4 pixels per thread; pad_x0 and the kept region are multiples of 4 away from each other
    for (int x = a.pad_x0 + 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x); x < a.pad_x1; x += 4 * (int)(gridDim.x * blockDim.x))
    {
        if (row_inside && x >= a.keep_x0 && x + 4 <= a.keep_x1)
            continue;                                   // interior: untouched
        uint32_t w;
        if (x + 4 <= a.keep_x0)      w = 0x01010101u * src[a.keep_x0];
        else if (x >= a.keep_x1)     w = 0x01010101u * src[a.keep_x1 - 1];
        else                         w = *(const uint32_t *)(src + x);
        *(uint32_t *)(dst + x) = w;
    }
}

void launch_expand_border(uint8_t *p0, uint8_t *p1, uint8_t *p2, int nplanes, int stride,
                          int keep_x0, int keep_x1, int keep_y0, int keep_y1,
                          int pad_x0, int pad_x1, int pad_y0, int pad_y1, void *stream)
{
    BorderArgs a;
    a.plane[0] = p0; a.plane[1] = p1; a.plane[2] = p2;
    a.stride = stride;
    a.keep_x0 = keep_x0; a.keep_x1 = keep_x1; a.keep_y0 = keep_y0; a.keep_y1 = keep_y1;
    a.pad_x0 = pad_x0; a.pad_x1 = pad_x1; a.pad_y0 = pad_y0; a.pad_y1 = pad_y1;
    const int groups = (pad_x1 - pad_x0 + 3) / 4;
    dim3 grid((groups + 255) / 256, pad_y1 - pad_y0, nplanes);
    k_expand_border<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
}

// =====================================================================================================
// Half-pel filter.  Output region x in [-8, W+8), y in [-8, H+8) of the three planes
// (x264_frame_filter's offs = start*stride - 8, width + 16; common/mc.c:453-475).
//   v  = a - 5b + 20c + 20d - 5e + f   over rows y-2..y+3      (kept unclipped, int16)
//   V  = clip((v + 16) >> 5)
//   HV = clip((6-tap over v[x-2..x+3] + 512) >> 10)
//   H  = clip((6-tap over src[x-2..x+3] + 16) >> 5)
// One CTA = HP_TW x HP_TH outputs.  The source tile (HP_TH+5 rows of 16-byte-aligned 144-byte
// segments) is staged into shared memory with TMA bulk row copies completing on one mbarrier; the
// vertical intermediates live in shared memory as int16 so the HV pass never touches HBM.
// =====================================================================================================
#define HP_TW 128
#define HP_TH 16
#define HP_SEG 160                       // bytes staged per source row (16-byte aligned start, covers TW+5+14)
#define HP_ROWS (HP_TH + 5)
#define HP_VW (HP_TW + 8)                // intermediates per row (TW+5 used)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared (SASS: UBLKCP); size and both addresses multiples of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a + f - 5 * (b + e) + 20 * (c + d); }
__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)min(max(v, 0), 255); }

__global__ void __launch_bounds__(256) k_hpel_filter(const uint8_t *__restrict__ src, uint8_t *__restrict__ dsth,
                                                     uint8_t *__restrict__ dstv, uint8_t *__restrict__ dstc,
                                                     int stride, int width, int height)
{
    __shared__ __align__(128) uint8_t s_src[HP_ROWS][HP_SEG];
    __shared__ __align__(16) int16_t s_v[HP_TH][HP_VW];
    __shared__ __align__(8) uint64_t s_bar;

    // tile origin in pixel coordinates; x0 is congruent to -8 mod 128, so (x0 - 8) is 16-byte aligned
    // inside the padded buffer (pixel 0 sits at column 32 of a 16-byte-aligned row).
    const int x0 = -8 + (int)blockIdx.x * HP_TW;
    const int y0 = -8 + (int)blockIdx.y * HP_TH;
    const int seg_x = x0 - 8;                              // first staged column (covers x0-2)
    const int tid = threadIdx.x;

    if (tid == 0)
        mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid < 32)
    {
        if (tid == 0)
            mbar_expect_tx(&s_bar, HP_ROWS * HP_SEG);
        __syncwarp();
        if (tid < HP_ROWS)
            tma_load_1d(&s_src[tid][0], src + (ptrdiff_t)(y0 - 2 + tid) * stride + seg_x, HP_SEG, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    // vertical pass: intermediates for columns x0-2 .. x0+TW+2 (index c = column - (x0-2))
    for (int i = tid; i < HP_TH * (HP_TW + 5); i += 256)
    {
        const int r = i / (HP_TW + 5), c = i - r * (HP_TW + 5);
        const int sx = c + 6;                              // column x0-2+c sits at staged byte (x0-2+c) - seg_x = c + 6
        const int v = tap6(s_src[r][sx], s_src[r + 1][sx], s_src[r + 2][sx], s_src[r + 3][sx], s_src[r + 4][sx], s_src[r + 5][sx]);
        s_v[r][c] = (int16_t)v;
    }
    __syncthreads();

    // horizontal passes: each thread produces 4 consecutive pixels of one row for all three planes
    const int x_end = width + 8, y_end = height + 8;
    for (int i = tid; i < HP_TH * (HP_TW / 4); i += 256)
    {
        const int r = i / (HP_TW / 4), g = i - r * (HP_TW / 4);
        const int x = x0 + 4 * g, y = y0 + r;
        if (x >= x_end || y >= y_end)
            continue;
        uint32_t wh = 0, wv = 0, wc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int c = 4 * g + k + 2;                   // index of column x+k in s_v
            const int sx = 4 * g + k + 8;                  // index of column x+k in s_src
            const uint8_t *row = &s_src[r + 2][0];
            const int hv = tap6(s_v[r][c - 2], s_v[r][c - 1], s_v[r][c], s_v[r][c + 1], s_v[r][c + 2], s_v[r][c + 3]);
            const int hh = tap6(row[sx - 2], row[sx - 1], row[sx], row[sx + 1], row[sx + 2], row[sx + 3]);
            wv |= clip8((s_v[r][c] + 16) >> 5) << (8 * k);
            wc |= clip8((hv + 512) >> 10) << (8 * k);
            wh |= clip8((hh + 16) >> 5) << (8 * k);
        }
        const ptrdiff_t o = (ptrdiff_t)y * stride + x;
        *(uint32_t *)(dsth + o) = wh;
        *(uint32_t *)(dstv + o) = wv;
        *(uint32_t *)(dstc + o) = wc;
    }
}

void launch_hpel_filter(const uint8_t *src, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                        int stride, int width, int height, void *stream)
{
    dim3 grid((width + 16 + HP_TW - 1) / HP_TW, (height + 16 + HP_TH - 1) / HP_TH);
    k_hpel_filter<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dsth, dstv, dstc, stride, width, height);
}

// =====================================================================================================
// Stateless search batch: one warp (lane team) per recorded call.
// =====================================================================================================
#define SB_WARPS 4

__device__ __forceinline__ void stage_fenc_mb(const DevFrameCtx &fc, int mb_x, int mb_y, uint8_t *sy, uint8_t *su, uint8_t *sv)
{
    const int lane = threadIdx.x & 31;
    if (lane < 16)
    {
        const uint4 v = *(const uint4 *)(fc.fenc_y + (size_t)(16 * mb_y + lane) * fc.stride_y + 16 * mb_x);
        *(uint4 *)(sy + 16 * lane) = v;
    }
    else if (lane < 24)
    {
        const int r = lane - 16;
        *(uint2 *)(su + 8 * r) = *(const uint2 *)(fc.fenc_u + (size_t)(8 * mb_y + r) * fc.stride_c + 8 * mb_x);
    }
    else
    {
        const int r = lane - 24;
        *(uint2 *)(sv + 8 * r) = *(const uint2 *)(fc.fenc_v + (size_t)(8 * mb_y + r) * fc.stride_c + 8 * mb_x);
    }
    __syncwarp();
}

__device__ __forceinline__ void make_block(const DevFrameCtx &fc, MeBlock &b, int slot, int mb_x, int mb_y, int xoff, int yoff,
                                           int i_pixel, const uint8_t *sy, const uint8_t *su, const uint8_t *sv)
{
    const DevRef &rf = fc.ref[slot];
    b.i_pixel = i_pixel; b.bw = pix_w(i_pixel); b.bh = pix_h(i_pixel);
    b.fenc = sy + yoff * 16 + xoff;
    b.fenc_u = su + (yoff >> 1) * 8 + (xoff >> 1);
    b.fenc_v = sv + (yoff >> 1) * 8 + (xoff >> 1);
    b.stride = fc.stride_y; b.stride_c = fc.stride_c;
    const ptrdiff_t off = (ptrdiff_t)(16 * mb_y + yoff) * fc.stride_y + 16 * mb_x + xoff;
#pragma unroll
    for (int k = 0; k < 4; k++) b.ref[k] = rf.y[k] + off;
    const ptrdiff_t offc = (ptrdiff_t)(8 * mb_y + (yoff >> 1)) * fc.stride_c + 8 * mb_x + (xoff >> 1);
    b.ref_u = rf.u + offc; b.ref_v = rf.v + offc;
    b.integral = rf.integral ? rf.integral + off : nullptr;
    b.integral4 = rf.integral4 ? rf.integral4 + off : nullptr;
#if defined(PCAMV_CHECKED)
    b.chk_lo = rf.base; b.chk_hi = rf.base + rf.bytes;
#endif
}

// mvsads: --me tesa candidate lists, one of mvsads_cap entries per call of this launch (null otherwise)
__global__ void __launch_bounds__(SB_WARPS * 32) k_search_batch(const __grid_constant__ DevFrameCtx fc,
                                                               const pcamv_me_call *__restrict__ calls, int n,
                                                               pcamv_me_result *__restrict__ results,
                                                               unsigned long long *mvsads, int mvsads_cap)
{
    __shared__ __align__(16) uint8_t s_fenc[SB_WARPS][384];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * SB_WARPS + warp;
    if (idx >= n)
        return;
    const pcamv_me_call &c = calls[idx];
    uint8_t *sy = s_fenc[warp], *su = sy + 256, *sv = sy + 320;
    stage_fenc_mb(fc, c.mb_x, c.mb_y, sy, su, sv);

    MeEnv env;
    env.cost_mv = fc.tab.cost_mv;
    env.cost_mv_fpel[0] = env.cost_mv_fpel[1] = env.cost_mv_fpel[2] = env.cost_mv_fpel[3] = nullptr;
    env.me_method = fc.me_method; env.me_range = fc.me_range; env.subme = fc.subme; env.chroma_me = fc.chroma_me && fc.subme >= 5;
    env.mbcmp_satd = fc.subme > 1;
    env.mvsads = mvsads ? mvsads + (size_t)idx * mvsads_cap : nullptr;
#pragma unroll
    for (int k = 0; k < 2; k++)
    {
        env.mv_min_fpel[k] = c.mv_min_fpel[k]; env.mv_max_fpel[k] = c.mv_max_fpel[k];
        env.mv_min_spel[k] = c.mv_min_spel[k]; env.mv_max_spel[k] = c.mv_max_spel[k];
    }
    MeBlock b;
    make_block(fc, b, c.ref_slot, c.mb_x, c.mb_y, c.xoff, c.yoff, c.i_pixel, sy, su, sv);
    block_set_mvp(b, env, c.mvp[0], c.mvp[1]);

    MeResult m;
    int thresh = c.thresh_in;
    if (c.mode == 0)
    {
        int mvc[PCAMV_MAX_MVC][2];
        const int nm = min(c.i_mvc, PCAMV_MAX_MVC);
        for (int i = 0; i < nm; i++) { mvc[i][0] = c.mvc[i][0]; mvc[i][1] = c.mvc[i][1]; }
        m.mv[0] = m.mv[1] = 0; m.cost = 0; m.cost_mv = 0;
        me_search_ref<1>(env, b, mvc, nm, c.has_thresh ? &thresh : nullptr, m);
    }
    else
    {
        m.mv[0] = c.mv_in[0]; m.mv[1] = c.mv_in[1]; m.cost = c.cost_in; m.cost_mv = c.cost_mv_in;
        me_refine_qpel(env, b, m, c.i_ref_cost);
    }
    if ((threadIdx.x & 31) == 0)
    {
        pcamv_me_result r;
        r.mv[0] = (int16_t)m.mv[0]; r.mv[1] = (int16_t)m.mv[1];
        r.cost = m.cost; r.cost_mv = m.cost_mv; r.thresh_out = thresh;
        results[idx] = r;
    }
}

// with --me tesa the calls go out in chunks of `chunk` (what the scratch holds lists for), else in one launch
void launch_search_batch(const DevFrameCtx &fc, const pcamv_me_call *calls, int n, pcamv_me_result *results,
                         unsigned long long *mvsads, int mvsads_cap, int chunk, void *stream)
{
    if (n <= 0) return;
    if (!mvsads) chunk = n;
    for (int at = 0; at < n; at += chunk)
    {
        const int m = n - at < chunk ? n - at : chunk;
        k_search_batch<<<(m + SB_WARPS - 1) / SB_WARPS, SB_WARPS * 32, 0, (cudaStream_t)stream>>>(fc, calls + at, m, results + at, mvsads, mvsads_cap);
    }
}

// =====================================================================================================
// Integral plane for the exhaustive searches (reference common/mc.c:311-345 integral_init8h / integral_init8v as
// x264_frame_filter strings them together, mc.c:477-511): out[y][x] = sum of the 8x8 pixels of the PADDED integer luma
// plane whose top-left corner is (x, y), as uint16 (<= 64 * 255, so the reference's modular running sums give exactly
// this).  Positions whose box leaves the padded plane are never read by a search (MV limits keep a 16x16 block + the
// 3-column overshoot inside the 32-pixel border) and are written as 0.  Horizontal 8-sums of a tile go through
// shared memory, then 8 of them are added vertically: 1 byte read + 2 bytes written per pixel, HBM-bound.
// =====================================================================================================
// N = 8, and N = 4 when sub-8x8 partitions are searched exhaustively (the reference's second integral plane, mc.c:322-337).
#define BS_TW 64
#define BS_TH 32
template <int N>
__global__ void __launch_bounds__(256) k_box_sum(const uint8_t *__restrict__ src, uint16_t *__restrict__ dst, int stride, int rows)
{
    __shared__ uint16_t hs[BS_TH + N - 1][BS_TW];
    const int x0 = blockIdx.x * BS_TW, y0 = blockIdx.y * BS_TH;
    for (int i = threadIdx.x; i < (BS_TH + N - 1) * BS_TW; i += 256)
    {
        const int r = i / BS_TW, c = i - r * BS_TW;
        const int y = y0 + r, x = x0 + c;
        int s = 0;
        if (y < rows && x + N <= stride)
        {
            const uint8_t *p = src + (size_t)y * stride + x;
#pragma unroll
            for (int k = 0; k < N; k++) s += p[k];
        }
        hs[r][c] = (uint16_t)s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BS_TH * BS_TW; i += 256)
    {
        const int r = i / BS_TW, c = i - r * BS_TW;
        const int y = y0 + r, x = x0 + c;
        if (y < rows && x < stride)
        {
            int s = 0;
            if (y + N <= rows && x + N <= stride)
            {
#pragma unroll
                for (int k = 0; k < N; k++) s += hs[r + k][c];
            }
            dst[(size_t)y * stride + x] = (uint16_t)s;
        }
    }
}

void launch_box_sum(const uint8_t *src_padded, uint16_t *dst_padded, int stride, int rows, int box, void *stream)
{
    dim3 grid((stride + BS_TW - 1) / BS_TW, (rows + BS_TH - 1) / BS_TH);
    if (box == 4) k_box_sum<4><<<grid, 256, 0, (cudaStream_t)stream>>>(src_padded, dst_padded, stride, rows);
    else          k_box_sum<8><<<grid, 256, 0, (cudaStream_t)stream>>>(src_padded, dst_padded, stride, rows);
}

// =====================================================================================================
// Integer issue-rate microbenchmark: the roofline denominator for the search kernels (SURVEY.md 8(d):
// "measure the int32 issue peak with a micro-benchmark on the box").  Every thread runs `iters` rounds
// of 8 independent chains of the instruction mix the SAD/SATD inner loops are made of
// (VABSDIFF4.ACC, IADD3, LOP3); the result is written so nothing is optimised away.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_int_peak(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed * (threadIdx.x + 1) + k; b[k] = seed ^ (0x9E3779B9u * (k + 1)); }
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            a[k] = __vsadu4(a[k], b[k]) + a[k];       // VABSDIFF4.U8.ACC
            b[k] = (b[k] + a[k]) + 0x01010101u;       // IADD3
            a[k] = (a[k] ^ b[k]) & 0x7f7f7f7fu;       // LOP3
            b[k] = b[k] - (a[k] >> 1);                // SHF/IADD
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r += a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

void launch_int_peak(uint32_t *out, int blocks, int iters, void *stream)
{
    k_int_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 12345u);
}

} // namespace pcamv
