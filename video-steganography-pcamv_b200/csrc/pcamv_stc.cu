// pcamv_stc.cu — the embedder's syndrome-trellis code on the GPU.
//
// Behavioural contract (bit-exact): reference embed.h:309-548 (stc_embed): Viterbi over 2^h states (h = 10 in the encoder,
// encoder/encoder.c:1843), one trellis column per cover element, the state shifted right by one message bit at the end of
// every block of the banded parity-check matrix; float path metrics, `<=` tie-breaks, a bit-packed survivor path, then the
// backward trace that emits the stego bits.  Only float additions and comparisons occur, each computed exactly as the
// scalar code computes it (no contraction possible: there is no multiply), so the metrics are IEEE-identical.
//
//   forward   one CTA of 2^h threads per vector: thread s owns state s.  Per cover element:
//               stay  = price[s]       + (cover ? rho : 0)      (stego bit 0, state unchanged)
//               other = price[s ^ col] + (cover ? 0 : rho)      (stego bit 1)
//               price'[s] = other <= stay ? other : stay;  path bit = other <= stay
//             which is the reference's pairwise update (embed.h:439-468) written per state: the pair (m, m ^ col) sets the
//             survivor bit of either member exactly when the transition from the other member is not worse.
//             Path bits leave as one 32-bit ballot per warp: 2^h / 8 bytes per cover element, coalesced.
//   backward  one warp: the 128-byte path line of element idx is known before the state is, so lines are fetched eight
//             elements ahead and the state-dependent word is picked with a shuffle.
// The sub-matrix columns come from the caller (the host encoder draws them with the reference's getMatrix, embed.h:276-306);
// the block schedule (embed.h:376-392) is rebuilt here in the same double arithmetic.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "pcamv_ctx.h"

namespace pcamv {

// per cover element: the column as the forward pass masks it, as the backward trace masks it, end-of-block flag, the block's
// message bit.  The two masks agree whenever the message has at least `matrixheight` bits; for shorter messages the
// reference's forward pass starts from the full mask and only then shrinks it (embed.h:415,482-483) while its backward
// trace grows the mask from zero (embed.h:519-524), and both are reproduced as they are.
struct StcElem { uint32_t col, colb; uint8_t last, msg, pad[2]; };

template <int H>
__global__ void __launch_bounds__(1 << H) k_stc_forward(const uint8_t *__restrict__ cover, const float *__restrict__ rho,
                                                        const StcElem *__restrict__ elem, int n, uint32_t *__restrict__ path,
                                                        float *__restrict__ total_price)
{
    constexpr int S = 1 << H;
    __shared__ float pr[2][S];
    const int s = threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    pr[0][s] = s == 0 ? 0.0f : inf;
    __syncthreads();
    int cur = 0;
    for (int idx = 0; idx < n; idx++)
    {
        const StcElem e = elem[idx];
        const float p = rho[idx];
        const bool one = cover[idx] != 0;
        const float stay = __fadd_rn(pr[cur][s], one ? p : 0.0f);
        const float other = __fadd_rn(pr[cur][s ^ (int)e.col], one ? 0.0f : p);
        const bool take = other <= stay;
        const unsigned bits = __ballot_sync(0xffffffffu, take);
        if ((s & 31) == 0)
            path[(size_t)idx * (S / 32) + (s >> 5)] = bits;
        pr[cur ^ 1][s] = take ? other : stay;
        __syncthreads();
        cur ^= 1;
        if (e.last)
        {
            // end of a block: the next message bit selects the states that survive, the state register shifts right
            pr[cur ^ 1][s] = s < S / 2 ? pr[cur][2 * s + e.msg] : inf;
            __syncthreads();
            cur ^= 1;
        }
    }
    if (s == 0)
        *total_price = pr[cur][0];
}

// one warp; stego[idx] = survivor bit of the running state, which then moves along the column when the bit is set
template <int H>
__global__ void __launch_bounds__(32) k_stc_backward(const StcElem *__restrict__ elem, int n, const uint32_t *__restrict__ path,
                                                     uint8_t *__restrict__ stego)
{
    constexpr int W = (1 << H) / 32;              // path words per element (<= 32)
    const int lane = threadIdx.x;
    uint32_t state = 0;
    for (int top = n - 1; top >= 0; top -= 8)
    {
        uint32_t line[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
            line[k] = (top - k >= 0 && lane < W) ? path[(size_t)(top - k) * W + lane] : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int idx = top - k;
            if (idx < 0) break;
            const StcElem e = elem[idx];
            if (e.last)
                state = (state << 1) | e.msg;             // entering the block from behind: its message bit comes back in
            const uint32_t word = __shfl_sync(0xffffffffu, line[k], (int)(state >> 5));
            const uint32_t bit = (word >> (state & 31)) & 1u;
            if (lane == 0) stego[idx] = (uint8_t)bit;
            if (bit) state ^= e.colb;
        }
    }
}

} // namespace pcamv

using namespace pcamv;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, #call, e_); } while (0)

// The trellis on device-resident cover / rho (n elements); `message` (an bits) and the sub-matrix columns are host arrays;
// `total` = sum of rho over the elements (what the reference compares the trellis result with, embed.h:496-505).
// Writes the stego bits to d_stego (device, n bytes; elements the block schedule does not reach stay untouched).
// Returns 0 = embedded, 1 = not embeddable (d_stego untouched), -1 = error.
namespace pcamv {
int stc_run_device(pcamv_ctx *ctx, const uint8_t *d_cover, const float *d_rho, int n, const uint8_t *message, int an, int matrixheight,
                   const uint32_t *cols_short, int w_short, const uint32_t *cols_long, int w_long, double total, uint8_t *d_stego)
{
    if (!message || !cols_short || !cols_long || n <= 0 || an <= 0 || an > n)
        return ctx_fail(ctx, "pcamv_stc_embed: bad argument", cudaSuccess);
    if (matrixheight < 7 || matrixheight > 10)
        return ctx_fail(ctx, "pcamv_stc_embed: matrix height must be 7..10 (the encoder uses 10)", cudaSuccess);
    // The per-state survivor rule of k_stc_forward equals the reference's pairwise update (embed.h:439-468) only for columns
    // whose masked value is non-zero; the mask never drops bit 0, so an odd column is always safe.  Every column getMatrix
    // can produce is odd (embed.h:276-306: the tabulated ones, and the generated ones have bit 0 forced), so an even one is a
    // caller error, not a case to emulate.
    for (int k = 0; k < w_short; k++)
        if (!(cols_short[k] & 1u)) return ctx_fail(ctx, "pcamv_stc_embed: sub-matrix columns must be odd (bit 0 set)", cudaSuccess);
    for (int k = 0; k < w_long; k++)
        if (!(cols_long[k] & 1u)) return ctx_fail(ctx, "pcamv_stc_embed: sub-matrix columns must be odd (bit 0 set)", cudaSuccess);
    // block schedule (embed.h:376-392) and per-element columns with the shrinking mask of the last h blocks (embed.h:482-483)
    const double invalpha = (double)n / an;
    const int shorter = (int)floor(invalpha), longer = (int)ceil(invalpha);
    if (shorter != w_short || longer != w_long)
        return ctx_fail(ctx, "pcamv_stc_embed: sub-matrix widths must be floor / ceil of n / an", cudaSuccess);
    std::vector<StcElem> el((size_t)n);
    uint32_t colmask = (1u << matrixheight) - 1;
    int worm = 0, index = 0;
    for (int b = 0; b < an; b++)
    {
        const bool wide = worm + longer <= (b + 1) * invalpha + 0.5;
        const int width = wide ? longer : shorter;
        const uint32_t *cols = wide ? cols_long : cols_short;
        worm += width;
        if (index + width > n)
            return ctx_fail(ctx, "pcamv_stc_embed: block schedule overruns the cover", cudaSuccess);
        const int left = an - b;                              // blocks from here to the end
        const uint32_t backmask = left <= matrixheight ? (1u << left) - 1u : (1u << matrixheight) - 1u;
        for (int k = 0; k < width; k++, index++)
        {
            el[index].col = cols[k] & colmask;
            el[index].colb = cols[k] & backmask;
            el[index].last = k == width - 1;
            el[index].msg = message[b] ? 1 : 0;
            el[index].pad[0] = el[index].pad[1] = 0;
        }
        if (an - b <= matrixheight)
            colmask >>= 1;
    }
    const int used = index;           // the schedule's total width; elements past it (if any) are never touched, as in the reference
    const size_t words = (size_t)used * ((1u << matrixheight) / 32);
    // one device block per context, grown on demand: path | elems | total
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_path = 0, o_el = up(words * sizeof(uint32_t)), o_total = o_el + up(used * sizeof(StcElem)), need = o_total + 256;
    if (ctx->stc_bytes < need)
    {
        cudaFree(ctx->d_stc); ctx->d_stc = nullptr; ctx->stc_bytes = 0;
        CK(cudaMalloc(&ctx->d_stc, need + need / 4));
        ctx->stc_bytes = need + need / 4;
    }
    uint32_t *d_path = (uint32_t *)(ctx->d_stc + o_path); StcElem *d_el = (StcElem *)(ctx->d_stc + o_el);
    float *d_total = (float *)(ctx->d_stc + o_total);
    CK(cudaMemcpyAsync(d_el, el.data(), used * sizeof(StcElem), cudaMemcpyHostToDevice, ctx->stream));
    switch (matrixheight)
    {
    case 7:  k_stc_forward<7><<<1, 128, 0, ctx->stream>>>(d_cover, d_rho, d_el, used, d_path, d_total); break;
    case 8:  k_stc_forward<8><<<1, 256, 0, ctx->stream>>>(d_cover, d_rho, d_el, used, d_path, d_total); break;
    case 9:  k_stc_forward<9><<<1, 512, 0, ctx->stream>>>(d_cover, d_rho, d_el, used, d_path, d_total); break;
    default: k_stc_forward<10><<<1, 1024, 0, ctx->stream>>>(d_cover, d_rho, d_el, used, d_path, d_total); break;
    }
    float total_price = 0;
    CK(cudaMemcpyAsync(&total_price, d_total, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));          // (also: `el` may go out of scope)
    ctx->launches += 1;
    if ((double)total_price >= total)
        return 1;               // "The syndrome is not in the range of the syndrome matrix." (embed.h:503-512)
    switch (matrixheight)
    {
    case 7:  k_stc_backward<7><<<1, 32, 0, ctx->stream>>>(d_el, used, d_path, d_stego); break;
    case 8:  k_stc_backward<8><<<1, 32, 0, ctx->stream>>>(d_el, used, d_path, d_stego); break;
    case 9:  k_stc_backward<9><<<1, 32, 0, ctx->stream>>>(d_el, used, d_path, d_stego); break;
    default: k_stc_backward<10><<<1, 32, 0, ctx->stream>>>(d_el, used, d_path, d_stego); break;
    }
    ctx->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
} // namespace pcamv

// returns 0 = embedded, 1 = the syndrome is not in the range of the matrix (the reference's stc_embed returns 0 and leaves
// stego untouched), -1 = error (pcamv_last_error)
extern "C" int pcamv_stc_embed(pcamv_ctx *ctx, const uint8_t *cover, int n, const uint8_t *message, int an, const float *rho,
                               uint8_t *stego, int matrixheight, const uint32_t *cols_short, int w_short,
                               const uint32_t *cols_long, int w_long)
{
    if (!ctx || ctx->failed) return -1;
    cudaSetDevice(ctx->cfg.device);
    if (!cover || !message || !rho || !stego || !cols_short || !cols_long || n <= 0 || an <= 0 || an > n)
        return ctx_fail(ctx, "pcamv_stc_embed: bad argument", cudaSuccess);
    // host-pointer entry: stage cover / rho in a block of their own, run, fetch the stego bits
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t need = 2 * up(n) + up((size_t)n * sizeof(float));
    if (ctx->stc_io_bytes < need)
    {
        cudaFree(ctx->d_stc_io); ctx->d_stc_io = nullptr; ctx->stc_io_bytes = 0;
        CK(cudaMalloc(&ctx->d_stc_io, need + need / 4));
        ctx->stc_io_bytes = need + need / 4;
    }
    uint8_t *d_cover = ctx->d_stc_io, *d_stego = ctx->d_stc_io + up(n);
    float *d_rho = (float *)(ctx->d_stc_io + 2 * up(n));
    CK(cudaMemcpyAsync(d_cover, cover, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_rho, rho, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    // the reference's total: rho of the elements the block schedule covers, summed in order in double (embed.h:405-470)
    const double invalpha = (double)n / an;
    const int shorter = (int)floor(invalpha), longer = (int)ceil(invalpha);
    int worm = 0, used = 0;
    for (int b = 0; b < an; b++)
    {
        const bool wide = worm + longer <= (b + 1) * invalpha + 0.5;
        worm += wide ? longer : shorter;
        used += wide ? longer : shorter;
    }
    if (used > n) used = n;
    double total = 0;
    for (int i = 0; i < used; i++) total += rho[i];
    const int rc = stc_run_device(ctx, d_cover, d_rho, n, message, an, matrixheight, cols_short, w_short, cols_long, w_long, total, d_stego);
    if (rc) return rc;
    CK(cudaMemcpyAsync(stego, d_stego, used, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
