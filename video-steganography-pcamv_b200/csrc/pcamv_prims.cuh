// pcamv_prims.cuh — lane-team primitives and packed-pixel arithmetic.
//
// All hot-path device code in this repo is written against a small "team" abstraction: a team is
// one warp on the GPU, split into PCAMV_NGRP groups of PCAMV_LPG lanes.  A group evaluates one
// candidate motion vector (its lanes own disjoint 8x4 / 4x4 units of the block), so four candidates
// of a diamond / hexagon step are costed concurrently and reduced with three xor-shuffles.
// Control flow (which candidate is tried next, strict-< tie-breaks) is executed redundantly and
// uniformly by every lane, which is what keeps the results bit-exact with the scalar reference.
//
// With -DPCAMV_EMU the same source compiles as plain C++ with a team of ONE lane; tests/ use that
// build to check the search logic against reference dumps on the CPU-only container.  It is a
// development/test aid only: libpcamv_cuda.so never contains or calls it.
#pragma once
#include <stdint.h>

#if defined(PCAMV_EMU)
  #include <string.h>
  #include <stdlib.h>
  #define PCAMV_DEV static inline
  #define PCAMV_MEM inline
  #define PCAMV_FN static
  #define PCAMV_MEMFN inline
  #define PCAMV_NGRP 1
  #define PCAMV_LPG  1
#else
  #include <cuda_runtime.h>
  #define PCAMV_DEV __device__ __forceinline__
  #define PCAMV_MEM __device__ __forceinline__
  // out-of-line device functions: the search code is far larger than the instruction cache when everything
  // is inlined (first profile: 678 KB of SASS, warps stalled on instruction fetch), so the block-cost
  // evaluators and the pattern walkers exist exactly once.
  #define PCAMV_FN static __device__ __noinline__
  #define PCAMV_MEMFN __device__ __noinline__
  #define PCAMV_NGRP 4
  #define PCAMV_LPG  8
#endif

namespace pcamv {

// ---- team geometry ------------------------------------------------------------------------------
PCAMV_DEV int team_lane()
{
#if defined(PCAMV_EMU)
    return 0;
#else
    return threadIdx.x & 31;
#endif
}
PCAMV_DEV int team_grp() { return team_lane() / PCAMV_LPG; }
PCAMV_DEV int team_sub() { return team_lane() % PCAMV_LPG; }

// sum over the lanes of one group; every lane of the group gets the total
PCAMV_DEV int grp_sum(int v)
{
#if !defined(PCAMV_EMU)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
#endif
    return v;
}
// sum over the whole team
PCAMV_DEV int team_sum(int v)
{
#if !defined(PCAMV_EMU)
    v = grp_sum(v);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
#endif
    return v;
}
// value held by group g (any lane of it), delivered to every lane of the team
PCAMV_DEV int grp_bcast(int v, int g)
{
#if defined(PCAMV_EMU)
    (void)g; return v;
#else
    return __shfl_sync(0xffffffffu, v, g * PCAMV_LPG);
#endif
}
PCAMV_DEV int lane_bcast(int v, int lane)
{
#if defined(PCAMV_EMU)
    (void)lane; return v;
#else
    return __shfl_sync(0xffffffffu, v, lane);
#endif
}
// smallest value over the team, in every lane
PCAMV_DEV int team_min(int v)
{
#if !defined(PCAMV_EMU)
    v = __reduce_min_sync(0xffffffffu, v);
#endif
    return v;
}
// min over the LOWER lanes of the team (INT_MAX in lane 0): the running minimum a sequential scan would hold
PCAMV_DEV int team_prefix_min_excl(int v)
{
#if defined(PCAMV_EMU)
    (void)v; return 0x7fffffff;
#else
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = o < v ? o : v;
    }
    const int up = __shfl_up_sync(0xffffffffu, v, 1);
    return lane ? up : 0x7fffffff;
#endif
}
// bit l = predicate of lane l
PCAMV_DEV unsigned team_ballot(bool p)
{
#if defined(PCAMV_EMU)
    return p ? 1u : 0u;
#else
    return __ballot_sync(0xffffffffu, p);
#endif
}
PCAMV_DEV int popc_below(unsigned mask)       // set bits of `mask` in lanes below the caller
{
#if defined(PCAMV_EMU)
    (void)mask; return 0;
#else
    return __popc(mask & ((1u << (threadIdx.x & 31)) - 1u));
#endif
}
PCAMV_DEV int ctz32(unsigned mask)            // index of the lowest set bit (mask != 0)
{
#if defined(PCAMV_EMU)
    return __builtin_ctz(mask);
#else
    return __ffs((int)mask) - 1;
#endif
}
PCAMV_DEV int popc32(unsigned mask)
{
#if defined(PCAMV_EMU)
    return __builtin_popcount(mask);
#else
    return __popc(mask);
#endif
}
PCAMV_DEV void team_sync()
{
#if !defined(PCAMV_EMU)
    __syncwarp();
#endif
}

// ---- packed pixel arithmetic ---------------------------------------------------------------------
// unaligned 4-pixel load (little endian: pixel x in bits 0..7)
PCAMV_DEV uint32_t ld4(const uint8_t *p)
{
#if defined(PCAMV_EMU)
    uint32_t v; memcpy(&v, p, 4); return v;
#else
    uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t lo = __ldg(q), hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, (unsigned)(a & 3) * 8);
#endif
}
// 4-pixel load from a 4-byte aligned address (fenc / recon staging buffers)
PCAMV_DEV uint32_t ld4a(const uint8_t *p)
{
#if defined(PCAMV_EMU)
    uint32_t v; memcpy(&v, p, 4); return v;
#else
    return *(const uint32_t *)p;
#endif
}
// The block being matched (source pixels, or the reconstruction in the cost table) always sits in shared memory on the
// GPU.  Inside the out-of-line evaluators — where that memory is read-only — it is addressed as such: LDS needs no memory
// descriptor, while every generic / global load in an out-of-line function costs two extra R2UR instructions.
#if defined(PCAMV_EMU)
typedef const uint8_t *smem_ptr;
PCAMV_DEV smem_ptr to_smem(const uint8_t *p) { return p; }
PCAMV_DEV uint32_t ld4s(smem_ptr p) { uint32_t v; memcpy(&v, p, 4); return v; }
#else
typedef unsigned smem_ptr;
PCAMV_DEV smem_ptr to_smem(const uint8_t *p) { return (unsigned)__cvta_generic_to_shared(p); }
PCAMV_DEV uint32_t ld4s(smem_ptr p)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p));
    return v;
}
#endif
// per-byte (a+b+1)>>1 : the quarter-pel average of pixel_avg (reference common/mc.c:34-50)
PCAMV_DEV uint32_t avg4(uint32_t a, uint32_t b)
{
#if defined(PCAMV_EMU)
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        r |= ((((a >> (8 * i)) & 255) + ((b >> (8 * i)) & 255) + 1) >> 1) << (8 * i);
    return r;
#else
    return __vavgu4(a, b);
#endif
}
// sum over 4 bytes of |a-b| (reference common/pixel.c:40-55)
PCAMV_DEV int sad4(uint32_t a, uint32_t b)
{
#if defined(PCAMV_EMU)
    int s = 0;
    for (int i = 0; i < 4; i++)
        s += abs((int)((a >> (8 * i)) & 255) - (int)((b >> (8 * i)) & 255));
    return s;
#else
    return (int)__vsadu4(a, b);
#endif
}
PCAMV_DEV int px(uint32_t w, int i) { return (int)((w >> (8 * i)) & 255); }

// |x| + (|y| << 16) of the pseudo-SIMD pair x + (y << 16)   (reference common/pixel.c:177-181)
PCAMV_DEV uint32_t abs2(uint32_t a)
{
    uint32_t s = ((a >> 15) & 0x10001u) * 0xffffu;
    return (a + s) ^ s;
}

PCAMV_DEV int imax_abs(int a, int b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; return a > b ? a : b; }

#define PCAMV_HADAMARD4(d0, d1, d2, d3, s0, s1, s2, s3) \
    { uint32_t t0 = (s0) + (s1), t1 = (s0) - (s1), t2 = (s2) + (s3), t3 = (s2) - (s3); \
      d0 = t0 + t2; d2 = t0 - t2; d1 = t1 + t3; d3 = t1 - t3; }

// Sum of |4x4 Hadamard coefficients| over the two 4x4 halves of an 8x4 block, NOT yet halved.
// f[r], g[r] = left / right 4 pixels of fenc row r; a[r], b[r] = same for the prediction.
// Same packing as the reference's satd_8x4 (common/pixel.c:209-228): left half in bits 0..15,
// right half in bits 16..31 of each lane word.
PCAMV_DEV uint32_t hadamard_8x4_sum(const uint32_t f[4], const uint32_t g[4], const uint32_t a[4], const uint32_t b[4])
{
    uint32_t tmp[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        uint32_t a0 = (uint32_t)(px(f[r], 0) - px(a[r], 0)) + ((uint32_t)(px(g[r], 0) - px(b[r], 0)) << 16);
        uint32_t a1 = (uint32_t)(px(f[r], 1) - px(a[r], 1)) + ((uint32_t)(px(g[r], 1) - px(b[r], 1)) << 16);
        uint32_t a2 = (uint32_t)(px(f[r], 2) - px(a[r], 2)) + ((uint32_t)(px(g[r], 2) - px(b[r], 2)) << 16);
        uint32_t a3 = (uint32_t)(px(f[r], 3) - px(a[r], 3)) + ((uint32_t)(px(g[r], 3) - px(b[r], 3)) << 16);
        PCAMV_HADAMARD4(tmp[r][0], tmp[r][1], tmp[r][2], tmp[r][3], a0, a1, a2, a3);
    }
    uint32_t sum = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        uint32_t a0, a1, a2, a3;
        PCAMV_HADAMARD4(a0, a1, a2, a3, tmp[0][c], tmp[1][c], tmp[2][c], tmp[3][c]);
        sum += abs2(a0) + abs2(a1) + abs2(a2) + abs2(a3);
    }
    return (sum & 0xffffu) + (sum >> 16);
}

// Sum of |4x4 Hadamard coefficients| of one 4x4 block, NOT yet halved (common/pixel.c:187-207).
// Two differences per word throughout: the even and the odd bytes of a row are split with one mask each, so one
// subtraction yields (d0, d2) and one (d1, d3) in the x + (y << 16) representation the reference's own packing uses;
// sums and differences of such words are the packed sums and differences as long as every half stays within 16 bits
// (|coefficient| <= 16 * 255).  The first horizontal butterfly and the whole vertical transform run on packed words; the
// last horizontal butterfly is never formed: |a + b| + |a - b| = 2 * max(|a|, |b|).
PCAMV_DEV int pk_lo(uint32_t p) { return (int)(int16_t)(p & 0xffffu); }
PCAMV_DEV int pk_hi(uint32_t p) { return ((int)p - pk_lo(p)) >> 16; }
PCAMV_DEV uint32_t hadamard_4x4_sum(const uint32_t f[4], const uint32_t a[4])
{
    const uint32_t m = 0x00ff00ffu;
    uint32_t x[4], y[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const uint32_t de = (f[r] & m) - (a[r] & m);                   // (d0, d2)
        const uint32_t dq = ((f[r] >> 8) & m) - ((a[r] >> 8) & m);     // (d1, d3)
        x[r] = de + dq;                                               // (d0 + d1, d2 + d3)
        y[r] = de - dq;                                               // (d0 - d1, d2 - d3)
    }
    uint32_t xv[4], yv[4];
    PCAMV_HADAMARD4(xv[0], xv[1], xv[2], xv[3], x[0], x[1], x[2], x[3]);
    PCAMV_HADAMARD4(yv[0], yv[1], yv[2], yv[3], y[0], y[1], y[2], y[3]);
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        sum += imax_abs(pk_lo(xv[k]), pk_hi(xv[k]));
        sum += imax_abs(pk_lo(yv[k]), pk_hi(yv[k]));
    }
    return (uint32_t)(2 * sum);
}

PCAMV_DEV int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
PCAMV_DEV int imin(int a, int b) { return a < b ? a : b; }
PCAMV_DEV int imax(int a, int b) { return a > b ? a : b; }
PCAMV_DEV int iabs(int a) { return a < 0 ? -a : a; }
PCAMV_DEV int median3(int a, int b, int c) { return imax(imin(a, b), imin(imax(a, b), c)); }
PCAMV_DEV int clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }

} // namespace pcamv
