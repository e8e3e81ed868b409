// pcamv_me.cuh — block costs (SAD / SATD / qpel / chroma MC) and the motion search of one block.
//
// Behavioural contract (bit-exact): reference encoder/me.c:158-666 (x264_me_search_ref),
// :669-678 (x264_me_refine_qpel), :715-843 (refine_subpel); common/pixel.c:40-65,187-253
// (SAD, SATD); common/mc.c:194-277 (get_ref quarter-pel averaging, mc_chroma).
// The structure is NOT the reference's: candidates of one pattern step are costed concurrently by
// the groups of a lane team (pcamv_prims.cuh) and then folded in the reference's evaluation order
// with its strict-< rule, so the first minimum wins exactly as in the scalar code.
#pragma once
#include "pcamv_prims.cuh"
#if defined(PCAMV_CHECKED) && !defined(PCAMV_EMU)
#include <stdio.h>
#endif

namespace pcamv {

enum { PIX_16x16 = 0, PIX_16x8, PIX_8x16, PIX_8x8, PIX_8x4, PIX_4x8, PIX_4x4 };
enum { ME_DIA = 0, ME_HEX, ME_UMH, ME_ESA, ME_TESA };
#define PCAMV_COST_MAX (1 << 28)

PCAMV_DEV int pix_w(int i_pixel) { return i_pixel <= PIX_16x8 ? 16 : i_pixel <= PIX_8x4 ? 8 : 4; }
PCAMV_DEV int pix_h(int i_pixel)
{
    return (i_pixel == PIX_16x16 || i_pixel == PIX_8x16) ? 16 : (i_pixel == PIX_8x4 || i_pixel == PIX_4x4) ? 4 : 8;
}

// Encoder-wide search parameters (one per context / QP).
struct MeEnv
{
    const int16_t *cost_mv;          // centre of the lambda*bits table, valid index [-16384, 16384]
    const uint16_t *cost_mv_fpel[4]; // ESA: cost_mv[i*4+j] tables, centre pointers (may be null)
    int me_method, me_range, subme, chroma_me;
    int mbcmp_satd;                  // mbcmp is SATD when subme > 1, else SAD (reference encoder/encoder.c:615-625)
    int mv_min_fpel[2], mv_max_fpel[2];   // per-MB limits (reference encoder/analyse.c:279-318)
    int mv_min_spel[2], mv_max_spel[2];
    unsigned long long *mvsads;      // --me tesa: candidate list of the team ((2*range+4) * (2*range+1) entries), else null
};

// One block to be searched: fenc pixels, the 4 luma planes + 2 chroma planes of ONE reference,
// already offset to the block's position, and the MV-cost tables biased by the predictor.
struct MeBlock
{
    const uint8_t *fenc;      // luma, row stride 16
    const uint8_t *fenc_u;    // chroma, row stride 8
    const uint8_t *fenc_v;
    const uint8_t *ref[4];    // integer, H, V, HV planes at the block origin
    const uint8_t *ref_u, *ref_v;
    const uint16_t *integral, *integral4; // exhaustive searches only: 8x8 / 4x4 box sums of the reference at the block origin
    int stride, stride_c;
    int i_pixel, bw, bh;
    int mvp[2];
    const int16_t *cost_mvx, *cost_mvy;   // cost_mv - mvp[0], cost_mv - mvp[1]
    int k_fpel, k_qsad;       // what fpelcmp is at integer / quarter-pel positions: SAD, or SATD with --me tesa (encoder/encoder.c:621-624)
#if defined(PCAMV_CHECKED)
    const uint8_t *chk_lo, *chk_hi;    // the reference slot's allocation: every pixel load of the evaluators must stay inside
#endif
};

// PCAMV_CHECKED build (tools/checked_build.sh): the address range a call is about to read from the reference planes is checked
// against the slot's allocation, and the kernel traps on a violation — the pool this project is measured on refuses
// compute-sanitizer (profiles/r02_sanitizer_attempt.txt), so the out-of-bounds class that matters here (a motion vector, a
// border or a stride computed wrongly sends the evaluator outside the padded planes) is checked by the code itself.
#if defined(PCAMV_CHECKED) && !defined(PCAMV_EMU)
  #define PCAMV_CHK_RANGE(b, first, last, what) do { \
      if ((const uint8_t *)(first) < (b).chk_lo || (const uint8_t *)(last) > (b).chk_hi) { \
          if ((threadIdx.x & 31) == 0) printf("PCAMV_CHECKED: %s reads [%p, %p) outside the reference slot [%p, %p)\n", what, (const void *)(first), (const void *)(last), (const void *)(b).chk_lo, (const void *)(b).chk_hi); \
          __trap(); } } while (0)
#else
  #define PCAMV_CHK_RANGE(b, first, last, what) do { } while (0)
#endif

// search state / result (the in/out part of the reference's x264_me_t, encoder/me.h:30-51)
struct MeResult
{
    int mv[2];
    int cost;
    int cost_mv;
};

PCAMV_DEV void block_set_mvp(MeBlock &b, const MeEnv &env, int mvpx, int mvpy)
{
    b.mvp[0] = mvpx; b.mvp[1] = mvpy;
    b.cost_mvx = env.cost_mv - mvpx;
    b.cost_mvy = env.cost_mv - mvpy;
    const int tesa = env.me_method == 4 && env.mbcmp_satd;
    b.k_fpel = tesa ? 1 : 3;        // COST_SATD : COST_SAD_FPEL
    b.k_qsad = tesa ? 1 : 0;        // COST_SATD : COST_SAD
}

// ---- quarter-pel reference addressing (reference common/mc.c:192-243) ---------------------------
PCAMV_DEV void qpel_sources(const MeBlock &b, int qmx, int qmy, const uint8_t *&s1, const uint8_t *&s2)
{
    const int fx = qmx & 3, fy = qmy & 3;
    const int idx = (fy << 2) + fx;
    // first source: integer/H plane, or V/HV plane on the half-pel row (table hpel_ref0)
    const int h0 = (fy == 2 ? ((fx == 0) ? 2 : 3) : ((fx == 0) ? 0 : 1));
    // second source (table hpel_ref1): V plane off the integer row, HV plane in the half-pel column
    const int h1 = (fy == 0) ? 0 : ((fx == 2) ? 3 : 2);
    const int off = (qmy >> 2) * b.stride + (qmx >> 2);
    s1 = b.ref[h0] + off + (fy == 3 ? b.stride : 0);
    s2 = (idx & 5) ? b.ref[h1] + off + (fx == 3 ? 1 : 0) : s1;      // (a + a + 1) >> 1 == a: no branch for the unaveraged positions
}

// 4 predicted luma pixels at block-relative (x, y) for sources prepared by qpel_sources
PCAMV_DEV uint32_t pred4(const uint8_t *s1, const uint8_t *s2, int stride, int x, int y)
{
    return avg4(ld4(s1 + y * stride + x), ld4(s2 + y * stride + x));
}

// =====================================================================================================
// Candidate costs.  One out-of-line evaluator serves every caller (integer search, predictor test, half/quarter-
// pel refinement, the cost-table rings): the instruction cache is the scarce resource of the per-macroblock
// code, so there is exactly ONE copy of the SAD loop, the SATD unit loop and the 4x4 Hadamard in a kernel.
// Candidates travel in registers (packed x | y << 16), never through local-memory arrays.
// =====================================================================================================
enum { COST_SAD = 0, COST_SATD = 1, COST_SATD_CHROMA = 2, COST_SAD_FPEL = 3 };

PCAMV_DEV int pk(int x, int y) { return (int)(((uint32_t)x & 0xffffu) | ((uint32_t)y << 16)); }
PCAMV_DEV int pk_x(int p) { return (int)(int16_t)(p & 0xffff); }
PCAMV_DEV int pk_y(int p) { return p >> 16; }

// 4 chroma pixels of the bilinear 1/8-pel interpolation (reference common/mc.c:246-277) from the two source rows
// (t = top row words at s and s+1, u = bottom row), two pixels per 32-bit multiply-add: every 16-bit half stays
// below 64*255+32, so the halves never interact and the result equals the scalar formula bit for bit.
PCAMV_DEV uint32_t bilin4(uint32_t t0, uint32_t t1, uint32_t u0, uint32_t u1, int cA, int cB, int cC, int cD)
{
    const uint32_t m = 0x00ff00ffu;
    const uint32_t e = (t0 & m) * cA + (t1 & m) * cB + (u0 & m) * cC + (u1 & m) * cD + 0x00200020u;
    const uint32_t o = ((t0 >> 8) & m) * cA + ((t1 >> 8) & m) * cB + ((u0 >> 8) & m) * cC + ((u1 >> 8) & m) * cD + 0x00200020u;
    return ((e >> 6) & m) | (((o >> 6) & m) << 8);
}
// words at p and p+1 (unaligned) from one pair of aligned loads
PCAMV_DEV void ld4x2(const uint8_t *p, uint32_t &w0, uint32_t &w1)
{
#if defined(PCAMV_EMU)
    memcpy(&w0, p, 4); memcpy(&w1, p + 1, 4);
#else
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned long long v = (unsigned long long)__ldg(q) | ((unsigned long long)__ldg(q + 1) << 32);
    const unsigned sh = (unsigned)(a & 3) * 8;
    w0 = (uint32_t)(v >> sh);
    w1 = (uint32_t)(v >> (sh + 8));
#endif
}

// 4 predicted chroma pixels at block-relative (x, y) (reference common/mc.c:246-277), src already at the block origin
PCAMV_DEV uint32_t chroma4(const uint8_t *src, int stride, int qmx, int qmy, int x, int y)
{
    const int dx = qmx & 7, dy = qmy & 7;
    const uint8_t *s = src + ((qmy >> 3) + y) * stride + (qmx >> 3) + x;
    uint32_t t0, t1, u0, u1;
    ld4x2(s, t0, t1);
    ld4x2(s + stride, u0, u1);
    return bilin4(t0, t1, u0, u1, (8 - dx) * (8 - dy), dx * (8 - dy), (8 - dx) * dy, dx * dy);
}

// 16 pixels of the unaligned row at p as four words (five aligned loads + funnel shifts)
PCAMV_DEV void ld_row16(const uint8_t *p, uint32_t w[4])
{
#if defined(PCAMV_EMU)
    memcpy(w, p, 16);
#else
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t v[5];
#pragma unroll
    for (int j = 0; j < 5; j++) v[j] = __ldg(q + j);
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = __funnelshift_r(v[j], v[j + 1], sh);
#endif
}

#if defined(PCAMV_EMU)
  #define PCAMV_ROWS 16      // rows of a block handled by one lane
#else
  #define PCAMV_ROWS 2
#endif

// Cost of candidate `grp` of (c0..c3) — distortion of the block against the quarter-pel prediction at the packed MV
// plus the MV bit cost — returned in every lane of that group; groups >= n return COST_MAX (they still compute, on
// whatever valid candidate they were handed, so that the whole team runs one straight-line instruction stream:
// a lone warp per macroblock pays ~25 cycles for every divergent branch and reconvergence).
// kind: COST_SAD_FPEL (fpelcmp at integer positions: one source plane), COST_SAD (fpelcmp at any quarter-pel position),
// COST_SATD (mbcmp, luma), COST_SATD_CHROMA (luma + both chroma planes, each halved separately as in the reference's
// three calls; every 4x4 Hadamard sum is even, so halving per unit is exact).
// SAD: a lane owns whole rows (sub, sub + 8) of up to 16 pixels; SATD: a lane owns 4x4 units.
PCAMV_FN int cand_cost(const MeBlock &b, int kind, int n, int c0, int c1, int c2, int c3)
{
    // a single candidate gets the whole team (32 lanes) instead of one group of 8
#if defined(PCAMV_EMU)
    const int wide = 0, lpg = 1, grp = 0, sub = 0;
#else
    const int wide = n == 1;
    const int lpg = wide ? 32 : PCAMV_LPG;
    const int grp = wide ? 0 : team_grp(), sub = wide ? team_lane() : team_sub();
#endif
    const int c = grp == 0 ? c0 : grp == 1 ? c1 : grp == 2 ? c2 : c3;
    const int qx = pk_x(c), qy = pk_y(c);
    int acc = 0;
    // the block descriptor lives in memory the compiler cannot prove unaliased: read each field once
    const smem_ptr fenc = to_smem(b.fenc);
    const int stride = b.stride, bw = b.bw, bh = b.bh;
    const uint8_t *s1, *s2;
    if (kind == COST_SAD_FPEL)
        s1 = s2 = b.ref[0] + (qy >> 2) * stride + (qx >> 2);          // integer position: the plane itself
    else
    {
        const int fx = qx & 3, fy = qy & 3;
        const int h0 = (fy == 2 ? ((fx == 0) ? 2 : 3) : ((fx == 0) ? 0 : 1));
        const int h1 = (fy == 0) ? 0 : ((fx == 2) ? 3 : 2);
        const int off = (qy >> 2) * stride + (qx >> 2);
        s1 = b.ref[h0] + off + (fy == 3 ? stride : 0);
        s2 = (((fy << 2) + fx) & 5) ? b.ref[h1] + off + (fx == 3 ? 1 : 0) : s1;
    }
    // (ld_row16 / ld4 read whole aligned words: up to 3 bytes before and 7 after the pixels they deliver)
    PCAMV_CHK_RANGE(b, (s1 < s2 ? s1 : s2) - 3, (s1 > s2 ? s1 : s2) + (bh - 1) * stride + 16 + 7, "cand_cost (luma)");
#ifdef PCAMV_CHECKED
    if (kind == COST_SATD_CHROMA)
    {
        const uint8_t *cu = b.ref_u + (qy >> 3) * b.stride_c + (qx >> 3), *cv = b.ref_v + (qy >> 3) * b.stride_c + (qx >> 3);
        PCAMV_CHK_RANGE(b, cu - 3, cu + (bh >> 1) * b.stride_c + (bw >> 1) + 1 + 7, "cand_cost (U)");
        PCAMV_CHK_RANGE(b, cv - 3, cv + (bh >> 1) * b.stride_c + (bw >> 1) + 1 + 7, "cand_cost (V)");
    }
#endif
    if (kind == COST_SAD_FPEL || kind == COST_SAD)
    {
        const int w4 = bw >> 2;
#pragma unroll 1
        for (int r = 0; r < PCAMV_ROWS && lpg * r < bh; r++)
        {
            const int y = sub + lpg * r;
            const int yy = imin(y, bh - 1);
            uint32_t p[4];
            ld_row16(s1 + yy * stride, p);
            if (kind == COST_SAD)
            {
                uint32_t q[4];
                ld_row16(s2 + yy * stride, q);
#pragma unroll
                for (int j = 0; j < 4; j++) p[j] = avg4(p[j], q[j]);
            }
            int s = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                s += j < w4 ? sad4(ld4s(fenc + yy * 16 + 4 * j), p[j]) : 0;
            acc += y < bh ? s : 0;
        }
    }
    else
    {
        // 4x4 units: nl luma units, then nc units of U, then nc units of V; one Hadamard per unit
        const int llx = bw == 16 ? 2 : bw == 8 ? 1 : 0;         // log2(luma units per row)
        const int nl = (bh >> 2) << llx;
        const int lcx = bw == 16 ? 1 : 0;                        // log2(chroma units per row)
        const int nc = kind == COST_SATD_CHROMA ? (bh >> 3) << lcx : 0;
        const int stride_c = b.stride_c;
        const int total = nl + 2 * nc;
        const int dx = qx & 7, dy = qy & 7;
        const int cA = (8 - dx) * (8 - dy), cB = dx * (8 - dy), cC = (8 - dx) * dy, cD = dx * dy;
#pragma unroll 1
        for (int k = 0; k * lpg < total; k++)
        {
            const int u = sub + lpg * k;
            const int uu = u < total ? u : 0;
            uint32_t f[4], a[4];
            if (uu < nl)
            {
                const int x = (uu & ((1 << llx) - 1)) << 2, y = (uu >> llx) << 2;
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    f[r] = ld4s(fenc + (y + r) * 16 + x);
                    a[r] = pred4(s1, s2, stride, x, y + r);
                }
            }
            else
            {
                int v = uu - nl;
                const int pl = v >= nc;
                v -= pl ? nc : 0;
                const int x = (v & ((1 << lcx) - 1)) << 2, y = (v >> lcx) << 2;
                const smem_ptr fe = to_smem(pl ? b.fenc_v : b.fenc_u) + y * 8 + x;
                const uint8_t *s = (pl ? b.ref_v : b.ref_u) + ((qy >> 3) + y) * stride_c + (qx >> 3) + x;
                uint32_t t0, t1;
                ld4x2(s, t0, t1);
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    uint32_t u0, u1;
                    ld4x2(s + (r + 1) * stride_c, u0, u1);
                    f[r] = ld4s(fe + r * 8);
                    a[r] = bilin4(t0, t1, u0, u1, cA, cB, cC, cD);
                    t0 = u0; t1 = u1;
                }
            }
            const int h = (int)(hadamard_4x4_sum(f, a) >> 1);
            acc += u < total ? h : 0;
        }
    }
#if !defined(PCAMV_EMU)
    {
        // group totals, then (always, branch-free) the team total for the single-candidate layout
        // (three shuffles beat one REDUX here: a lone warp waits out the reduction's latency, measured 7 % per pass;
        // summing the four group totals from eval4's broadcasts instead of the two extra shuffles: measured no faster)
        acc = grp_sum(acc);
        int all = acc + __shfl_xor_sync(0xffffffffu, acc, 8);
        all += __shfl_xor_sync(0xffffffffu, all, 16);
        acc = wide ? all : acc;
    }
#endif
    return grp < n ? acc + b.cost_mvx[qx] + b.cost_mvy[qy] : PCAMV_COST_MAX;
}

// costs[i] = cost of candidate i (i < n <= 4) in EVERY lane
#if defined(PCAMV_EMU)
static long g_eval_calls[4][5][7];      // [kind][n][i_pixel]: evaluator calls made, for the candidate-mix statistics of tests/emu
#endif
PCAMV_DEV void eval4(const MeBlock &b, int kind, int n, int c0, int c1, int c2, int c3, int costs[4])
{
#if defined(PCAMV_EMU)
    g_eval_calls[kind][n][b.i_pixel]++;
    const int c[4] = { c0, c1, c2, c3 };
    for (int i = 0; i < 4; i++)
        costs[i] = i < n ? cand_cost(b, kind, 1, c[i], c[i], c[i], c[i]) : PCAMV_COST_MAX;
#else
    const int mine = cand_cost(b, kind, n, c0, c1, c2, c3);
#pragma unroll
    for (int g = 0; g < 4; g++)
        costs[g] = grp_bcast(mine, g);
#endif
}
PCAMV_DEV int eval1(const MeBlock &b, int kind, int c)
{
    int costs[4];
    eval4(b, kind, 1, c, c, c, c, costs);
    return costs[0];
}

// =====================================================================================================
// Integer-pel search + sub-pel refinement of one block
// =====================================================================================================
// running best of the integer search, in registers: cost in the high word, packed full-pel MV in the low word
typedef unsigned long long best_t;
PCAMV_DEV best_t best_make(int cost, int mv) { return ((best_t)(uint32_t)cost << 32) | (uint32_t)mv; }
PCAMV_DEV int best_cost(best_t s) { return (int)(s >> 32); }
PCAMV_DEV int best_mv(best_t s) { return (int)(uint32_t)s; }

// four signed byte offsets in one word
#define PCAMV_OFF4(a, b, c, d) ((uint32_t)((a) & 0xff) | ((uint32_t)((b) & 0xff) << 8) | ((uint32_t)((c) & 0xff) << 16) | ((uint32_t)((d) & 0xff) << 24))
PCAMV_DEV int off_at(uint32_t offs, int i) { return (int)(int8_t)(offs >> (8 * i)); }

// Cost the full-pel candidates (ox + dx_i, oy + dy_i), i < n, and fold them into the best in order with the
// reference's strict-< rule (COST_MV / COST_MV_X3_DIR / COST_MV_X4, encoder/me.c:57-120).
PCAMV_FN best_t try4(const MeBlock &b, best_t best, int n, int ox, int oy, uint32_t dxs, uint32_t dys)
{
    int mv[4], costs[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        mv[i] = pk(ox + off_at(dxs, i), oy + off_at(dys, i));
    eval4(b, b.k_fpel, n, pk(pk_x(mv[0]) << 2, pk_y(mv[0]) << 2), pk(pk_x(mv[1]) << 2, pk_y(mv[1]) << 2),
          pk(pk_x(mv[2]) << 2, pk_y(mv[2]) << 2), pk(pk_x(mv[3]) << 2, pk_y(mv[3]) << 2), costs);
    int bcost = best_cost(best), bmv = best_mv(best);
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < n && costs[i] < bcost) { bcost = costs[i]; bmv = mv[i]; }
    return best_make(bcost, bmv);
}
PCAMV_DEV best_t try1(const MeBlock &b, best_t best, int mx, int my) { return try4(b, best, 1, mx, my, 0, 0); }
PCAMV_DEV best_t dia1(const MeBlock &b, best_t best, int ox, int oy)
{
    return try4(b, best, 4, ox, oy, PCAMV_OFF4(0, 0, -1, 1), PCAMV_OFF4(-1, 1, 0, 0));
}
PCAMV_DEV bool fpel_in_range(const MeEnv &e, int mx, int my)
{
    return mx >= e.mv_min_fpel[0] && mx <= e.mv_max_fpel[0] && my >= e.mv_min_fpel[1] && my <= e.mv_max_fpel[1];
}

// ---- one candidate per LANE --------------------------------------------------------------------------
// Candidate sets that do not depend on intermediate results (the rings of the UMH hexagon grid, the rows of the exhaustive
// search) are costed 32 at a time: every lane computes the whole SAD of its own full-pel position, so there is no
// cross-lane reduction per candidate, and the 32 results are folded in evaluation order (first minimum wins, strict <
// against the running best — the same outcome as the reference's sequential COST_MV loop).
#if defined(PCAMV_EMU)
  #define PCAMV_WIDE 1
#else
  #define PCAMV_WIDE 32
#endif

// SAD + MV cost of the block at full-pel (mx, my), computed by the calling lane alone
PCAMV_FN int lane_sad(const MeBlock &b, int mx, int my)
{
    const smem_ptr fenc = to_smem(b.fenc);
    const int stride = b.stride, bh = b.bh, w4 = b.bw >> 2;
    const uint8_t *s = b.ref[0] + my * stride + mx;
    PCAMV_CHK_RANGE(b, s - 3, s + (bh - 1) * stride + 16 + 7, "lane_sad");
    int acc = 0;
    // (not unrolled: measured no faster with four rows in flight, and the instruction cache is the scarcer resource)
#pragma unroll 1
    for (int y = 0; y < bh; y++)
    {
        uint32_t p[4];
        ld_row16(s + y * stride, p);
#pragma unroll
        for (int j = 0; j < 4; j++)
            acc += j < w4 ? sad4(ld4s(fenc + y * 16 + 4 * j), p[j]) : 0;
    }
    return acc + b.cost_mvx[mx << 2] + b.cost_mvy[my << 2];
}

// fold the lanes' (cost, mv) of candidates k0 + lane (cost = COST_MAX for lanes without a candidate) into the best
PCAMV_FN best_t fold_wide(best_t best, int cost, int mv)
{
#if !defined(PCAMV_EMU)
    // first minimum in lane order: the smallest cost (one REDUX), then the lowest lane holding it
    const int m = __reduce_min_sync(0xffffffffu, cost);
    const unsigned who = __ballot_sync(0xffffffffu, cost == m);
    mv = __shfl_sync(0xffffffffu, mv, __ffs((int)who) - 1);
    cost = m;
#endif
    return cost < best_cost(best) ? best_make(cost, mv) : best;
}

// symmetric cross around (ox,oy)  (reference encoder/me.c:131-155)
PCAMV_FN best_t search_cross(const MeEnv &e, const MeBlock &b, best_t best, int ox, int oy, int start, int x_max, int y_max)
{
    // evaluation order of the reference: (+i,0) (-i,0) for i = start, start+2, .. < x_max, then (0,+i) (0,-i) for i < y_max;
    // a point outside the MV range is skipped (the reference's unchecked 4-at-a-time loops only run when none can be).
    // The centre is fixed, so the points are independent: one per lane.
    const int nx = x_max > start ? 2 * ((x_max - start + 1) >> 1) : 0;
    const int ny = y_max > start ? 2 * ((y_max - start + 1) >> 1) : 0;
#pragma unroll 1
    for (int k0 = 0; k0 < nx + ny; k0 += PCAMV_WIDE)
    {
        const int k = k0 + team_lane();
        const int vert = k >= nx;
        const int kk = vert ? k - nx : k;
        const int i = start + 2 * (kk >> 1);
        const int d = (kk & 1) ? -i : i;
        const int mx = vert ? ox : ox + d, my = vert ? oy + d : oy;
        const bool valid = k < nx + ny && (vert ? (my <= e.mv_max_fpel[1] && my >= e.mv_min_fpel[1])
                                                : (mx <= e.mv_max_fpel[0] && mx >= e.mv_min_fpel[0]));
        const int cost = valid ? lane_sad(b, mx, my) : PCAMV_COST_MAX;
        best = fold_wide(best, cost, pk(mx, my));
    }
    return best;
}

// hexagon corner k = 0..5: (-2,0) (-1,2) (1,2) (2,0) (1,-2) (-1,-2), one signed nibble per corner
PCAMV_DEV int hex_dx(int k) { return ((int)(0xF121FEu << (28 - 4 * k))) >> 28; }     // nibbles, k = 0 lowest: E F 1 2 1 F
PCAMV_DEV int hex_dy(int k) { return ((int)(0xEE0220u << (28 - 4 * k))) >> 28; }      // 0 2 2 0 E E

// hexagon (radius 2) walk followed by the 8-point square  (reference encoder/me.c:263-341)
PCAMV_FN best_t search_hex(const MeEnv &e, const MeBlock &b, best_t best, int me_range)
{
    int cx = pk_x(best_mv(best)), cy = pk_y(best_mv(best));
    // first step: all six corners in the order 0..5
    const int cost0 = best_cost(best);
    best = try4(b, best, 3, cx, cy, PCAMV_OFF4(-2, -1, 1, 0), PCAMV_OFF4(0, 2, 2, 0));
    best = try4(b, best, 3, cx, cy, PCAMV_OFF4(2, 1, -1, 0), PCAMV_OFF4(0, -2, -2, 0));
    if (best_cost(best) < cost0)
    {
        int nx = pk_x(best_mv(best)), ny = pk_y(best_mv(best));
        // direction index of the winning corner
        int dir = 0;
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (nx - cx == hex_dx(k) && ny - cy == hex_dy(k)) dir = k;
        cx = nx; cy = ny;
        // half hexagons: only the three corners not covered by the previous position
#pragma unroll 1
        for (int i = 1; i < me_range / 2 && fpel_in_range(e, cx, cy); i++)
        {
            const int d0 = dir == 0 ? 5 : dir - 1, d2 = dir == 5 ? 0 : dir + 1;
            const int costi = best_cost(best);
            best = try4(b, best, 3, cx, cy, PCAMV_OFF4(hex_dx(d0), hex_dx(dir), hex_dx(d2), 0),
                        PCAMV_OFF4(hex_dy(d0), hex_dy(dir), hex_dy(d2), 0));
            if (!(best_cost(best) < costi))
                break;
            nx = pk_x(best_mv(best)); ny = pk_y(best_mv(best));
            dir = (nx - cx == hex_dx(d0) && ny - cy == hex_dy(d0)) ? d0 : (nx - cx == hex_dx(dir) && ny - cy == hex_dy(dir)) ? dir : d2;
            cx = nx; cy = ny;
        }
    }
    // square refine around the final centre
    best = try4(b, best, 4, cx, cy, PCAMV_OFF4(0, 0, -1, 1), PCAMV_OFF4(-1, 1, 0, 0));
    best = try4(b, best, 4, cx, cy, PCAMV_OFF4(-1, -1, 1, 1), PCAMV_OFF4(-1, 1, -1, 1));
    return best;
}

PCAMV_DEV best_t search_dia(const MeEnv &e, const MeBlock &b, best_t best, int me_range)
{
    int i = 0;
    do
    {
        const int omv = best_mv(best);
        best = dia1(b, best, pk_x(omv), pk_y(omv));
        if (best_mv(best) == omv) break;
        if (!fpel_in_range(e, pk_x(best_mv(best)), pk_y(best_mv(best)))) break;
    } while (++i < me_range);
    return best;
}

// uneven-cross multi-hexagon-grid search (reference encoder/me.c:342-482); returns with *run_hex = 1 when the
// reference goes on to the hexagon refinement (executed by the caller, so that search_hex is expanded once)
PCAMV_FN best_t search_umh(const MeEnv &e, const MeBlock &b, best_t best, int pmx, int pmy, const int (*mvc)[2], int i_mvc,
                            int *run_hex, int *hex_range)
{
    int me_range = e.me_range;
    const int shift = b.i_pixel == PIX_16x16 ? 0 : b.i_pixel <= PIX_8x16 ? 1 : b.i_pixel == PIX_8x8 ? 2
                    : b.i_pixel <= PIX_4x8 ? 3 : 4;
    int cross_start = 1;
    *run_hex = 0; *hex_range = me_range;
    const int ucost1 = best_cost(best);
    best = dia1(b, best, pmx, pmy);
    if (pmx | pmy)
        best = dia1(b, best, 0, 0);
    if (b.i_pixel == PIX_4x4) { *run_hex = 1; return best; }

    const int ucost2 = best_cost(best);
    {
        const int bmx = pk_x(best_mv(best)), bmy = pk_y(best_mv(best));
        if ((bmx | bmy) && ((bmx - pmx) | (bmy - pmy)))
            best = dia1(b, best, bmx, bmy);
    }
    if (best_cost(best) == ucost2)
        cross_start = 3;
    int ox = pk_x(best_mv(best)), oy = pk_y(best_mv(best));

    if (best_cost(best) == ucost2 && best_cost(best) < (2000 >> shift))
    {
        best = try4(b, best, 4, ox, oy, PCAMV_OFF4(0, -1, 1, -2), PCAMV_OFF4(-2, -1, -1, 0));
        best = try4(b, best, 4, ox, oy, PCAMV_OFF4(2, -1, 1, 0), PCAMV_OFF4(0, 1, 1, 2));
        if (best_cost(best) == ucost1 && best_cost(best) < (500 >> shift))
            return best;
        if (best_cost(best) == ucost2)
        {
            const int range = (me_range >> 1) | 1;
            best = search_cross(e, b, best, ox, oy, 3, range, range);
            best = try4(b, best, 4, ox, oy, PCAMV_OFF4(-1, 1, -2, 2), PCAMV_OFF4(-2, -2, -1, -1));
            best = try4(b, best, 4, ox, oy, PCAMV_OFF4(-2, 2, -1, 1), PCAMV_OFF4(1, 1, 2, 2));
            if (best_cost(best) == ucost2)
                return best;
            cross_start = range + 2;
        }
    }

    // adaptive search range from predictor agreement
    if (i_mvc)
    {
        int mvd, denom = 1;
        if (i_mvc == 1)
        {
            if (b.i_pixel == PIX_16x16)
                mvd = 25;
            else
                mvd = iabs(b.mvp[0] - mvc[0][0]) + iabs(b.mvp[1] - mvc[0][1]);
        }
        else
        {
            denom = i_mvc - 1;
            mvd = 0;
            if (b.i_pixel != PIX_16x16)
            {
                mvd = iabs(b.mvp[0] - mvc[0][0]) + iabs(b.mvp[1] - mvc[0][1]);
                denom++;
            }
#pragma unroll 1
            for (int i = 0; i < i_mvc - 1; i++)
                mvd += iabs(mvc[i][0] - mvc[i + 1][0]) + iabs(mvc[i][1] - mvc[i + 1][1]);
        }
        const int bc = best_cost(best);
        const int sad_ctx = bc < (1000 >> shift) ? 0 : bc < (2000 >> shift) ? 1 : bc < (4000 >> shift) ? 2 : 3;
        const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
        // multiplier table rows = mvd_ctx, cols = sad_ctx
        const int mul = mvd_ctx == 0 ? (sad_ctx < 2 ? 3 : 4)
                      : mvd_ctx == 1 ? (sad_ctx < 1 ? 3 : 4)
                      : mvd_ctx == 2 ? (sad_ctx < 3 ? 4 : 5)
                      : (sad_ctx < 2 ? 4 : sad_ctx == 2 ? 5 : 6);
        me_range = me_range * mul / 4;
    }

    // still centred on (ox,oy) — the reference keeps the stale centre here on purpose
    best = search_cross(e, b, best, ox, oy, cross_start, me_range, me_range / 2);
    best = try4(b, best, 4, ox, oy, PCAMV_OFF4(-2, -2, 2, 2), PCAMV_OFF4(-2, 2, -2, 2));

    // 16-point hexagon grid, radius 4*i for i = 1 .. me_range/4 (at least one ring): point k = 16*(i-1) + j, one per lane
    ox = pk_x(best_mv(best)); oy = pk_y(best_mv(best));
    {
        const int rings = imax(me_range / 4, 1);
        const int n_pts = 16 * rings;
#pragma unroll 1
        for (int k0 = 0; k0 < n_pts; k0 += PCAMV_WIDE)
        {
            const int k = k0 + team_lane();
            const int i = (k >> 4) + 1, j = k & 15;
            // { -4,-4,-4,-4,-4, 4,4,4,4,4, 2,0,-2,-2,0,2 } and { 2,1,0,-1,-2, -2,-1,0,1,2, 3,4,3,-3,-4,-3 } as signed nibbles
            const int gx = (int)(((long long)(0x20EE0244444CCCCCull << (60 - 4 * j))) >> 60);
            const int gy = (int)(((long long)(0xDCD343210FEEF012ull << (60 - 4 * j))) >> 60);
            const int mx = ox + gx * i, my = oy + gy * i;
            // the reference tests every point against the MV range only when the ring can leave it (me.c:453-470)
            const bool near_edge = 4 * i > imin(imin(e.mv_max_fpel[0] - ox, ox - e.mv_min_fpel[0]),
                                                imin(e.mv_max_fpel[1] - oy, oy - e.mv_min_fpel[1]));
            const bool valid = k < n_pts && (!near_edge || fpel_in_range(e, mx, my));
            const int cost = valid ? lane_sad(b, mx, my) : PCAMV_COST_MAX;
            best = fold_wide(best, cost, pk(mx, my));
        }
    }
    if (pk_y(best_mv(best)) <= e.mv_max_fpel[1])
    {
        *run_hex = 1; *hex_range = me_range;
    }
    return best;
}

// half-pel / quarter-pel refinement (reference encoder/me.c:715-843)
PCAMV_FN void refine_subpel(const MeEnv &env, const MeBlock &b, MeResult &m, int hpel_iters, int qpel_iters,
                             int *p_halfpel_thresh, int b_refine_qpel)
{
    const int chroma = env.chroma_me && b.i_pixel <= PIX_8x8;
    const int mbcmp = !env.mbcmp_satd ? COST_SAD : chroma ? COST_SATD_CHROMA : COST_SATD;   // --subme 1: mbcmp is plain SAD
    int bmx = m.mv[0], bmy = m.mv[1], bcost = m.cost;
    int costs[4];

    if (hpel_iters && env.subme < 3)
    {
        const int mx = clip3(b.mvp[0], env.mv_min_spel[0], env.mv_max_spel[0]);
        const int my = clip3(b.mvp[1], env.mv_min_spel[1], env.mv_max_spel[1]);
        if ((mx - bmx) | (my - bmy))
        {
            const int s = eval1(b, b.k_qsad, pk(mx, my));
            if (s < bcost) { bcost = s; bmx = mx; bmy = my; }
        }
    }

#pragma unroll 1
    for (int i = hpel_iters; i > 0; i--)
    {
        const int omx = bmx, omy = bmy;
        eval4(b, b.k_qsad, 4, pk(omx, omy - 2), pk(omx, omy + 2), pk(omx - 2, omy), pk(omx + 2, omy), costs);
        // the vertical pair only ever moves bmy (reference COPY2_IF_LT on bmy alone)
        if (costs[0] < bcost) { bcost = costs[0]; bmy = omy - 2; }
        if (costs[1] < bcost) { bcost = costs[1]; bmy = omy + 2; }
        if (costs[2] < bcost) { bcost = costs[2]; bmx = omx - 2; bmy = omy; }
        if (costs[3] < bcost) { bcost = costs[3]; bmx = omx + 2; bmy = omy; }
        if (bmx == omx && bmy == omy)
            break;
    }

    if (!b_refine_qpel)
    {
        if (bmy > env.mv_max_spel[1])
            bmy = env.mv_max_spel[1];
        bcost = eval1(b, mbcmp, pk(bmx, bmy));     // bcost was reset to COST_MAX: always taken
    }

    if (p_halfpel_thresh)
    {
        if (((bcost * 7) >> 3) > *p_halfpel_thresh)
        {
            m.cost = bcost; m.mv[0] = bmx; m.mv[1] = bmy;
            return;     // cost_mv deliberately left untouched, as in the reference
        }
        else if (bcost < *p_halfpel_thresh)
            *p_halfpel_thresh = bcost;
    }

    int bdir = -1;
#pragma unroll 1
    for (int i = qpel_iters; i > 0; i--)
    {
        const int odir = bdir;
        const int omx = bmx, omy = bmy;
        // direction d (0 up, 1 down, 2 left, 3 right) is skipped when it would step straight back (d^1 == odir), except in
        // the final refine.  All four are costed (four lane groups are there anyway); the skipped one is ignored in the fold,
        // which is what not evaluating it amounts to.
        eval4(b, mbcmp, 4, pk(omx, omy - 1), pk(omx, omy + 1), pk(omx - 1, omy), pk(omx + 1, omy), costs);
#pragma unroll
        for (int d = 0; d < 4; d++)
        {
            const bool take = b_refine_qpel || (d ^ 1) != odir;
            if (take && costs[d] < bcost)
            {
                bcost = costs[d]; bdir = d;
                bmx = omx + (d == 2 ? -1 : d == 3 ? 1 : 0); bmy = omy + (d == 0 ? -1 : d == 1 ? 1 : 0);
            }
        }
        if (bmx == omx && bmy == omy)
            break;
    }

    if (bmy > env.mv_max_spel[1])
    {
        bmy = env.mv_max_spel[1];
        bcost = eval1(b, mbcmp, pk(bmx, bmy));
    }

    m.cost = bcost;
    m.mv[0] = bmx;
    m.mv[1] = bmy;
    m.cost_mv = b.cost_mvx[bmx] + b.cost_mvy[bmy];
}

PCAMV_DEV void subpel_iters(int subme, int out[4])
{
    // { refine_hpel, refine_qpel, me_hpel, me_qpel } per --subme level (reference encoder/me.c:34-44), one nibble each
    const uint32_t t = subme == 1 ? 0x0011u : subme == 2 ? 0x0110u : subme == 3 ? 0x0120u : subme == 4 ? 0x1120u
                     : subme == 5 ? 0x2120u : subme == 6 || subme == 7 ? 0x2200u : subme >= 8 ? 0xa400u : 0u;
    out[0] = t & 15; out[1] = (t >> 4) & 15; out[2] = (t >> 8) & 15; out[3] = (t >> 12) & 15;
}

// Exhaustive search (reference encoder/me.c:483-634, --me esa / tesa): a raster scan (y outer, x inner, strict <) of the
// window around the best predictor; the scanned width is rounded to a multiple of 4 exactly as the reference rounds it
// (me.c:491), so up to 3 columns right of max_x are visited and the last one may be cut.
// Successive elimination (common/pixel.c:515-559): sum |DC(fenc sub-block) - DC(reference sub-block)| + cost_mvx is a lower
// bound of SAD + cost_mvx; the 8x8 box sums of the reference come from the integral plane k_box_sum8 builds
// (common/mc.c:311-345,477-511).  One candidate column per lane.

// DCs of the (up to four) u x u quadrants of the block, u = 8 (blocks of 8x8 and up) or 4 (sub-8x8 blocks): enc_dc of
// me.c:507-523
PCAMV_DEV int block_unit(const MeBlock &b) { return (b.bw < 8 || b.bh < 8) ? 4 : 8; }
PCAMV_FN void block_dcs(const MeBlock &b, int dc[4])
{
    const smem_ptr fenc = to_smem(b.fenc);
    const int bw = b.bw, bh = b.bh, u = block_unit(b);
    const int words = u * u / 4, wpr = u / 4;         // 4-pixel words per quadrant / per quadrant row
#pragma unroll 1
    for (int q = 0; q < 4; q++)
    {
        const int qx = q & 1, qy = q >> 1;
        int part = 0;
        if (u * qx < bw && u * qy < bh)
        {
#if defined(PCAMV_EMU)
            for (int it = 0; it < words; it++)
                part += sad4(ld4s(fenc + (u * qy + it / wpr) * 16 + u * qx + 4 * (it % wpr)), 0u);
#else
            const int it = team_lane();
            part = it < words ? sad4(ld4s(fenc + (u * qy + it / wpr) * 16 + u * qx + 4 * (it % wpr)), 0u) : 0;
            part = team_sum(part);
#endif
        }
        dc[q] = part;
    }
}

PCAMV_DEV int ld_u16(const uint16_t *p)
{
#if defined(PCAMV_EMU)
    return *p;
#else
    return __ldg(p);
#endif
}

// x264_pixel_ads4 / ads2 / ads1 at full-pel (mx, my)
PCAMV_DEV int ads_at(const MeBlock &b, const int dc[4], int mx, int my)
{
    const int stride = b.stride, u = block_unit(b);
    const uint16_t *s = (u == 4 ? b.integral4 : b.integral) + my * stride + mx;
    const bool two_x = b.bw == 2 * u, two_y = b.bh == 2 * u;
    int a = iabs(dc[0] - ld_u16(s));
    if (two_x) a += iabs(dc[1] - ld_u16(s + u));
    if (two_y) a += iabs(dc[2] - ld_u16(s + u * stride));
    if (two_x && two_y) a += iabs(dc[3] - ld_u16(s + u * stride + u));
    return a + b.cost_mvx[mx << 2];
}
PCAMV_DEV bool have_box_sums(const MeBlock &b) { return block_unit(b) == 4 ? b.integral4 != nullptr : b.integral != nullptr; }

PCAMV_FN best_t search_esa(const MeEnv &e, const MeBlock &b, best_t best)
{
    const int range = e.me_range;
    const int bmx = pk_x(best_mv(best)), bmy = pk_y(best_mv(best));
    const int min_x = imax(bmx - range, e.mv_min_fpel[0]), min_y = imax(bmy - range, e.mv_min_fpel[1]);
    const int max_x = imin(bmx + range, e.mv_max_fpel[0]), max_y = imin(bmy + range, e.mv_max_fpel[1]);
    const int width = (max_x - min_x + 3) & ~3;
    if (!have_box_sums(b))
    {
        // without box sums that bound this block (contexts opened without the matching integral plane): every position gets its SAD
#pragma unroll 1
        for (int my = min_y; my <= max_y; my++)
#pragma unroll 1
            for (int x0 = 0; x0 < width; x0 += PCAMV_WIDE)
            {
                const int x = x0 + team_lane();
                const int cost = x < width ? lane_sad(b, min_x + x, my) : PCAMV_COST_MAX;
                best = fold_wide(best, cost, pk(min_x + x, my));
            }
        return best;
    }
    // A position whose lower bound is not below the running best cannot win under strict <: skipping it (and whole
    // rows whose y cost alone reaches the best, me.c:617-619) leaves the first minimum of the raster scan unchanged.
    int dc[4];
    block_dcs(b, dc);
#pragma unroll 1
    for (int my = min_y; my <= max_y; my++)
    {
        const int ycost = b.cost_mvy[my << 2];
        if (best_cost(best) <= ycost)
            continue;
#pragma unroll 1
        for (int x0 = 0; x0 < width; x0 += PCAMV_WIDE)
        {
            const int x = x0 + team_lane();
            const bool pass = x < width && ads_at(b, dc, min_x + x, my) < best_cost(best) - ycost;
            unsigned m = team_ballot(pass);
            if (!m)
                continue;
            if (popc32(m) > 8)
            {
                const int cost = pass ? lane_sad(b, min_x + x, my) : PCAMV_COST_MAX;
                best = fold_wide(best, cost, pk(min_x + x, my));
                continue;
            }
            // few survivors: the whole team costs them four at a time, in x order
#pragma unroll 1
            while (m)
            {
                int l[4], n = 0;
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    l[k] = 0;
                    if (m) { l[k] = ctz32(m); m &= m - 1; n++; }
                }
                best = try4(b, best, n, min_x + x0, my, PCAMV_OFF4(l[0], l[1], l[2], l[3]), 0);
            }
        }
    }
    return best;
}

// --me tesa (me.c:524-613): ADS threshold, then SAD threshold, keep the best few SADs, then SATD on those.
// Unlike ESA the prefilters decide the outcome here, so thresholds, list order and the partial selection sort are the
// reference's own; only the evaluation is spread over the lanes (the running SAD threshold a sequential scan would hold
// before each candidate is the prefix minimum over the lower lanes).
PCAMV_DEV unsigned long long mvsad_make(int sad, int mx, int my)
{
    return ((unsigned long long)(uint32_t)sad << 32) | ((uint32_t)(mx & 0xffff) << 16) | (uint32_t)(my & 0xffff);
}
PCAMV_DEV int mvsad_sad(unsigned long long v) { return (int)(v >> 32); }
PCAMV_DEV int mvsad_mv(unsigned long long v) { return pk((int)(int16_t)(v >> 16), (int)(int16_t)v); }

PCAMV_FN best_t search_tesa(const MeEnv &e, const MeBlock &b, best_t best)
{
    const int range = e.me_range;
    const int bmx = pk_x(best_mv(best)), bmy = pk_y(best_mv(best));
    const int min_x = imax(bmx - range, e.mv_min_fpel[0]), min_y = imax(bmy - range, e.mv_min_fpel[1]);
    const int max_x = imin(bmx + range, e.mv_max_fpel[0]), max_y = imin(bmy + range, e.mv_max_fpel[1]);
    const int width = (max_x - min_x + 3) & ~3;
    const int sad_thresh = range <= 16 ? 10 : range <= 24 ? 11 : 12;
    unsigned long long *mvsads = e.mvsads;
    int dc[4];
    block_dcs(b, dc);
    int n = 0;
    int bsad = eval1(b, COST_SAD_FPEL, pk(bmx << 2, bmy << 2));      // plain SAD + BITS_MVD(bmx, bmy)
#pragma unroll 1
    for (int my = min_y; my <= max_y; my++)
    {
        const int ycost = b.cost_mvy[my << 2];
        if (bsad <= ycost)
            continue;
        bsad -= ycost;
        const int thresh = bsad * 17 / 16;          // fixed for the row, as in the reference's one ads call per row
#pragma unroll 1
        for (int x0 = 0; x0 < width; x0 += PCAMV_WIDE)
        {
            const int x = x0 + team_lane();
            const bool pass = x < width && ads_at(b, dc, min_x + x, my) < thresh;
            if (!team_ballot(pass))
                continue;
            // Reference quirk, reproduced: the SAD stage looks the x cost up with the column index RELATIVE to min_x
            // (cost_fpel_mvx[xs[i]], me.c:549,565) where the ads call was handed cost_fpel_mvx + min_x, so the MV cost in
            // these SADs is that of column x, not of min_x + x.
            const int sad = pass ? lane_sad(b, min_x + x, my) - ycost - b.cost_mvx[(min_x + x) << 2] + b.cost_mvx[x << 2] : 0x7fffffff;
            const int run = imin(bsad, team_prefix_min_excl(sad));
            const bool take = pass && sad < ((run * sad_thresh) >> 3);
            const unsigned tm = team_ballot(take);
            if (take)
                mvsads[n + popc_below(tm)] = mvsad_make(sad + ycost, min_x + x, my);
            n += popc32(tm);
            bsad = imin(bsad, team_min(sad));
        }
        bsad += ycost;
    }
    team_sync();
    const int limit = range / 2;
    if (n > limit * 2)
    {
        // halve the range if the domain is too large: keep, in order, the entries within the relaxed threshold
        const int cut = (bsad * (sad_thresh + 8)) >> 4;
        int w = 0;
#pragma unroll 1
        for (int j0 = 0; j0 < n; j0 += PCAMV_WIDE)
        {
            const int j = j0 + team_lane();
            const unsigned long long v = j < n ? mvsads[j] : 0ull;
            const bool keep = j < n && mvsad_sad(v) <= cut;
            const unsigned km = team_ballot(keep);
            team_sync();                        // the chunk is in registers before any of it is overwritten
            if (keep)
                mvsads[w + popc_below(km)] = v;
            w += popc32(km);
            team_sync();
        }
        n = w;
    }
    if (n > limit)
    {
        // partial selection sort of the first `limit` entries: first minimum, swapped into place
#pragma unroll 1
        for (int i = 0; i < limit; i++)
        {
            int bs = 0x7fffffff, bj = 0x7fffffff;
#pragma unroll 1
            for (int j = i + team_lane(); j < n; j += PCAMV_WIDE)
            {
                const int sj = mvsad_sad(mvsads[j]);
                if (sj < bs) { bs = sj; bj = j; }
            }
            const int mn = team_min(bs);
            bj = team_min(bs == mn ? bj : 0x7fffffff);
            if (bj > i && team_lane() == 0)
            {
                const unsigned long long t = mvsads[i];
                mvsads[i] = mvsads[bj];
                mvsads[bj] = t;
            }
            team_sync();
        }
        n = limit;
    }
    // fpelcmp (SATD here) of the survivors, in list order
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 4)
    {
        const int nn = imin(4, n - i0);
        int mv[4], costs[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            mv[k] = mvsad_mv(mvsads[i0 + (k < nn ? k : 0)]);
        eval4(b, b.k_fpel, nn, pk(pk_x(mv[0]) << 2, pk_y(mv[0]) << 2), pk(pk_x(mv[1]) << 2, pk_y(mv[1]) << 2),
              pk(pk_x(mv[2]) << 2, pk_y(mv[2]) << 2), pk(pk_x(mv[3]) << 2, pk_y(mv[3]) << 2), costs);
        int bcost = best_cost(best), bmv = best_mv(best);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nn && costs[k] < bcost) { bcost = costs[k]; bmv = mv[k]; }
        best = best_make(bcost, bmv);
    }
    return best;
}

// XS: the exhaustive searches (--me esa / tesa) are compiled in.  The kernels of the pattern searches are instantiated
// without them: merely having that code in the call graph cost the UMH workload 7 % (register allocation and code layout
// of the whole per-macroblock call tree; measured on B200).
template <int XS>
PCAMV_FN void me_search_ref(const MeEnv &env, const MeBlock &b, const int (*mvc)[2], int i_mvc,
                             int *p_halfpel_thresh, MeResult &m)
{
    const int x_min = env.mv_min_fpel[0], y_min = env.mv_min_fpel[1], x_max = env.mv_max_fpel[0], y_max = env.mv_max_fpel[1];
    int bpred_mv = 0, bpred_cost = PCAMV_COST_MAX;
    int bmx = clip3(b.mvp[0], x_min * 4, x_max * 4);
    int bmy = clip3(b.mvp[1], y_min * 4, y_max * 4);
    const int pmx = (bmx + 2) >> 2, pmy = (bmy + 2) >> 2;
    best_t best;

    if (env.subme >= 3)
    {
        // predictors at quarter-pel precision: mvp first, then each distinct non-zero candidate; costed four at a
        // time and folded in order (strict <)
        int c[4], n = 0;
        c[0] = c[1] = c[2] = c[3] = pk(bmx, bmy);
        n = 1;
#pragma unroll 1
        for (int i = 0; i <= i_mvc; i++)
        {
            bool have = false;
            int v = 0;
            if (i < i_mvc)
            {
                const int cx = mvc[i][0], cy = mvc[i][1];
                const bool nonzero = (cx | cy) != 0;
                const bool differs = ((cx & 0xffff) != (bmx & 0xffff)) || ((cy & 0xffff) != (bmy & 0xffff));
                have = nonzero && differs;
                v = pk(clip3(cx, x_min * 4, x_max * 4), clip3(cy, y_min * 4, y_max * 4));
            }
            if (have)
            {
                if (n == 0) c[0] = v; else if (n == 1) c[1] = v; else if (n == 2) c[2] = v; else c[3] = v;
                n++;
            }
            if (n == 4 || (i == i_mvc && n > 0))
            {
                int costs[4];
                eval4(b, b.k_qsad, n, c[0], c[1], c[2], c[3], costs);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < n && costs[k] < bpred_cost) { bpred_cost = costs[k]; bpred_mv = c[k]; }
                n = 0;
            }
        }
        bmx = (pk_x(bpred_mv) + 2) >> 2;
        bmy = (pk_y(bpred_mv) + 2) >> 2;
        // COST_MV(bmx,bmy) against bcost = COST_MAX: always taken
        best = try1(b, best_make(PCAMV_COST_MAX, 0), bmx, bmy);
    }
    else
    {
        // COST_MV then minus BITS_MVD(pmx,pmy): the plain SAD at the rounded predictor
        best = try1(b, best_make(PCAMV_COST_MAX, 0), pmx, pmy);
        best = best_make(best_cost(best) - (b.cost_mvx[pmx << 2] + b.cost_mvy[pmy << 2]), pk(pmx, pmy));
#pragma unroll 1
        for (int i = 0; i < i_mvc; i++)
        {
            int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
            if ((mx | my) && ((mx - pk_x(best_mv(best))) | (my - pk_y(best_mv(best)))))
            {
                mx = clip3(mx, x_min, x_max);
                my = clip3(my, y_min, y_max);
                best = try1(b, best, mx, my);
            }
        }
    }
    best = try1(b, best, 0, 0);

    int run_hex = env.me_method == ME_HEX, hex_range = env.me_range;
    if (env.me_method == ME_DIA) best = search_dia(env, b, best, env.me_range);
    else if (env.me_method == ME_UMH) best = search_umh(env, b, best, pmx, pmy, mvc, i_mvc, &run_hex, &hex_range);
    else if (XS && env.me_method == ME_TESA && env.mbcmp_satd) best = search_tesa(env, b, best);
    else if (XS && env.me_method != ME_HEX) best = search_esa(env, b, best);
    if (run_hex)
        best = search_hex(env, b, best, hex_range);

    bmx = pk_x(best_mv(best)); bmy = pk_y(best_mv(best));
    if (bpred_cost < best_cost(best))
    {
        m.mv[0] = pk_x(bpred_mv); m.mv[1] = pk_y(bpred_mv); m.cost = bpred_cost;
    }
    else
    {
        m.mv[0] = bmx << 2; m.mv[1] = bmy << 2; m.cost = best_cost(best);
    }
    m.cost_mv = b.cost_mvx[m.mv[0]] + b.cost_mvy[m.mv[1]];
    if (bmx == pmx && bmy == pmy && env.subme < 3)
        m.cost += m.cost_mv;

    if (env.subme >= 2)
    {
        int it[4];
        subpel_iters(env.subme, it);
        refine_subpel(env, b, m, it[2], it[3], p_halfpel_thresh, 0);
    }
    else if (m.mv[1] > env.mv_max_spel[1])
        m.mv[1] = env.mv_max_spel[1];
}

// x264_me_refine_qpel (reference encoder/me.c:669-678); i_ref_cost is removed first for P blocks <= 8x8
PCAMV_FN void me_refine_qpel(const MeEnv &env, const MeBlock &b, MeResult &m, int i_ref_cost)
{
    int it[4];
    subpel_iters(env.subme, it);
    if (b.i_pixel <= PIX_8x8)
        m.cost -= i_ref_cost;
    refine_subpel(env, b, m, it[0], it[1], (int *)0, 1);
}

} // namespace pcamv
