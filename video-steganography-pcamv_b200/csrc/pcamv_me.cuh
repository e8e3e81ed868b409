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
    const uint16_t *integral; // ESA only
    int stride, stride_c;
    int i_pixel, bw, bh;
    int mvp[2];
    const int16_t *cost_mvx, *cost_mvy;   // cost_mv - mvp[0], cost_mv - mvp[1]
};

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
    s2 = (idx & 5) ? b.ref[h1] + off + (fx == 3 ? 1 : 0) : (const uint8_t *)0;
}

// 4 predicted luma pixels at block-relative (x, y) for sources prepared by qpel_sources
PCAMV_DEV uint32_t pred4(const uint8_t *s1, const uint8_t *s2, int stride, int x, int y)
{
    uint32_t w = ld4(s1 + y * stride + x);
    if (s2)
        w = avg4(w, ld4(s2 + y * stride + x));
    return w;
}

// 4 predicted chroma pixels (reference common/mc.c:246-277), src already at the block origin
PCAMV_DEV uint32_t chroma4(const uint8_t *src, int stride, int qmx, int qmy, int x, int y)
{
    const int dx = qmx & 7, dy = qmy & 7;
    const int cA = (8 - dx) * (8 - dy), cB = dx * (8 - dy), cC = (8 - dx) * dy, cD = dx * dy;
    const uint8_t *s = src + ((qmy >> 3) + y) * stride + (qmx >> 3) + x;
    const uint32_t t0 = ld4(s), t1 = ld4(s + 1), b0 = ld4(s + stride), b1 = ld4(s + stride + 1);
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        r |= (uint32_t)((cA * px(t0, i) + cB * px(t1, i) + cC * px(b0, i) + cD * px(b1, i) + 32) >> 6) << (8 * i);
    return r;
}

// ---- distortion of up to N candidates, concurrently -----------------------------------------------
// SAD at quarter-pel candidates (integer/half positions take the single-plane path).
// out[i] = SAD only; the caller adds the MV cost.
PCAMV_FN void sad_cands(const MeBlock &b, int n, const int *qmx, const int *qmy, int *out)
{
    const int grp = team_grp(), sub = team_sub();
    const int w4 = b.bw >> 2, words = w4 * b.bh;
    for (int c0 = 0; c0 < n; c0 += PCAMV_NGRP)
    {
        const int c = c0 + grp;
        int acc = 0;
        if (c < n)
        {
            const uint8_t *s1, *s2;
            qpel_sources(b, qmx[c], qmy[c], s1, s2);
            for (int w = sub; w < words; w += PCAMV_LPG)
            {
                const int y = w / w4, x = (w - y * w4) << 2;
                acc += sad4(ld4a(b.fenc + y * 16 + x), pred4(s1, s2, b.stride, x, y));
            }
        }
        acc = grp_sum(acc);
#pragma unroll
        for (int g = 0; g < PCAMV_NGRP; g++)
            if (c0 + g < n)
                out[c0 + g] = grp_bcast(acc, g);
    }
}

// SAD at full-pel candidates (mx,my in pixels): integer plane only.
PCAMV_FN void sad_fpel_cands(const MeBlock &b, int n, const int *mx, const int *my, int *out)
{
    const int grp = team_grp(), sub = team_sub();
    const int w4 = b.bw >> 2, words = w4 * b.bh;
    for (int c0 = 0; c0 < n; c0 += PCAMV_NGRP)
    {
        const int c = c0 + grp;
        int acc = 0;
        if (c < n)
        {
            const uint8_t *s = b.ref[0] + my[c] * b.stride + mx[c];
            for (int w = sub; w < words; w += PCAMV_LPG)
            {
                const int y = w / w4, x = (w - y * w4) << 2;
                acc += sad4(ld4a(b.fenc + y * 16 + x), ld4(s + y * b.stride + x));
            }
        }
        acc = grp_sum(acc);
#pragma unroll
        for (int g = 0; g < PCAMV_NGRP; g++)
            if (c0 + g < n)
                out[c0 + g] = grp_bcast(acc, g);
    }
}

// SATD of a luma block against prepared sources; lanes of the group split the 8x4 / 4x4 units.
PCAMV_DEV int satd_luma_part(const uint8_t *fenc, int fstride, int bw, int bh,
                             const uint8_t *s1, const uint8_t *s2, int stride, int sub)
{
    uint32_t acc = 0;
    if (bw >= 8)
    {
        const int ux = bw >> 3, units = ux * (bh >> 2);
        for (int u = sub; u < units; u += PCAMV_LPG)
        {
            const int x = (u % ux) << 3, y = (u / ux) << 2;
            uint32_t f[4], g[4], a[4], c[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                f[r] = ld4a(fenc + (y + r) * fstride + x);
                g[r] = ld4a(fenc + (y + r) * fstride + x + 4);
                a[r] = pred4(s1, s2, stride, x, y + r);
                c[r] = pred4(s1, s2, stride, x + 4, y + r);
            }
            acc += hadamard_8x4_sum(f, g, a, c);
        }
    }
    else
    {
        const int units = bh >> 2;
        for (int u = sub; u < units; u += PCAMV_LPG)
        {
            const int y = u << 2;
            uint32_t f[4], a[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                f[r] = ld4a(fenc + (y + r) * fstride);
                a[r] = pred4(s1, s2, stride, 0, y + r);
            }
            acc += hadamard_4x4_sum(f, a);
        }
    }
    return (int)acc;
}

// SATD of one chroma plane block (cw x ch, cw in {4,8}) against bilinear MC at (qmx,qmy)
PCAMV_DEV int satd_chroma_part(const uint8_t *fenc, int fstride, int cw, int ch,
                               const uint8_t *src, int stride, int qmx, int qmy, int sub, int sub_off)
{
    uint32_t acc = 0;
    if (cw >= 8)
    {
        const int units = ch >> 2;
        for (int u = 0; u < units; u++)
        {
            if (((u + sub_off) % PCAMV_LPG) != sub) continue;
            const int y = u << 2;
            uint32_t f[4], g[4], a[4], c[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                f[r] = ld4a(fenc + (y + r) * fstride);
                g[r] = ld4a(fenc + (y + r) * fstride + 4);
                a[r] = chroma4(src, stride, qmx, qmy, 0, y + r);
                c[r] = chroma4(src, stride, qmx, qmy, 4, y + r);
            }
            acc += hadamard_8x4_sum(f, g, a, c);
        }
    }
    else
    {
        const int units = ch >> 2;
        for (int u = 0; u < units; u++)
        {
            if (((u + sub_off) % PCAMV_LPG) != sub) continue;
            const int y = u << 2;
            uint32_t f[4], a[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                f[r] = ld4a(fenc + (y + r) * fstride);
                a[r] = chroma4(src, stride, qmx, qmy, 0, y + r);
            }
            acc += hadamard_4x4_sum(f, a);
        }
    }
    return (int)acc;
}

// SATD (luma, plus both chroma planes when `chroma`) at quarter-pel candidates.
// out[i] = luma SATD + U SATD + V SATD (each already halved like the reference's functions).
// The reference adds chroma only while the running cost is still below the best (encoder/me.c:689-713);
// chroma terms are non-negative, so adding them unconditionally gives the same decisions.
PCAMV_FN void satd_cands(const MeBlock &b, int n, const int *qmx, const int *qmy, int chroma, int *out,
                          const uint8_t *fenc_y, const uint8_t *fenc_u, const uint8_t *fenc_v, int use_satd = 1)
{
    const int grp = team_grp(), sub = team_sub();
    if (!use_satd)
    {
        // --subme 1: mbcmp is plain SAD (chroma ME is never on below subme 5)
        sad_cands(b, n, qmx, qmy, out);
        return;
    }
    for (int c0 = 0; c0 < n; c0 += PCAMV_NGRP)
    {
        const int c = c0 + grp;
        int accY = 0, accU = 0, accV = 0;
        if (c < n)
        {
            const uint8_t *s1, *s2;
            qpel_sources(b, qmx[c], qmy[c], s1, s2);
            accY = satd_luma_part(fenc_y, 16, b.bw, b.bh, s1, s2, b.stride, sub);
            if (chroma)
            {
                // spread the chroma units over the lanes that have the fewest luma units
                accU = satd_chroma_part(fenc_u, 8, b.bw >> 1, b.bh >> 1, b.ref_u, b.stride_c, qmx[c], qmy[c], sub, PCAMV_LPG - 1);
                accV = satd_chroma_part(fenc_v, 8, b.bw >> 1, b.bh >> 1, b.ref_v, b.stride_c, qmx[c], qmy[c], sub, PCAMV_LPG - 3);
            }
        }
        // each plane's SATD is halved separately in the reference (three function calls)
        accY = grp_sum(accY) >> 1;
        if (chroma)
        {
            accU = grp_sum(accU) >> 1;
            accV = grp_sum(accV) >> 1;
        }
        const int tot = accY + accU + accV;
#pragma unroll
        for (int g = 0; g < PCAMV_NGRP; g++)
            if (c0 + g < n)
                out[c0 + g] = grp_bcast(tot, g);
    }
}

// =====================================================================================================
// Integer-pel search + sub-pel refinement of one block
// =====================================================================================================
struct MeSearch
{
    const MeEnv &env;
    const MeBlock &b;
    int bmx, bmy, bcost;           // running best (full-pel units during the integer search)
    int mv_x_min, mv_y_min, mv_x_max, mv_y_max;

    PCAMV_MEM MeSearch(const MeEnv &e, const MeBlock &blk) : env(e), b(blk)
    {
        mv_x_min = e.mv_min_fpel[0]; mv_y_min = e.mv_min_fpel[1];
        mv_x_max = e.mv_max_fpel[0]; mv_y_max = e.mv_max_fpel[1];
    }
    PCAMV_MEM int bits_fpel(int mx, int my) const { return b.cost_mvx[mx << 2] + b.cost_mvy[my << 2]; }
    PCAMV_MEM bool in_range(int mx, int my) const
    {
        return mx >= mv_x_min && mx <= mv_x_max && my >= mv_y_min && my <= mv_y_max;
    }
    // cost n (<=4) full-pel candidates and fold them into the best in order (strict <)
    PCAMV_MEM void try_fpel(int n, const int *mx, const int *my)
    {
        int sad[4];
        sad_fpel_cands(b, n, mx, my, sad);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < n)
            {
                const int c = sad[i] + bits_fpel(mx[i], my[i]);
                if (c < bcost) { bcost = c; bmx = mx[i]; bmy = my[i]; }
            }
    }
    PCAMV_MEM void try_fpel1(int mx, int my) { try_fpel(1, &mx, &my); }
    // four offsets around (ox,oy)  — the reference's COST_MV_X4
    PCAMV_MEM void try_x4(int ox, int oy, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3)
    {
        const int mx[4] = { ox + x0, ox + x1, ox + x2, ox + x3 };
        const int my[4] = { oy + y0, oy + y1, oy + y2, oy + y3 };
        try_fpel(4, mx, my);
    }
    PCAMV_MEM void dia1(int ox, int oy) { try_x4(ox, oy, 0, -1, 0, 1, -1, 0, 1, 0); }

    // symmetric cross around (ox,oy)  (reference encoder/me.c:131-155)
    PCAMV_MEMFN void cross(int ox, int oy, int start, int x_max, int y_max)
    {
        int i = start;
        if (x_max <= imin(mv_x_max - ox, ox - mv_x_min))
            for (; i < x_max - 2; i += 4)
                try_x4(ox, oy, i, 0, -i, 0, i + 2, 0, -i - 2, 0);
        for (; i < x_max; i += 2)
        {
            if (ox + i <= mv_x_max) try_fpel1(ox + i, oy);
            if (ox - i >= mv_x_min) try_fpel1(ox - i, oy);
        }
        i = start;
        if (y_max <= imin(mv_y_max - oy, oy - mv_y_min))
            for (; i < y_max - 2; i += 4)
                try_x4(ox, oy, 0, i, 0, -i, 0, i + 2, 0, -i - 2);
        for (; i < y_max; i += 2)
        {
            if (oy + i <= mv_y_max) try_fpel1(ox, oy + i);
            if (oy - i >= mv_y_min) try_fpel1(ox, oy - i);
        }
    }

    // hexagon (radius 2) walk followed by the 8-point square  (reference encoder/me.c:263-341)
    PCAMV_MEMFN void hex_then_square(int me_range)
    {
        // hexagon corner k (k = 0..5), walking order matches the reference's direction indices
        const int hx[6] = { -2, -1, 1, 2, 1, -1 };
        const int hy[6] = { 0, 2, 2, 0, -2, -2 };
        int mx[4], my[4], sad[4];
        int dir = -1;
        // first step: all six corners, in the order 0..5
        int best6 = bcost;
        for (int half = 0; half < 2; half++)
        {
            for (int k = 0; k < 3; k++) { mx[k] = bmx + hx[3 * half + k]; my[k] = bmy + hy[3 * half + k]; }
            sad_fpel_cands(b, 3, mx, my, sad);
            for (int k = 0; k < 3; k++)
            {
                const int c = sad[k] + bits_fpel(mx[k], my[k]);
                if (c < best6) { best6 = c; dir = 3 * half + k; }
            }
        }
        bcost = best6;
        if (dir >= 0)
        {
            bmx += hx[dir]; bmy += hy[dir];
            // half hexagons: only the three corners not covered by the previous position
            for (int i = 1; i < me_range / 2 && in_range(bmx, bmy); i++)
            {
                int nd = -1;
                for (int k = 0; k < 3; k++)
                {
                    const int d = (dir + 5 + k) % 6;     // dir-1, dir, dir+1
                    mx[k] = bmx + hx[d]; my[k] = bmy + hy[d];
                }
                sad_fpel_cands(b, 3, mx, my, sad);
                for (int k = 0; k < 3; k++)
                {
                    const int c = sad[k] + bits_fpel(mx[k], my[k]);
                    if (c < bcost) { bcost = c; nd = (dir + 5 + k) % 6; }
                }
                if (nd < 0)
                    break;
                dir = nd;
                bmx += hx[dir]; bmy += hy[dir];
            }
        }
        // square refine around the final centre
        const int ox = bmx, oy = bmy;
        try_x4(ox, oy, 0, -1, 0, 1, -1, 0, 1, 0);
        try_x4(ox, oy, -1, -1, -1, 1, 1, -1, 1, 1);
    }

    PCAMV_MEM void search_dia(int me_range)
    {
        int i = 0;
        do
        {
            const int ox = bmx, oy = bmy;
            dia1(ox, oy);
            if (bmx == ox && bmy == oy) break;
            if (!in_range(bmx, bmy)) break;
        } while (++i < me_range);
    }

    // uneven-cross multi-hexagon-grid search (reference encoder/me.c:342-482)
    PCAMV_MEMFN void search_umh(int pmx, int pmy, const int (*mvc)[2], int i_mvc)
    {
        int me_range = env.me_range;
        const int shift = b.i_pixel == PIX_16x16 ? 0 : b.i_pixel <= PIX_8x16 ? 1 : b.i_pixel == PIX_8x8 ? 2
                        : b.i_pixel <= PIX_4x8 ? 3 : 4;
        int cross_start = 1;
        const int ucost1 = bcost;
        dia1(pmx, pmy);
        if (pmx | pmy)
            dia1(0, 0);
        if (b.i_pixel == PIX_4x4) { hex_then_square(me_range); return; }

        const int ucost2 = bcost;
        if ((bmx | bmy) && ((bmx - pmx) | (bmy - pmy)))
            dia1(bmx, bmy);
        if (bcost == ucost2)
            cross_start = 3;
        int ox = bmx, oy = bmy;

        if (bcost == ucost2 && bcost < (2000 >> shift))
        {
            try_x4(ox, oy, 0, -2, -1, -1, 1, -1, -2, 0);
            try_x4(ox, oy, 2, 0, -1, 1, 1, 1, 0, 2);
            if (bcost == ucost1 && bcost < (500 >> shift))
                return;
            if (bcost == ucost2)
            {
                const int range = (me_range >> 1) | 1;
                cross(ox, oy, 3, range, range);
                try_x4(ox, oy, -1, -2, 1, -2, -2, -1, 2, -1);
                try_x4(ox, oy, -2, 1, 2, 1, -1, 2, 1, 2);
                if (bcost == ucost2)
                    return;
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
                for (int i = 0; i < i_mvc - 1; i++)
                    mvd += iabs(mvc[i][0] - mvc[i + 1][0]) + iabs(mvc[i][1] - mvc[i + 1][1]);
            }
            const int sad_ctx = bcost < (1000 >> shift) ? 0 : bcost < (2000 >> shift) ? 1 : bcost < (4000 >> shift) ? 2 : 3;
            const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            // multiplier table rows = mvd_ctx, cols = sad_ctx
            const int mul = mvd_ctx == 0 ? (sad_ctx < 2 ? 3 : 4)
                          : mvd_ctx == 1 ? (sad_ctx < 1 ? 3 : 4)
                          : mvd_ctx == 2 ? (sad_ctx < 3 ? 4 : 5)
                          : (sad_ctx < 2 ? 4 : sad_ctx == 2 ? 5 : 6);
            me_range = me_range * mul / 4;
        }

        // still centred on (ox,oy) — the reference keeps the stale centre here on purpose
        cross(ox, oy, cross_start, me_range, me_range / 2);
        try_x4(ox, oy, -2, -2, -2, 2, 2, -2, 2, 2);

        // 16-point hexagon grid, radius 4*i
        ox = bmx; oy = bmy;
        const int gx[16] = { -4, -4, -4, -4, -4, 4, 4, 4, 4, 4, 2, 0, -2, -2, 0, 2 };
        const int gy[16] = { 2, 1, 0, -1, -2, -2, -1, 0, 1, 2, 3, 4, 3, -3, -4, -3 };
        int i = 1;
        do
        {
            if (4 * i > imin(imin(mv_x_max - ox, ox - mv_x_min), imin(mv_y_max - oy, oy - mv_y_min)))
            {
                for (int j = 0; j < 16; j++)
                {
                    const int mx = ox + gx[j] * i, my = oy + gy[j] * i;
                    if (in_range(mx, my))
                        try_fpel1(mx, my);
                }
            }
            else
            {
                for (int j = 0; j < 16; j += 4)
                    try_x4(ox, oy, gx[j] * i, gy[j] * i, gx[j + 1] * i, gy[j + 1] * i,
                           gx[j + 2] * i, gy[j + 2] * i, gx[j + 3] * i, gy[j + 3] * i);
            }
        } while (++i <= me_range / 4);
        if (bmy <= mv_y_max)
            hex_then_square(me_range);
    }
};

// half-pel / quarter-pel refinement (reference encoder/me.c:715-843)
PCAMV_FN void refine_subpel(const MeEnv &env, const MeBlock &b, MeResult &m, int hpel_iters, int qpel_iters,
                             int *p_halfpel_thresh, int b_refine_qpel)
{
    const int chroma = env.chroma_me && b.i_pixel <= PIX_8x8;
    int bmx = m.mv[0], bmy = m.mv[1], bcost = m.cost;
    int odir = -1, bdir;

    if (hpel_iters && env.subme < 3)
    {
        const int mx = clip3(b.mvp[0], env.mv_min_spel[0], env.mv_max_spel[0]);
        const int my = clip3(b.mvp[1], env.mv_min_spel[1], env.mv_max_spel[1]);
        if ((mx - bmx) | (my - bmy))
        {
            int s;
            sad_cands(b, 1, &mx, &my, &s);
            s += b.cost_mvx[mx] + b.cost_mvy[my];
            if (s < bcost) { bcost = s; bmx = mx; bmy = my; }
        }
    }

    for (int i = hpel_iters; i > 0; i--)
    {
        const int omx = bmx, omy = bmy;
        const int qx[4] = { omx, omx, omx - 2, omx + 2 };
        const int qy[4] = { omy - 2, omy + 2, omy, omy };
        int s[4];
        sad_cands(b, 4, qx, qy, s);
        // the vertical pair only ever moves bmy (reference COPY2_IF_LT on bmy alone)
        int c = s[0] + b.cost_mvx[omx] + b.cost_mvy[omy - 2];
        if (c < bcost) { bcost = c; bmy = omy - 2; }
        c = s[1] + b.cost_mvx[omx] + b.cost_mvy[omy + 2];
        if (c < bcost) { bcost = c; bmy = omy + 2; }
        c = s[2] + b.cost_mvx[omx - 2] + b.cost_mvy[omy];
        if (c < bcost) { bcost = c; bmx = omx - 2; bmy = omy; }
        c = s[3] + b.cost_mvx[omx + 2] + b.cost_mvy[omy];
        if (c < bcost) { bcost = c; bmx = omx + 2; bmy = omy; }
        if (bmx == omx && bmy == omy)
            break;
    }

    if (!b_refine_qpel)
    {
        if (bmy > env.mv_max_spel[1])
            bmy = env.mv_max_spel[1];
        int s;
        satd_cands(b, 1, &bmx, &bmy, chroma, &s, b.fenc, b.fenc_u, b.fenc_v, env.mbcmp_satd);
        bcost = s + b.cost_mvx[bmx] + b.cost_mvy[bmy];     // bcost was reset to COST_MAX: always taken
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

    bdir = -1;
    for (int i = qpel_iters; i > 0; i--)
    {
        odir = bdir;
        const int omx = bmx, omy = bmy;
        // direction d is skipped when it would step straight back (d^1 == odir), except in final refine
        int qx[4], qy[4], dirs[4], n = 0;
        const int dx[4] = { 0, 0, -1, 1 }, dy[4] = { -1, 1, 0, 0 };
        for (int d = 0; d < 4; d++)
            if (b_refine_qpel || (d ^ 1) != odir)
            {
                qx[n] = omx + dx[d]; qy[n] = omy + dy[d]; dirs[n] = d; n++;
            }
        int s[4];
        satd_cands(b, n, qx, qy, chroma, s, b.fenc, b.fenc_u, b.fenc_v, env.mbcmp_satd);
        for (int k = 0; k < n; k++)
        {
            const int c = s[k] + b.cost_mvx[qx[k]] + b.cost_mvy[qy[k]];
            if (c < bcost) { bcost = c; bmx = qx[k]; bmy = qy[k]; bdir = dirs[k]; }
        }
        if (bmx == omx && bmy == omy)
            break;
    }

    if (bmy > env.mv_max_spel[1])
    {
        bmy = env.mv_max_spel[1];
        int s;
        satd_cands(b, 1, &bmx, &bmy, chroma, &s, b.fenc, b.fenc_u, b.fenc_v, env.mbcmp_satd);
        bcost = s + b.cost_mvx[bmx] + b.cost_mvy[bmy];
    }

    m.cost = bcost;
    m.mv[0] = bmx;
    m.mv[1] = bmy;
    m.cost_mv = b.cost_mvx[bmx] + b.cost_mvy[bmy];
}

PCAMV_DEV void subpel_iters(int subme, int out[4])
{
    // { refine_hpel, refine_qpel, me_hpel, me_qpel } per --subme level (reference encoder/me.c:34-44)
    const int t[10][4] = { {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2},
                           {0,0,2,2}, {0,0,2,2}, {0,0,4,10}, {0,0,4,10} };
    for (int i = 0; i < 4; i++) out[i] = t[subme][i];
}

// Exhaustive search (reference encoder/me.c:483-634, --me esa).  The reference prunes with the
// successive-elimination lower bound (ads) and skips rows whose y-cost alone is not below the best;
// both are pure accelerations of a raster scan (y outer, x inner) with strict-< updates, which is
// what is restated here.  The scanned width is rounded to a multiple of 4 exactly as the reference
// rounds it (me.c:491), so up to 3 columns right of max_x are visited and the last one may be cut.
PCAMV_FN void esa_search(MeSearch &s)
{
    const int range = s.env.me_range;
    const int min_x = imax(s.bmx - range, s.mv_x_min), min_y = imax(s.bmy - range, s.mv_y_min);
    const int max_x = imin(s.bmx + range, s.mv_x_max), max_y = imin(s.bmy + range, s.mv_y_max);
    const int width = (max_x - min_x + 3) & ~3;
    for (int my = min_y; my <= max_y; my++)
        for (int x = 0; x < width; x += 4)
            s.try_x4(min_x + x, my, 0, 0, 1, 0, 2, 0, 3, 0);
}

PCAMV_FN void me_search_ref(const MeEnv &env, const MeBlock &b, const int (*mvc)[2], int i_mvc,
                             int *p_halfpel_thresh, MeResult &m)
{
    MeSearch s(env, b);
    int bpred_mx = 0, bpred_my = 0, bpred_cost = PCAMV_COST_MAX;

    s.bmx = clip3(b.mvp[0], s.mv_x_min * 4, s.mv_x_max * 4);
    s.bmy = clip3(b.mvp[1], s.mv_y_min * 4, s.mv_y_max * 4);
    const int pmx = (s.bmx + 2) >> 2, pmy = (s.bmy + 2) >> 2;
    s.bcost = PCAMV_COST_MAX;

    if (env.subme >= 3)
    {
        // predictors at quarter-pel precision: mvp first, then each distinct non-zero candidate
        int qx[10], qy[10], n = 0;
        qx[n] = s.bmx; qy[n] = s.bmy; n++;
        for (int i = 0; i < i_mvc; i++)
        {
            const bool nonzero = (mvc[i][0] | mvc[i][1]) != 0;
            const bool differs = ((mvc[i][0] & 0xffff) != (s.bmx & 0xffff)) || ((mvc[i][1] & 0xffff) != (s.bmy & 0xffff));
            if (nonzero && differs)
            {
                qx[n] = clip3(mvc[i][0], s.mv_x_min * 4, s.mv_x_max * 4);
                qy[n] = clip3(mvc[i][1], s.mv_y_min * 4, s.mv_y_max * 4);
                n++;
            }
        }
        for (int k0 = 0; k0 < n; k0 += 4)
        {
            const int nn = imin(4, n - k0);
            int sad[4];
            sad_cands(b, nn, qx + k0, qy + k0, sad);
            for (int k = 0; k < nn; k++)
            {
                const int c = sad[k] + b.cost_mvx[qx[k0 + k]] + b.cost_mvy[qy[k0 + k]];
                if (c < bpred_cost) { bpred_cost = c; bpred_mx = qx[k0 + k]; bpred_my = qy[k0 + k]; }
            }
        }
        s.bmx = (bpred_mx + 2) >> 2;
        s.bmy = (bpred_my + 2) >> 2;
        {
            // COST_MV(bmx,bmy) against bcost = COST_MAX: always taken
            int sad;
            const int mx = s.bmx, my = s.bmy;
            sad_fpel_cands(b, 1, &mx, &my, &sad);
            s.bcost = sad + s.bits_fpel(mx, my);
        }
    }
    else
    {
        {
            int sad;
            sad_fpel_cands(b, 1, &pmx, &pmy, &sad);
            s.bcost = sad;               // COST_MV then minus BITS_MVD(pmx,pmy)
            s.bmx = pmx; s.bmy = pmy;
        }
        for (int i = 0; i < i_mvc; i++)
        {
            int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
            if ((mx | my) && ((mx - s.bmx) | (my - s.bmy)))
            {
                mx = clip3(mx, s.mv_x_min, s.mv_x_max);
                my = clip3(my, s.mv_y_min, s.mv_y_max);
                s.try_fpel1(mx, my);
            }
        }
    }
    s.try_fpel1(0, 0);

    switch (env.me_method)
    {
    case ME_DIA: s.search_dia(env.me_range); break;
    case ME_HEX: s.hex_then_square(env.me_range); break;
    case ME_UMH: s.search_umh(pmx, pmy, mvc, i_mvc); break;
    default:     esa_search(s); break;
    }

    if (bpred_cost < s.bcost)
    {
        m.mv[0] = bpred_mx; m.mv[1] = bpred_my; m.cost = bpred_cost;
    }
    else
    {
        m.mv[0] = s.bmx << 2; m.mv[1] = s.bmy << 2; m.cost = s.bcost;
    }
    m.cost_mv = b.cost_mvx[m.mv[0]] + b.cost_mvy[m.mv[1]];
    if (s.bmx == pmx && s.bmy == pmy && env.subme < 3)
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
