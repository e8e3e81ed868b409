// pcamv_cost.cuh — the PCAMV candidate-MV cost table (x264_ih_get_mv_cost, reference encoder/analyse.c:2391-2550).
//
// For one motion vector of the final partitioning: reconstruct the whole macroblock (MC -> 4x4 DCT -> quant ->
// decimate -> dequant -> IDCT, luma + chroma; encoder/macroblock.c:605-755,277-372), cost the 9-point ring around
// the vector AGAINST THE RECONSTRUCTION (MV_SATD_FDEC_IH, analyse.c:2364-2385), then do the same for up to 12
// replacement vectors and keep the cheapest one whose "locally optimal" status matches the original's.
// One lane team handles one (macroblock, partition); lanes 0..15 own the luma 4x4 blocks, 16..23 the chroma ones.
#pragma once
#include "pcamv_frame.cuh"

namespace pcamv {

// x264_macroblock_encode for an inter MB whose partition k_over uses MV (omx, omy) instead of its own.
// Leaves the reconstruction in c.w.pred_y / pred_u / pred_v.
// partition k of the macroblock's final mode: the record holds up to four, P_8x8 macroblocks keep theirs in the side array
PCAMV_DEV const PartInfo &mb_part(const MbCtx &c, const MbResult &res, int k)
{
    return (res.type == MB_P_8x8 && c.fp.subparts) ? c.fp.subparts[(size_t)16 * c.mb_xy + k] : res.part[k];
}

// Transform / quantisation / decimation / reconstruction of the macroblock whose prediction is staged in c.w.pred_* (the part of
// x264_macroblock_encode behind the motion compensation, encoder/macroblock.c:690-755 luma, :277-372 chroma).  Leaves the
// reconstruction in c.w.pred_*.  Returns the luma 4x4 blocks (bit = block_idx) that keep coefficients after decimation: the
// non_zero_count the reference stores for them is non-zero (what the deblocking filter's bS = 2 test reads).
PCAMV_FN int encode_mb_residual(MbCtx &c)
{
    const DevTables &t = c.fc.tab;
    const int b_decimate = c.fc.b_dct_decimate;
    // scratch layout: [0..23] decimate score | nz << 8 per block; raw chroma DC terms behind them
    int *sc = c.w.scratch;
    int16_t *dcs = (int16_t *)(c.w.scratch + 24);           // 8 x int16
    PCAMV_FOR_ITEMS(it, 24)
        sc[it] = quant_block<0>(c, it, dcs);
    team_sync();

    // ---- decisions (uniform) -------------------------------------------------------------------------------
    int add_luma8 = 0;          // bit i: 8x8 block i gets its residual added
    {
        int s8[4], any8[4], mb = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            s8[i] = 0; any8[i] = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                const int v = sc[4 * i + j];
                any8[i] |= v >> 8;
                // the reference stops accumulating once an 8x8 reaches 6 (macroblock.c:712); totals compare the same
                s8[i] += v & 0xff;
            }
            mb += s8[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            if (b_decimate) { if (mb >= 6 && s8[i] >= 4) add_luma8 |= 1 << i; }
            else if (any8[i]) add_luma8 |= 1 << i;
        }
    }
    int chroma_mode[2];         // 0 = nothing, 1 = DC only, 2 = full
    int dcq[2][4];              // dequantised DC term per 4x4 block
#pragma unroll
    for (int pl = 0; pl < 2; pl++)
    {
        int nz_ac = 0, score = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { const int v = sc[16 + 4 * pl + j]; nz_ac |= v >> 8; score += v & 0xff; }
        const int b0 = dcs[4 * pl], b1 = dcs[4 * pl + 1], b2 = dcs[4 * pl + 2], b3 = dcs[4 * pl + 3];
        const int D0 = b0 + b1, D1 = b2 + b3, D2 = b0 - b1, D3 = b2 - b3;
        const int mf = t.quant4_mf[1][0] >> 1, bias = t.quant4_bias[1][0] << 1;
        const int q00 = quant_one((int16_t)(D0 + D1), mf, bias), q10 = quant_one((int16_t)(D2 + D3), mf, bias);
        const int q01 = quant_one((int16_t)(D0 - D1), mf, bias), q11 = quant_one((int16_t)(D2 - D3), mf, bias);
        const int nz_dc = (q00 | q01 | q10 | q11) != 0;
        const int e0 = q00 + q01, e1 = q10 + q11, e2 = q00 - q01, e3 = q10 - q11;
        int dmf = t.dequant4_mf[1][(t.chroma_qp % 6) * 16], qbits = t.chroma_qp / 6 - 5;
        if (qbits > 0) { dmf <<= qbits; qbits = 0; }
        dcq[pl][0] = (int16_t)(((e0 + e1) * dmf) >> -qbits); dcq[pl][1] = (int16_t)(((e0 - e1) * dmf) >> -qbits);
        dcq[pl][2] = (int16_t)(((e2 + e3) * dmf) >> -qbits); dcq[pl][3] = (int16_t)(((e2 - e3) * dmf) >> -qbits);
        if ((b_decimate && score < 7) || !nz_ac)
            chroma_mode[pl] = nz_dc ? 1 : 0;
        else
        {
            chroma_mode[pl] = 2;
            if (!nz_dc) dcq[pl][0] = dcq[pl][1] = dcq[pl][2] = dcq[pl][3] = 0;
        }
    }
    team_sync();

    // ---- reconstruction --------------------------------------------------------------------------------------
    PCAMV_FOR_ITEMS(it, 24)
    {
        if (it < 16)
        {
            // a block whose quantised coefficients are all zero adds nothing (its w.coef entry is stale)
            if (((add_luma8 >> (it >> 2)) & 1) && (sc[it] >> 8))
                recon_block(c, it, 0, 0);
        }
        else
        {
            const int pl = (it - 16) >> 2, blk = (it - 16) & 3;
            const int mode = pl ? chroma_mode[1] : chroma_mode[0];
            const int dc = pl ? (blk == 0 ? dcq[1][0] : blk == 1 ? dcq[1][1] : blk == 2 ? dcq[1][2] : dcq[1][3])
                              : (blk == 0 ? dcq[0][0] : blk == 1 ? dcq[0][1] : blk == 2 ? dcq[0][2] : dcq[0][3]);
            if (mode == 1)
                recon_block(c, it, 1, dc);
            else if (mode == 2)
            {
                // AC part all zero: only the DC term contributes, through the full IDCT path
                if (!(sc[it] >> 8))
                {
                    int16_t *co = c.w.coef[it];
                    for (int i = 0; i < 16; i++) co[i] = 0;
                }
                recon_block(c, it, 2, dc);
            }
        }
    }
    team_sync();
    int luma_mask = 0;
#pragma unroll 1
    for (int it = 0; it < 16; it++)
        if (((add_luma8 >> (it >> 2)) & 1) && (sc[it] >> 8)) luma_mask |= 1 << it;
    return luma_mask;
}


// x264_macroblock_encode for an inter MB whose partition k_over uses MV (omx, omy) instead of its own.
// Leaves the reconstruction in c.w.pred_y / pred_u / pred_v.
PCAMV_FN void encode_mb_inter(MbCtx &c, const MbResult &res, int k_over, int omx, int omy)
{
    for (int p = 0; p < res.n_part; p++)
    {
        const PartInfo &pi = mb_part(c, res, p);
        const int mx = clip3(p == k_over ? omx : pi.mv[0], c.mv_min[0], c.mv_max[0]);
        const int my = clip3(p == k_over ? omy : pi.mv[1], c.mv_min[1], c.mv_max[1]);
        mc_rect(c, c.fp.ref_slot[pi.ref], pi.xoff, pi.yoff, pix_w(pi.i_pixel), pix_h(pi.i_pixel), mx, my);
    }
    encode_mb_residual(c);
}

// dx / dy of replacement candidate ii = 0..11 (four at distance 1, eight knight moves), one signed nibble each:
//   dx = 0 1 0 -1 -2 -1 1 2 2 1 -1 -2      dy = -1 0 1 0 1 2 2 1 -1 -2 -2 -1
PCAMV_DEV int cand_dx(int ii) { return (int)(((long long)(0xEF1221FEF010ull << (60 - 4 * ii))) >> 60); }
PCAMV_DEV int cand_dy(int ii) { return (int)(((long long)(0xFEEF1221010Full << (60 - 4 * ii))) >> 60); }

// ---- x264_ih_get_mv_cost (analyse.c:2391-2550), in two parts so that the candidates can be costed by different teams ----
// (1) one vector: reconstruct the macroblock with it, cost its 9-point ring against the reconstruction
struct IhCand { int centre, min9, r[4]; };        // r[] = the four distance-1 ring costs (used for the original vector only)

// ii = -1: the original vector of partition k; ii = 0..11: replacement candidate ii.  c.w.blk must describe partition k with
// its source-pixel pointers aimed at c.w.pred_* (ih_setup_block).
PCAMV_FN void ih_eval(MbCtx &c, const MbResult &res, int k, int ii, int kind, int bmx, int bmy, IhCand &o)
{
    MeBlock &b = c.w.blk;
    const int cx = bmx + (ii < 0 ? 0 : cand_dx(ii)), cy = bmy + (ii < 0 ? 0 : cand_dy(ii));
    encode_mb_inter(c, res, ii < 0 ? -1 : k, cx, cy);
    // 9-point ring around (cx, cy) against the reconstruction: up, right, down, left, then the diagonals, then the centre
    int r[4];
    eval4(b, kind, 4, pk(cx, cy - 1), pk(cx + 1, cy), pk(cx, cy + 1), pk(cx - 1, cy), r);
    int min9 = imin(imin(r[0], r[1]), imin(r[2], r[3]));
    o.r[0] = r[0]; o.r[1] = r[1]; o.r[2] = r[2]; o.r[3] = r[3];
    eval4(b, kind, 4, pk(cx - 1, cy - 1), pk(cx - 1, cy + 1), pk(cx + 1, cy - 1), pk(cx + 1, cy + 1), r);
    min9 = imin(min9, imin(imin(r[0], r[1]), imin(r[2], r[3])));
    o.centre = eval1(b, kind, pk(cx, cy));
    o.min9 = imin(min9, o.centre);
}

PCAMV_DEV int ih_setup_block(MbCtx &c, const MbResult &res, int k)
{
    const PartInfo &pi = mb_part(c, res, k);
    MeBlock &b = c.w.blk;
    setup_block(c, b, pi.ref, pi.i_pixel, pi.xoff, pi.yoff);
    block_set_mvp(b, c.env, pi.mvp[0], pi.mvp[1]);
    // the block being compared is the RECONSTRUCTION of this partition, not the source
    b.fenc = c.w.pred_y + pi.yoff * 16 + pi.xoff;
    b.fenc_u = c.w.pred_u + (pi.yoff >> 1) * 8 + (pi.xoff >> 1);
    b.fenc_v = c.w.pred_v + (pi.yoff >> 1) * 8 + (pi.xoff >> 1);
    return !c.env.mbcmp_satd ? COST_SAD : (c.env.chroma_me && pi.i_pixel <= PIX_8x8) ? COST_SATD_CHROMA : COST_SATD;
}

// (2) the selection, folded over the candidates in the reference's order
struct IhFold { int orig, non_opt, fb_cost, fb_idx, best, ii_best, m_x, m_y; };
PCAMV_DEV void ih_fold_orig(IhFold &f, const IhCand &o)
{
    f.fb_cost = PCAMV_COST_MAX; f.fb_idx = -1; f.best = PCAMV_COST_MAX; f.ii_best = -1; f.m_x = 0; f.m_y = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (o.r[i] < f.fb_cost) { f.fb_cost = o.r[i]; f.fb_idx = i; }
    f.orig = o.centre;
    f.non_opt = o.min9 < f.orig;          // the original vector is not a local optimum of its ring
}
// returns 1 when the reference stops looking at further candidates (after the first four, if one of them qualified)
PCAMV_DEV int ih_fold_cand(IhFold &f, int ii, const IhCand &o)
{
    const int qualifies = f.non_opt ? (o.min9 != o.centre) : (o.min9 == o.centre);
    if (qualifies && o.centre < f.best) { f.best = o.centre; f.m_x = cand_dx(ii); f.m_y = cand_dy(ii); f.ii_best = ii; }
    return ii == 3 && f.best != PCAMV_COST_MAX;
}
PCAMV_DEV int ih_fold_finish(IhFold &f)
{
    int b_1_neighbor, b_error_pos = 0;
    if (f.best == PCAMV_COST_MAX)
    {
        b_error_pos = 1; b_1_neighbor = 1;
        f.m_x = 0; f.m_y = 0;
        if (f.fb_idx >= 0) { f.best = f.fb_cost; f.m_x = cand_dx(f.fb_idx); f.m_y = cand_dy(f.fb_idx); }
    }
    else
        b_1_neighbor = f.ii_best <= 3;
    int cost_opt = f.best > f.orig ? f.best - f.orig : 1;
    if (!b_1_neighbor)
        cost_opt = (int)(1.4f * (float)cost_opt);
    else if (b_error_pos)
        cost_opt = (int)(4.0f * (float)cost_opt);
    return cost_opt;
}

// x264_ih_get_mv_cost for partition k of macroblock `res` by ONE team; returns cost_opt, writes the chosen delta
PCAMV_FN int ih_get_mv_cost(MbCtx &c, const MbResult &res, int k, int &m_x, int &m_y)
{
    const PartInfo &pi = mb_part(c, res, k);
    const int bmx = pi.mv[0], bmy = pi.mv[1];
    const int kind = ih_setup_block(c, res, k);
    IhFold f;
    IhCand o;
    // ii = -1: the original vector; ii = 0..11: the replacement candidates (the first four decide whether the rest run)
#pragma unroll 1
    for (int ii = -1; ii < 12; ii++)
    {
        ih_eval(c, res, k, ii, kind, bmx, bmy, o);
        if (ii < 0) ih_fold_orig(f, o);
        else if (ih_fold_cand(f, ii, o)) break;
    }
    const int cost_opt = ih_fold_finish(f);
    m_x = f.m_x; m_y = f.m_y;
    return cost_opt;
}

// cost-table entries of one macroblock (all its partitions), appended to the MB's log after the search entries
PCAMV_FN void cost_table_mb(MbCtx &c, MbResult &res)
{
    if (res.type == MB_P_SKIP)
        return;
    c.n_log = res.n_log;
    init_limits(c);
    c.env.cost_mv = c.fc.tab.cost_mv;
    c.env.me_method = c.fc.me_method; c.env.me_range = c.fc.me_range; c.env.subme = c.fc.subme;
    c.env.chroma_me = c.fc.chroma_me && c.fc.subme >= 5;
    c.env.mbcmp_satd = c.fc.subme > 1;
    c.env.mvsads = nullptr;
    for (int k = 0; k < res.n_part; k++)
    {
        int m_x, m_y;
        const int cost_opt = ih_get_mv_cost(c, res, k, m_x, m_y);
        log_push(c, LOG_IHCOST, mb_part(c, res, k).i_pixel, mb_part(c, res, k).ref, m_x, m_y, cost_opt, 0);
    }
    if (team_lane() == 0)
        c.fp.results[c.mb_xy].n_log = c.n_log;
}

} // namespace pcamv
