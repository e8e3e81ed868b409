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

#if defined(PCAMV_EMU)
  #define PCAMV_COEF_SLOTS 24
  #define PCAMV_SLOT(it) (it)
#else
  #define PCAMV_COEF_SLOTS 1
  #define PCAMV_SLOT(it) 0
#endif

// x264_macroblock_encode for an inter MB whose partition k_over uses MV (omx, omy) instead of its own.
// Leaves the reconstruction in c.w.pred_y / pred_u / pred_v.
PCAMV_FN void encode_mb_inter(MbCtx &c, const MbResult &res, int k_over, int omx, int omy)
{
    const DevTables &t = c.fc.tab;
    const int b_decimate = c.fc.b_dct_decimate;
    for (int p = 0; p < res.n_part; p++)
    {
        const PartInfo &pi = res.part[p];
        const int mx = clip3(p == k_over ? omx : pi.mv[0], c.mv_min[0], c.mv_max[0]);
        const int my = clip3(p == k_over ? omy : pi.mv[1], c.mv_min[1], c.mv_max[1]);
        mc_rect(c, c.fp.ref_slot[pi.ref], pi.xoff, pi.yoff, pix_w(pi.i_pixel), pix_h(pi.i_pixel), mx, my);
    }
    int coef[PCAMV_COEF_SLOTS][16];
    // scratch layout: [0..23] decimate score, bits: score | nz << 8 ; chroma DC terms in dcs
    int *sc = c.w.scratch;
    int16_t *dcs = (int16_t *)(c.w.scratch + 24);           // 8 x int16
    PCAMV_FOR_ITEMS(it, 24)
    {
        int *co = coef[PCAMV_SLOT(it)];
        int d[16];
        if (it < 16)
        {
            const int bx = (it & 1) | ((it >> 1) & 2), by = ((it >> 1) & 1) | ((it >> 2) & 2);
            load_residual(c.w.fenc_y + 64 * by + 4 * bx, 16, c.w.pred_y + 64 * by + 4 * bx, 16, d);
            dct4x4(d, co);
            const int nz = quant4x4(co, t.quant4_mf[0], t.quant4_bias[0]);
            int score = 0;
            if (nz)
            {
                score = decimate_score(co, 0);
                dequant4x4(co, t.dequant4_mf[0], t.qp);
            }
            sc[it] = score | (nz << 8);
        }
        else
        {
            const int pl = (it - 16) >> 2, blk = (it - 16) & 3;
            const uint8_t *fe = pl ? c.w.fenc_v : c.w.fenc_u, *pr = pl ? c.w.pred_v : c.w.pred_u;
            load_residual(fe + 32 * (blk >> 1) + 4 * (blk & 1), 8, pr + 32 * (blk >> 1) + 4 * (blk & 1), 8, d);
            dct4x4(d, co);
            dcs[it - 16] = (int16_t)co[0];
            co[0] = 0;
            const int nz = quant4x4(co, t.quant4_mf[1], t.quant4_bias[1]);
            int score = 0;
            if (nz)
            {
                score = decimate_score(co, 1);
                dequant4x4(co, t.dequant4_mf[1], t.chroma_qp);
            }
            sc[it] = score | (nz << 8);
        }
    }
    team_sync();

    // ---- decisions (uniform) -------------------------------------------------------------------------------
    int add_luma8 = 0;          // bit i: 8x8 block i gets its residual added
    {
        int s8[4], any8[4], mb = 0;
        for (int i = 0; i < 4; i++)
        {
            s8[i] = 0; any8[i] = 0;
            for (int j = 0; j < 4; j++)
            {
                const int v = sc[4 * i + j];
                any8[i] |= v >> 8;
                // the reference stops accumulating once an 8x8 reaches 6 (macroblock.c:712); totals compare the same
                s8[i] += v & 0xff;
            }
            mb += s8[i];
        }
        for (int i = 0; i < 4; i++)
        {
            if (b_decimate) { if (mb >= 6 && s8[i] >= 4) add_luma8 |= 1 << i; }
            else if (any8[i]) add_luma8 |= 1 << i;
        }
    }
    int chroma_mode[2];         // 0 = nothing, 1 = DC only, 2 = full
    int dcq[2][4];              // dequantised DC term per 4x4 block
    for (int pl = 0; pl < 2; pl++)
    {
        int nz_ac = 0, score = 0;
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
        int *co = coef[PCAMV_SLOT(it)];
        int r[16];
        if (it < 16)
        {
            if (!((add_luma8 >> (it >> 2)) & 1)) continue;
            const int bx = (it & 1) | ((it >> 1) & 2), by = ((it >> 1) & 1) | ((it >> 2) & 2);
            idct4x4(co, r);
            add_residual(c.w.pred_y + 64 * by + 4 * bx, 16, r);
        }
        else
        {
            const int pl = (it - 16) >> 2, blk = (it - 16) & 3;
            uint8_t *pr = (pl ? c.w.pred_v : c.w.pred_u) + 32 * (blk >> 1) + 4 * (blk & 1);
            if (chroma_mode[pl] == 0) continue;
            if (chroma_mode[pl] == 1)
            {
                const int dc = (int16_t)((dcq[pl][blk] + 32) >> 6);
                for (int k = 0; k < 16; k++) r[k] = dc;
            }
            else
            {
                co[0] = dcq[pl][blk];
                idct4x4(co, r);
            }
            add_residual(pr, 8, r);
        }
    }
    team_sync();
}

// costs of the 9-point ring around (cx, cy) for partition k against the current reconstruction
PCAMV_FN void ring_costs(MbCtx &c, const MbResult &res, int k, int cx, int cy, int out[9])
{
    const PartInfo &pi = res.part[k];
    MeBlock b;
    setup_block(c, b, pi.ref, pi.i_pixel, pi.xoff, pi.yoff);
    block_set_mvp(b, c.env, pi.mvp[0], pi.mvp[1]);
    const int rx[9] = { 0, 1, 0, -1, -1, -1, 1, 1, 0 }, ry[9] = { -1, 0, 1, 0, -1, 1, -1, 1, 0 };
    int qx[9], qy[9];
    for (int i = 0; i < 9; i++) { qx[i] = cx + rx[i]; qy[i] = cy + ry[i]; }
    const int chroma = c.env.chroma_me && pi.i_pixel <= PIX_8x8;
    // the block being compared is the RECONSTRUCTION of this partition, not the source
    b.fenc = c.w.pred_y + pi.yoff * 16 + pi.xoff;
    b.fenc_u = c.w.pred_u + (pi.yoff >> 1) * 8 + (pi.xoff >> 1);
    b.fenc_v = c.w.pred_v + (pi.yoff >> 1) * 8 + (pi.xoff >> 1);
    satd_cands(b, 9, qx, qy, chroma, out, b.fenc, b.fenc_u, b.fenc_v, c.env.mbcmp_satd);
    for (int i = 0; i < 9; i++) out[i] += b.cost_mvx[qx[i]] + b.cost_mvy[qy[i]];
}

// x264_ih_get_mv_cost for partition k of macroblock `res`; returns cost_opt, writes the chosen delta
PCAMV_FN int ih_get_mv_cost(MbCtx &c, const MbResult &res, int k, int &m_x, int &m_y)
{
    const int dmx[12] = { 0, 1, 0, -1, -2, -1, 1, 2, 2, 1, -1, -2 }, dmy[12] = { -1, 0, 1, 0, 1, 2, 2, 1, -1, -2, -2, -1 };
    const int rx[4] = { 0, 1, 0, -1 }, ry[4] = { -1, 0, 1, 0 };
    const int bmx = res.part[k].mv[0], bmy = res.part[k].mv[1];
    int c1[9];
    encode_mb_inter(c, res, -1, 0, 0);
    ring_costs(c, res, k, bmx, bmy, c1);
    int min_cost = PCAMV_COST_MAX;
    for (int i = 0; i < 9; i++) if (c1[i] < min_cost) min_cost = c1[i];
    const int orig = c1[8];
    const int non_opt = min_cost < orig;        // the original vector is not a local optimum of its ring
    int best = PCAMV_COST_MAX, ii_best = -1;
    m_x = 0; m_y = 0;
    for (int ii = 0; ii < 12; ii++)
    {
        const int cx = bmx + dmx[ii], cy = bmy + dmy[ii];
        int c2[9];
        encode_mb_inter(c, res, k, cx, cy);
        ring_costs(c, res, k, cx, cy, c2);
        int m1 = PCAMV_COST_MAX;
        for (int i = 0; i < 9; i++) if (c2[i] < m1) m1 = c2[i];
        const int cost = c2[8];
        const int qualifies = non_opt ? (m1 != cost) : (m1 == cost);
        if (qualifies && cost < best) { best = cost; m_x = dmx[ii]; m_y = dmy[ii]; ii_best = ii; }
        if (ii == 3 && best != PCAMV_COST_MAX)
            break;
    }
    int b_1_neighbor, b_error_pos = 0;
    if (best == PCAMV_COST_MAX)
    {
        b_error_pos = 1; b_1_neighbor = 1;
        m_x = 0; m_y = 0;
        for (int i = 0; i < 4; i++)
            if (c1[i] < best) { best = c1[i]; m_x = rx[i]; m_y = ry[i]; }
    }
    else
        b_1_neighbor = ii_best <= 3;
    int cost_opt = best > orig ? best - orig : 1;
    if (!b_1_neighbor)
        cost_opt = (int)(1.4f * (float)cost_opt);
    else if (b_error_pos)
        cost_opt = (int)(4.0f * (float)cost_opt);
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
    for (int k = 0; k < res.n_part; k++)
    {
        int m_x, m_y;
        const int cost_opt = ih_get_mv_cost(c, res, k, m_x, m_y);
        log_push(c, LOG_IHCOST, res.part[k].i_pixel, res.part[k].ref, m_x, m_y, cost_opt, 0);
    }
    if (team_lane() == 0)
        c.fp.results[c.mb_xy].n_log = c.n_log;
}

} // namespace pcamv
