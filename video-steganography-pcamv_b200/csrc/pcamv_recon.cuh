// pcamv_recon.cuh — reconstruction of a P frame in its final mode and the in-loop deblocking filter, on the device
// (SURVEY.md 8(f) row 2): the frame the encoder keeps as reference is rebuilt from the decisions the analysis left in HBM,
// so that the reference planes of the next frame (integer plane, then border + half-pel planes by the kernels of
// pcamv_kernels.cu) never have to come from the host.
//
// Behavioural contract (bit-exact with the reference's fdec planes after x264_fdec_filter_row):
//   reconstruction   encoder/macroblock.c:605-755 (inter macroblock: x264_mb_mc, 4x4 DCT / quant / decimation / dequant / IDCT),
//                    :277-372 (chroma), :387-411 (P_SKIP: motion compensation with the cached vector of block 0, reference 0)
//   deblocking       common/frame.c:627-800 (x264_frame_deblock_row: edge order, bS from non_zero_count / reference index /
//                    vector differences, the no-sub8x8 shortcut, P_SKIP and low-QP macroblocks filter their boundary edge only),
//                    :404-470 (deblock_luma_c / deblock_chroma_c), :370-402 (alpha / beta / tc0 tables = H.264 tables 8-16, 8-17)
// Not reproduced, by construction: a macroblock whose pass-2 probe found it skippable although pass 1 coded it keeps
// b_skip_mc set on the host, and the host then takes the residual against whatever its intra analysis left in
// fdec — state the device does not have.  Such macroblocks are rare (1 in ~10^4 on low-noise content, none on the bench
// clips); the caller decides per frame whether to trust this path (pcamv_reconstruct_ref reports nothing about them — the
// host glue's check mode compares against the host's planes).
#pragma once
#include "pcamv_cost.cuh"

namespace pcamv {

struct ReconPlanes
{
    uint8_t *y, *u, *v;          // pixel (0,0) of the padded planes being built (reference-slot layout)
    int stride_y, stride_c;
    uint16_t *nnz;               // [n_mb]: bit (x + 4 y) = luma 4x4 block at raster position (x, y) keeps coefficients
};

struct DeblockParams
{
    int disable;                 // sh.i_disable_deblocking_filter_idc == 1
    int alpha_c0_offset, beta_offset;     // sh.i_alpha_c0_offset, sh.i_beta_offset (already doubled)
    int qp, qp_chroma;           // constant QP: every macroblock of the slice has these
    int chroma_qp_offset;
    int no_sub8x8_all;           // !(analyse.inter & X264_ANALYSE_PSUB8x8)
};

// pixels another SM may have written within this launch: bypass the non-coherent L1
#if defined(PCAMV_EMU)
  #define PCAMV_PX_LD(p) (*(p))
  #define PCAMV_PX_ST(p, v) (*(p) = (v))
  #define PCAMV_STV(p, v) (*(p) = (v))
#else
  #define PCAMV_PX_LD(p) __ldcg(p)
  #define PCAMV_PX_ST(p, v) __stcg((p), (v))
  #define PCAMV_STV(p, v) __stcg((p), (v))
#endif

// ---- reconstruction of one macroblock in its final mode ------------------------------------------------------------
// r = the record the analysis of the frame's last pass left for this macroblock (final cache vectors per 4x4 block, references
// per 8x8 block, type).  The source pixels must be staged in c.w.fenc_*.
PCAMV_FN void recon_mb(MbCtx &c, const MbResult &r, const ReconPlanes &rp)
{
    init_limits(c);
    if (r.type == MB_P_SKIP)
    {
        const int mvx = clip3(mv_x(r.mv[0]), c.mv_min[0], c.mv_max[0]), mvy = clip3(mv_y(r.mv[0]), c.mv_min[1], c.mv_max[1]);
        mc_rect(c, c.fp.ref_slot[0], 0, 0, 16, 16, mvx, mvy);
    }
    else
    {
#pragma unroll 1
        for (int i8 = 0; i8 < 4; i8++)
        {
            const int slot = c.fp.ref_slot[r.ref[i8]];
            const uint32_t m0 = r.mv[4 * i8];
            if (r.mv[4 * i8 + 1] == m0 && r.mv[4 * i8 + 2] == m0 && r.mv[4 * i8 + 3] == m0)
                mc_rect(c, slot, 8 * (i8 & 1), 8 * (i8 >> 1), 8, 8, clip3(mv_x(m0), c.mv_min[0], c.mv_max[0]), clip3(mv_y(m0), c.mv_min[1], c.mv_max[1]));
            else
#pragma unroll 1
                for (int j = 0; j < 4; j++)
                {
                    const uint32_t m = r.mv[4 * i8 + j];
                    mc_rect(c, slot, 8 * (i8 & 1) + 4 * (j & 1), 8 * (i8 >> 1) + 4 * (j >> 1), 4, 4,
                            clip3(mv_x(m), c.mv_min[0], c.mv_max[0]), clip3(mv_y(m), c.mv_min[1], c.mv_max[1]));
                }
        }
    }
    int mask = 0;
    if (r.type != MB_P_SKIP)
        mask = encode_mb_residual(c);
    team_sync();
    // the reconstructed macroblock into the frame
    {
        uint8_t *dy = rp.y + (size_t)(16 * c.mb_y) * rp.stride_y + 16 * c.mb_x;
        uint8_t *du = rp.u + (size_t)(8 * c.mb_y) * rp.stride_c + 8 * c.mb_x, *dv = rp.v + (size_t)(8 * c.mb_y) * rp.stride_c + 8 * c.mb_x;
        PCAMV_FOR_ITEMS(it, 64 + 32)
        {
            if (it < 64)
            {
                const int y = it >> 2, x = (it & 3) << 2;
                st4a(dy + (size_t)y * rp.stride_y + x, ld4a(c.w.pred_y + 16 * y + x));
            }
            else
            {
                const int k = it - 64, pl = k >> 4, y = (k >> 1) & 7, x = (k & 1) << 2;
                st4a((pl ? dv : du) + (size_t)y * rp.stride_c + x, ld4a((pl ? c.w.pred_v : c.w.pred_u) + 8 * y + x));
            }
        }
    }
    if (team_lane() == 0)
    {
        int raster = 0;
#pragma unroll 1
        for (int it = 0; it < 16; it++)
            if ((mask >> it) & 1)
                raster |= 1 << (((it & 1) | ((it >> 1) & 2)) + 4 * (((it >> 1) & 1) | ((it >> 2) & 2)));
        rp.nnz[c.mb_xy] = (uint16_t)raster;
    }
    team_sync();
}

// ---- deblocking ------------------------------------------------------------------------------------------------------
// H.264 tables 8-16 / 8-17, indexed by qp + offset in [-12, 63]
PCAMV_DEV int db_alpha(int i)
{
    i = clip3(i, 0, 51);
    const unsigned char tab[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,
                                    80,90,101,113,127,144,162,182,203,226,255,255 };
    return tab[i];
}
PCAMV_DEV int db_beta(int i)
{
    i = clip3(i, 0, 51);
    const unsigned char tab[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,
                                    13,13,14,14,15,15,16,16,17,17,18,18 };
    return tab[i];
}
PCAMV_DEV int db_tc0(int i, int bS)
{
    i = clip3(i, 0, 51);
    const unsigned char t1[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13 };
    const unsigned char t2[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,5,5,6,7,8,8,10,11,12,13,15,17 };
    const unsigned char t3[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13,14,16,18,20,23,25 };
    return bS == 1 ? t1[i] : bS == 2 ? t2[i] : t3[i];
}

struct DbFrame               // what the filter reads of the frame's final motion state
{
    const int8_t *type;      // [n_mb]
    const int8_t *ref8;      // [2 mb_h][2 mb_w]
    const uint32_t *mv4;     // [4 mb_h][4 mb_w] packed
    const uint16_t *nnz;     // [n_mb]
    int mb_w;
};

// one line across a luma edge (deblock_luma_c): pix = q0, xs = step across the edge; works on the staged copy of the
// macroblock's neighbourhood (plain loads and stores: team-private memory)
PCAMV_DEV void db_luma_line(uint8_t *pix, int xs, int alpha, int beta, int tc0)
{
    const int p2 = pix[-3 * xs], p1 = pix[-2 * xs], p0 = pix[-xs];
    const int q0 = pix[0], q1 = pix[xs], q2 = pix[2 * xs];
    if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)
    {
        int tc = tc0;
        if (iabs(p2 - p0) < beta)
        {
            pix[-2 * xs] = (uint8_t)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0));
            tc++;
        }
        if (iabs(q2 - q0) < beta)
        {
            pix[xs] = (uint8_t)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0));
            tc++;
        }
        const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        pix[-xs] = (uint8_t)clip_u8(p0 + delta);
        pix[0] = (uint8_t)clip_u8(q0 - delta);
    }
}
PCAMV_DEV void db_chroma_line(uint8_t *pix, int xs, int alpha, int beta, int tc)
{
    const int p1 = pix[-2 * xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs];
    if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)
    {
        const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        pix[-xs] = (uint8_t)clip_u8(p0 + delta);
        pix[0] = (uint8_t)clip_u8(q0 - delta);
    }
}

// Boundary strength of piece i (4 pixels) of edge e in direction dir (0 = vertical edges) of macroblock (mb_x, mb_y), it = 16 dir
// + 4 e + i (common/frame.c DEBLOCK_STRENGTH; inter macroblocks only).  0 for an edge that is not filtered at all: the frame
// border, the inner edges of a P_SKIP macroblock or of a slice whose QP is too low to change anything.  Depends on the frame's
// final motion state and coefficient flags only — not on any filtered pixel — so the strengths of the whole frame are computed
// ahead of the filter wavefront, one thread per piece.
PCAMV_DEV int db_raw_strength(const DbFrame &f, int mb_x, int mb_y, int it, int edge_end, int no_sub8x8)
{
    const int mb_xy = mb_y * f.mb_w + mb_x;
    const int dir = it >> 4, e = (it >> 2) & 3, i = it & 3;
    if ((e == 0 && (dir ? mb_y == 0 : mb_x == 0)) || (e >= 1 && e >= edge_end))
        return 0;
    const int s8 = 2 * f.mb_w, s4 = 4 * f.mb_w;
    const int mbn_xy = e ? mb_xy : (dir == 0 ? mb_xy - 1 : mb_xy - f.mb_w);
    const int nx = e ? mb_x : (dir == 0 ? mb_x - 1 : mb_x), ny = e ? mb_y : (dir == 0 ? mb_y : mb_y - 1);
    const unsigned nz_p = PCAMV_LDV(f.nnz + mb_xy), nz_q = PCAMV_LDV(f.nnz + mbn_xy);
    const int x = dir == 0 ? e : i, y = dir == 0 ? i : e;
    const int xn = dir == 0 ? (x - 1) & 3 : x, yn = dir == 0 ? y : (y - 1) & 3;
    if (((nz_p >> (x + 4 * y)) & 1) || ((nz_q >> (xn + 4 * yn)) & 1))
        return 2;
    if (e & no_sub8x8)
        return 0;
    // 0x80: "the reference would compare here" — unless it copies the previous piece's strength (db_strength_piece)
    const int i8p = (2 * mb_y + (y >> 1)) * s8 + 2 * mb_x + (x >> 1), i8q = (2 * ny + (yn >> 1)) * s8 + 2 * nx + (xn >> 1);
    const int i4p = (4 * mb_y + y) * s4 + 4 * mb_x + x, i4q = (4 * ny + yn) * s4 + 4 * nx + xn;
    const uint32_t mp = PCAMV_LDV(f.mv4 + i4p), mq = PCAMV_LDV(f.mv4 + i4q);
    const int differ = PCAMV_LDV(f.ref8 + i8p) != PCAMV_LDV(f.ref8 + i8q) || iabs(mv_x(mp) - mv_x(mq)) >= 4 || iabs(mv_y(mp) - mv_y(mq)) >= 4;
    return 0x80 | differ;
}
PCAMV_DEV int db_strength_piece(const DbFrame &f, const DeblockParams &dp, int mb_x, int mb_y, int it)
{
    const int type = PCAMV_LDV(f.type + mb_y * f.mb_w + mb_x);
    const int qp_thresh = 15 - imin(dp.alpha_c0_offset, dp.beta_offset) - imax(0, dp.chroma_qp_offset);
    const int edge_end = (type == MB_P_SKIP || dp.qp <= qp_thresh) ? 1 : 4;
    const int no_sub8x8 = (type != MB_P_8x8 || dp.no_sub8x8_all) ? 1 : 0;
    int bs = db_raw_strength(f, mb_x, mb_y, it, edge_end, no_sub8x8);
    if (bs & 0x80)
    {
        bs &= 0x7f;
        // an odd piece of a macroblock without sub-8x8 partitions takes the strength of the piece before it unless that is 2
        // (the reference's loop carries `prev`; an even piece never copies)
        if ((it & 1) && no_sub8x8)
        {
            const int prev = db_raw_strength(f, mb_x, mb_y, it - 1, edge_end, no_sub8x8) & 0x7f;
            if (prev != 2) bs = prev;
        }
    }
    return bs;
}

// Team-private staging of one macroblock's filter neighbourhood: luma rows -4..15 x columns -4..15 (stride 24), U and V rows
// -4..7 x columns -4..7 (stride 16), and the 32 boundary strengths of its 2 x 4 edges x 4 pieces.
#define DB_LS 24
#define DB_CS 16
#define DB_STAGE_Y 0
#define DB_STAGE_C (20 * DB_LS)
#define DB_STAGE_BS (DB_STAGE_C + 2 * 12 * DB_CS)
#define DB_STAGE_BYTES (DB_STAGE_BS + 32)

// All edges of macroblock (mb_x, mb_y), in the reference's order: vertical edges left to right (luma, and chroma on the even
// ones), then horizontal edges top to bottom.  The caller guarantees that the left neighbour and the row above up to the
// top-right neighbour are finished (their pixels are read and written here) and that nobody else touches the macroblock's
// neighbourhood until this call returns — which the wavefront order gives: the right neighbour and the row below start after it.
// The neighbourhood and the macroblock's 32 boundary strengths (bs32, computed ahead: db_strength_piece) are copied into
// team-private memory once (one global round trip instead of one per edge), the edges are filtered in the copy and the copy is
// written back.
// Team layout while filtering: lanes 0..15 = the 16 luma lines of an edge, lanes 16..23 / 24..31 = the 8 lines of U / V.
PCAMV_FN void deblock_mb(const ReconPlanes &rp, const DeblockParams &dp, int mb_x, int mb_y, const uint8_t *bs32, uint8_t *stage)
{
    const int alpha = db_alpha(dp.qp + dp.alpha_c0_offset), beta = db_beta(dp.qp + dp.beta_offset);
    const int alpha_c = db_alpha(dp.qp_chroma + dp.alpha_c0_offset), beta_c = db_beta(dp.qp_chroma + dp.beta_offset);
    uint8_t *py = rp.y + (size_t)(16 * mb_y) * rp.stride_y + 16 * mb_x;
    uint8_t *pu = rp.u + (size_t)(8 * mb_y) * rp.stride_c + 8 * mb_x, *pv = rp.v + (size_t)(8 * mb_y) * rp.stride_c + 8 * mb_x;
    uint8_t *sy = stage + DB_STAGE_Y + 4 * DB_LS + 4;                   // (0, 0) of the macroblock in the luma copy
    uint8_t *sbs = stage + DB_STAGE_BS;

    // (1) copy in: 20 luma rows of 20 bytes, 12 rows of 12 bytes per chroma plane (the planes are padded: rows / columns -4..-1
    // exist in memory on the frame border too, and are not filtered there)
    PCAMV_FOR_ITEMS(it, 32)
    {
        if (it < 8)
            ((uint32_t *)sbs)[it] = PCAMV_LDV((const uint32_t *)bs32 + it);
        if (it < 20)
        {
            const uint32_t *src = (const uint32_t *)(py + (ptrdiff_t)(it - 4) * rp.stride_y - 4);
            uint32_t *dst = (uint32_t *)(stage + DB_STAGE_Y + it * DB_LS);
#pragma unroll
            for (int k = 0; k < 5; k++) dst[k] = PCAMV_LDV(src + k);
        }
        else
        {
            const int r = it - 20;
#pragma unroll
            for (int pl = 0; pl < 2; pl++)
            {
                const uint32_t *src = (const uint32_t *)((pl ? pv : pu) + (ptrdiff_t)(r - 4) * rp.stride_c - 4);
                uint32_t *dst = (uint32_t *)(stage + DB_STAGE_C + pl * 12 * DB_CS + r * DB_CS);
#pragma unroll
                for (int k = 0; k < 3; k++) dst[k] = PCAMV_LDV(src + k);
            }
        }
    }
    // (3) the edges, in the copy
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++)
    {
#pragma unroll 1
        for (int e = 0; e < 4; e++)
        {
            team_sync();             // an edge reads what the previous edge (and, for dir 1, the vertical edges of all lines) wrote
            const uint32_t bS = (uint32_t)sbs[16 * dir + 4 * e] | ((uint32_t)sbs[16 * dir + 4 * e + 1] << 8) |
                                ((uint32_t)sbs[16 * dir + 4 * e + 2] << 16) | ((uint32_t)sbs[16 * dir + 4 * e + 3] << 24);
            if (bS)
            {
                PCAMV_FOR_ITEMS(it, 32)
                {
                    if (it < 16)
                    {
                        const int bs = (bS >> (8 * (it >> 2))) & 255;
                        if (bs && alpha && beta)
                        {
                            uint8_t *q = dir == 0 ? sy + it * DB_LS + 4 * e : sy + (4 * e) * DB_LS + it;
                            db_luma_line(q, dir == 0 ? 1 : DB_LS, alpha, beta, db_tc0(dp.qp + dp.alpha_c0_offset, bs));
                        }
                    }
                    else if (!(e & 1))
                    {
                        const int k = it - 16, pl = k >> 3, line = k & 7;
                        const int bs = (bS >> (8 * (line >> 1))) & 255;
                        if (bs && alpha_c && beta_c)
                        {
                            uint8_t *base = stage + DB_STAGE_C + pl * 12 * DB_CS + 4 * DB_CS + 4;
                            uint8_t *q = dir == 0 ? base + line * DB_CS + 2 * e : base + (2 * e) * DB_CS + line;
                            db_chroma_line(q, dir == 0 ? 1 : DB_CS, alpha_c, beta_c, db_tc0(dp.qp_chroma + dp.alpha_c0_offset, bs) + 1);
                        }
                    }
                }
            }
        }
    }
    team_sync();
    // (4) copy out
    PCAMV_FOR_ITEMS(it, 32)
    {
        if (it < 20)
        {
            uint32_t *dst = (uint32_t *)(py + (ptrdiff_t)(it - 4) * rp.stride_y - 4);
            const uint32_t *src = (const uint32_t *)(stage + DB_STAGE_Y + it * DB_LS);
#pragma unroll
            for (int k = 0; k < 5; k++) PCAMV_STV(dst + k, src[k]);
        }
        else
        {
            const int r = it - 20;
#pragma unroll
            for (int pl = 0; pl < 2; pl++)
            {
                uint32_t *dst = (uint32_t *)((pl ? pv : pu) + (ptrdiff_t)(r - 4) * rp.stride_c - 4);
                const uint32_t *src = (const uint32_t *)(stage + DB_STAGE_C + pl * 12 * DB_CS + r * DB_CS);
#pragma unroll
                for (int k = 0; k < 3; k++) PCAMV_STV(dst + k, src[k]);
            }
        }
    }
    team_sync();
}

} // namespace pcamv
