// pcamv_intra.cuh — intra mode analysis of a macroblock, first pieces: the 16x16 luma modes and the 8x8 chroma
// modes as x264_mb_analyse_intra / x264_mb_analyse_intra_chroma cost them (encoder/analyse.c:628-684, 552-625; predictors
// common/predict.c:137-330 16x16, :335-520 8x8c; mode availability analyse.c:387-450).
//
// STATUS: NOT on the product path.  In the product the host runs the reference's intra analysis (for I frames, and in P frames for
// the macroblocks of quirk q1 only — intra is disabled in P slices).  What these pieces are for: the intra SATD
// costs are the threshold of RD mode decision (analyse.c:2826-2832) and the statistic / fdec leftovers of a P macroblock.  Covered:
// the 16x16 modes, the chroma modes and the 4x4 modes (nine predictors per block with the intra encode between blocks, analyse.c:770-879,
// on the product's transform / quantisation primitives).  Missing: the I-frame decision and encode, the 8x8 modes (8x8dct is not served), wiring.
// Straight one-lane code on the macroblock's source pixels and the reconstructed border of its neighbours (top-left, top row,
// left column), which is what a wavefront kernel has for the macroblock it works on once neighbours are reconstructed in place.
// Checked on the CPU against every call of the reference's analysis, I and P slices ('INTR' records of oracle/_ref/x264_dump_rd,
// tests/emu/emu_intra_check.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
  #define PCAMV_INTRA_HD __host__ __device__
#else
  #define PCAMV_INTRA_HD
#endif

namespace pcamv {

enum { I16_V = 0, I16_H = 1, I16_DC = 2, I16_P = 3, I16_DC_LEFT = 4, I16_DC_TOP = 5, I16_DC_128 = 6 };        // common/predict.h:49-58
enum { IC_DC = 0, IC_H = 1, IC_V = 2, IC_P = 3, IC_DC_LEFT = 4, IC_DC_TOP = 5, IC_DC_128 = 6 };               // common/predict.h:32-41

PCAMV_INTRA_HD static inline int intra_clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
PCAMV_INTRA_HD static inline int intra_ue_bits(unsigned v) { int n = 0; for (v += 1; v > 1; v >>= 1) n++; return 2 * n + 1; }

// SATD of an N x N block (N = 16 or 8) against its prediction: sum over 4x4 blocks of sum|H4(diff)| >> 1 — as x264_pixel_satd_WxH
// composes it from 8x4 / 4x4 pieces (common/pixel.c:187-253): every 4x4 contributes its unnormalised Hadamard sum, the halving is
// applied per 8x4 pair; 16x16 and 8x8 tile into whole pairs, so sum over pairs of (s0 + s1) >> 1.
PCAMV_INTRA_HD static inline int intra_satd(const uint8_t *src, int ss, const uint8_t *pred, int ps, int n)
{
    int total = 0;
    for (int by = 0; by < n; by += 4)
        for (int bx = 0; bx < n; bx += 8)
        {
            int pair = 0;
            for (int k = 0; k < 2; k++)
            {
                int d[16], r[16];
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++) d[4 * y + x] = src[(by + y) * ss + bx + 4 * k + x] - pred[(by + y) * ps + bx + 4 * k + x];
                for (int y = 0; y < 4; y++)
                {
                    const int a = d[4 * y], b = d[4 * y + 1], c = d[4 * y + 2], e = d[4 * y + 3];
                    r[4 * y] = a + b + c + e; r[4 * y + 1] = a - b + c - e; r[4 * y + 2] = a + b - c - e; r[4 * y + 3] = a - b - c + e;
                }
                for (int x = 0; x < 4; x++)
                {
                    const int a = r[x], b = r[4 + x], c = r[8 + x], e = r[12 + x];
                    const int t0 = a + b + c + e, t1 = a - b + c - e, t2 = a + b - c - e, t3 = a - b - c + e;
                    pair += (t0 < 0 ? -t0 : t0) + (t1 < 0 ? -t1 : t1) + (t2 < 0 ? -t2 : t2) + (t3 < 0 ? -t3 : t3);
                }
            }
            total += pair >> 1;
        }
    return total;
}

// 16x16 luma predictors (common/predict.c:137-330).  tl = top-left pixel, top[16], left[16]; out[256], row pitch 16.
PCAMV_INTRA_HD static inline void intra_predict_16x16(int mode, int tl, const uint8_t *top, const uint8_t *left, uint8_t *out)
{
    if (mode == I16_V) { for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) out[16 * y + x] = top[x]; return; }
    if (mode == I16_H) { for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) out[16 * y + x] = left[y]; return; }
    if (mode == I16_P)
    {
        int H = 0, V = 0;
        for (int i = 1; i <= 8; i++)
        {
            H += i * (top[7 + i] - (i == 8 ? tl : top[7 - i]));
            V += i * (left[7 + i] - (i == 8 ? tl : left[7 - i]));
        }
        const int a = 16 * (left[15] + top[15]), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++) out[16 * y + x] = (uint8_t)intra_clip8((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
        return;
    }
    int dc = 128;
    if (mode == I16_DC || mode == I16_DC_LEFT || mode == I16_DC_TOP)
    {
        int s = 0;
        if (mode != I16_DC_TOP) for (int i = 0; i < 16; i++) s += left[i];
        if (mode != I16_DC_LEFT) for (int i = 0; i < 16; i++) s += top[i];
        dc = mode == I16_DC ? (s + 16) >> 5 : (s + 8) >> 4;
    }
    for (int i = 0; i < 256; i++) out[i] = (uint8_t)dc;
}

// 8x8 chroma predictors (common/predict.c:335-520).  out[64], row pitch 8.
PCAMV_INTRA_HD static inline void intra_predict_8x8c(int mode, int tl, const uint8_t *top, const uint8_t *left, uint8_t *out)
{
    if (mode == IC_V) { for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) out[8 * y + x] = top[x]; return; }
    if (mode == IC_H) { for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) out[8 * y + x] = left[y]; return; }
    if (mode == IC_P)
    {
        int H = 0, V = 0;
        for (int i = 1; i <= 4; i++)
        {
            H += i * (top[3 + i] - (i == 4 ? tl : top[3 - i]));
            V += i * (left[3 + i] - (i == 4 ? tl : left[3 - i]));
        }
        const int a = 16 * (left[7] + top[7]), b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) out[8 * y + x] = (uint8_t)intra_clip8((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
        return;
    }
    // the DC family works per 4x4 quadrant: s0 / s1 = left / right half of the top row, s2 / s3 = upper / lower half of the left column
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0, dc[4] = { 128, 128, 128, 128 };
    for (int i = 0; i < 4; i++) { s0 += top[i]; s1 += top[4 + i]; s2 += left[i]; s3 += left[4 + i]; }
    if (mode == IC_DC) { dc[0] = (s0 + s2 + 4) >> 3; dc[1] = (s1 + 2) >> 2; dc[2] = (s3 + 2) >> 2; dc[3] = (s1 + s3 + 4) >> 3; }
    else if (mode == IC_DC_LEFT) { dc[0] = dc[1] = (s2 + 2) >> 2; dc[2] = dc[3] = (s3 + 2) >> 2; }
    else if (mode == IC_DC_TOP) { dc[0] = dc[2] = (s0 + 2) >> 2; dc[1] = dc[3] = (s1 + 2) >> 2; }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) out[8 * y + x] = (uint8_t)dc[2 * (y >> 2) + (x >> 2)];
}

struct IntraCosts { int satd16, pred16, dir16[7], satd_c, pred_c; };

// modes in the reference's evaluation order; a strict `<` keeps the first minimum (COPY2_IF_LT)
PCAMV_INTRA_HD static inline void intra_analyse_16x16(const uint8_t *fenc /*pitch 16*/, int has_left, int has_top, int has_topleft, int tl,
                                                      const uint8_t *top, const uint8_t *left, int lambda, IntraCosts &o)
{
    int modes[4], n;
    if (has_topleft) { modes[0] = I16_V; modes[1] = I16_H; modes[2] = I16_DC; modes[3] = I16_P; n = 4; }
    else if (has_left) { modes[0] = I16_DC_LEFT; modes[1] = I16_H; n = 2; }
    else if (has_top) { modes[0] = I16_DC_TOP; modes[1] = I16_V; n = 2; }
    else { modes[0] = I16_DC_128; n = 1; }
    uint8_t pred[256];
    o.satd16 = 1 << 28; o.pred16 = 0;
    for (int i = 0; i < 7; i++) o.dir16[i] = 0;
    for (int i = 0; i < n; i++)
    {
        const int m = modes[i], fix = m > I16_P ? I16_DC : m;
        intra_predict_16x16(m, tl, top, left, pred);
        const int cost = intra_satd(pred, 16, fenc, 16, 16) + lambda * intra_ue_bits(fix);
        o.dir16[m] = cost;
        if (cost < o.satd16) { o.satd16 = cost; o.pred16 = m; }
    }
}
PCAMV_INTRA_HD static inline void intra_analyse_chroma(const uint8_t *fenc_u, const uint8_t *fenc_v /*pitch 8*/, int has_left, int has_top, int has_topleft,
                                                       const int tl[2], const uint8_t *top_u, const uint8_t *left_u, const uint8_t *top_v, const uint8_t *left_v,
                                                       int lambda, IntraCosts &o)
{
    int modes[4], n;
    if (has_topleft) { modes[0] = IC_V; modes[1] = IC_H; modes[2] = IC_DC; modes[3] = IC_P; n = 4; }
    else if (has_left) { modes[0] = IC_DC_LEFT; modes[1] = IC_H; n = 2; }
    else if (has_top) { modes[0] = IC_DC_TOP; modes[1] = IC_V; n = 2; }
    else { modes[0] = IC_DC_128; n = 1; }
    uint8_t pu[64], pv[64];
    o.satd_c = 1 << 28; o.pred_c = 0;
    for (int i = 0; i < n; i++)
    {
        const int m = modes[i], fix = m > IC_P ? IC_DC : m;
        intra_predict_8x8c(m, tl[0], top_u, left_u, pu);
        intra_predict_8x8c(m, tl[1], top_v, left_v, pv);
        const int cost = intra_satd(pu, 8, fenc_u, 8, 8) + intra_satd(pv, 8, fenc_v, 8, 8) + lambda * intra_ue_bits(fix);
        if (cost < o.satd_c) { o.satd_c = cost; o.pred_c = m; }
    }
}

// ---- the 4x4 luma modes (analyse.c:770-879) -----------------------------------------------------------------------------------------
// Only when the frame-level device code is in the translation unit (pcamv_frame.cuh first): the intra encode between blocks uses
// the product's own transform / quantisation primitives.
#if defined(PCAMV_FN)
enum { I4_V = 0, I4_H = 1, I4_DC = 2, I4_DDL = 3, I4_DDR = 4, I4_VR = 5, I4_HD = 6, I4_VL = 7, I4_HU = 8, I4_DC_LEFT = 9, I4_DC_TOP = 10, I4_DC_128 = 11 };
enum { INB_LEFT = 1, INB_TOP = 2, INB_TOPRIGHT = 4, INB_TOPLEFT = 8 };                     // common/macroblock.h:28-34

// the nine predictors + the DC variants (common/predict.c:525-720 = H.264 8.3.1.2).  p = the block's top-left pixel inside a
// picture-like buffer with row pitch `s`: row -1 holds top-left, top[0..3] and top-right[4..7], column -1 the left pixels.
PCAMV_DEV void intra_predict_4x4(int mode, uint8_t *p, int s)
{
    int t[9], l[5];                                  // t[0] = top-left, t[1..8] = top / top-right; l[0] = top-left, l[1..4] = left
    for (int i = 0; i < 9; i++) t[i] = p[-s - 1 + i];
    l[0] = t[0];
    for (int i = 0; i < 4; i++) l[1 + i] = p[i * s - 1];
#define T(i) t[(i) + 1]
#define L(i) l[(i) + 1]
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
        {
            int v;
            switch (mode)
            {
            case I4_V: v = T(x); break;
            case I4_H: v = L(y); break;
            case I4_DC: v = (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3; break;
            case I4_DC_LEFT: v = (L(0) + L(1) + L(2) + L(3) + 2) >> 2; break;
            case I4_DC_TOP: v = (T(0) + T(1) + T(2) + T(3) + 2) >> 2; break;
            case I4_DDL: v = (x == 3 && y == 3) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2; break;
            case I4_DDR:
                if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
                else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
                else v = (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
                break;
            case I4_VR:
            {
                const int z = 2 * x - y;
                if (z >= 0 && !(z & 1)) v = (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
                else if (z >= 0) v = (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
                else v = (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
                break;
            }
            case I4_HD:
            {
                const int z = 2 * y - x;
                if (z >= 0 && !(z & 1)) v = (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
                else if (z >= 0) v = (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
                else v = (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
                break;
            }
            case I4_VL:
                v = !(y & 1) ? (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1 : (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
                break;
            case I4_HU:
            {
                const int z = x + 2 * y;
                if (z > 5) v = L(3);
                else if (z == 5) v = (L(2) + 3 * L(3) + 2) >> 2;
                else if (!(z & 1)) v = (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
                else v = (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2;
                break;
            }
            default: v = 128;
            }
            p[y * s + x] = (uint8_t)v;
        }
#undef T
#undef L
}

struct Intra4x4In
{
    const uint8_t *fenc;            // source macroblock, pitch 16, 4-byte aligned
    uint8_t *buf;                   // work picture, pitch 32, 16-byte aligned: the macroblock's pixel (0,0) at buf[32 + 4]; row 0 = top-left
                                    // (column 3), top row (4..19), top-right (20..23); column 3 = left pixels.  Reconstructed in place.
    uint8_t nb4[16];                // h->mb.i_neighbour4: which neighbours every 4x4 block has (INB_*)
    int8_t left_mode[4], top_mode[4];   // cached prediction modes of the blocks left of / above the macroblock (-1 = none)
    int lambda, qp, mbrd, fast_intra, satd_inter, satd16, satd8x8;
    const uint16_t *quant_mf, *quant_bias;      // h->quant4_mf / quant4_bias [CQM_4IY][qp]
    const int32_t *dequant_mf;                  // h->dequant4_mf[CQM_4IY]
};

// returns i_satd_i4x4 (1 << 28 when the analysis gives up before the last block), modes of the analysed blocks in pred[]
PCAMV_DEV int intra_analyse_4x4(const Intra4x4In &in, int pred[16])
{
    const int COSTMAX = 1 << 28;
    int thresh = in.satd_inter < in.satd16 ? in.satd_inter : in.satd16;
    if (in.satd8x8 < thresh) thresh = in.satd8x8;
    if (in.mbrd) thresh = thresh * (10 - in.fast_intra) / 8;
    int cost = in.lambda * 24;
    int8_t modes[16];                                               // chosen modes, block_idx order
    for (int idx = 0;; idx++)
    {
        const int bx = (idx & 1) | ((idx >> 1) & 2), by = ((idx >> 1) & 1) | ((idx >> 2) & 2);
        uint8_t *p = in.buf + (4 * by + 1) * 32 + 4 + 4 * bx;
        // x264_mb_predict_intra4x4_mode: the smaller of the left / top block's mode (DC family folded to DC), DC when either is missing
        const int idx_l = bx ? idx - ((bx & 1) ? 1 : 3) : -1, idx_t = by ? idx - ((by & 1) ? 2 : 6) : -1;
        int ma = bx ? modes[idx_l] : in.left_mode[by], mb = by ? modes[idx_t] : in.top_mode[bx];
        if (ma > I4_HU) ma = I4_DC;
        if (mb > I4_HU) mb = I4_DC;
        const int pm = (ma < mb ? ma : mb) < 0 ? I4_DC : (ma < mb ? ma : mb);
        int list[9], n;
        const int nb = in.nb4[idx];
        if ((nb & INB_LEFT) && (nb & INB_TOP))
        {
            n = 0; list[n++] = I4_DC; list[n++] = I4_H; list[n++] = I4_V; list[n++] = I4_DDL;
            if (nb & INB_TOPLEFT) { list[n++] = I4_DDR; list[n++] = I4_VR; list[n++] = I4_HD; }
            list[n++] = I4_VL; list[n++] = I4_HU;
        }
        else if (nb & INB_LEFT) { list[0] = I4_DC_LEFT; list[1] = I4_H; list[2] = I4_HU; n = 3; }
        else if (nb & INB_TOP) { list[0] = I4_DC_TOP; list[1] = I4_V; list[2] = I4_DDL; list[3] = I4_VL; n = 4; }
        else { list[0] = I4_DC_128; n = 1; }
        if ((nb & (INB_TOPRIGHT | INB_TOP)) == INB_TOP)              // emulate the missing top-right samples
            for (int i = 4; i < 8; i++) p[-32 + i] = p[-32 + 3];
        int best = COSTMAX, best_mode = list[0];
        for (int i = 0; i < n; i++)
        {
            const int m = list[i], fix = m > I4_HU ? I4_DC : m;
            intra_predict_4x4(m, p, 32);
            int d[16], r[16], sum = 0;
            load_residual(in.fenc + 64 * by + 4 * bx, 16, p, 32, d);
            for (int y = 0; y < 4; y++)
            {
                const int a = d[4 * y], b = d[4 * y + 1], c = d[4 * y + 2], e = d[4 * y + 3];
                r[4 * y] = a + b + c + e; r[4 * y + 1] = a - b + c - e; r[4 * y + 2] = a + b - c - e; r[4 * y + 3] = a - b - c + e;
            }
            for (int x = 0; x < 4; x++)
            {
                const int a = r[x], b = r[4 + x], c = r[8 + x], e = r[12 + x];
                const int t0 = a + b + c + e, t1 = a - b + c - e, t2 = a + b - c - e, t3 = a - b - c + e;
                sum += (t0 < 0 ? -t0 : t0) + (t1 < 0 ? -t1 : t1) + (t2 < 0 ? -t2 : t2) + (t3 < 0 ? -t3 : t3);
            }
            const int c4 = (sum >> 1) + in.lambda * (pm == fix ? 1 : 4);
            if (c4 < best) { best = c4; best_mode = m; }
        }
        pred[idx] = best_mode;
        cost += best;
        if (cost > thresh || idx == 15)
            return idx == 15 ? cost : COSTMAX;
        // the chosen mode's prediction, then the intra encode of the block (x264_mb_encode_i4x4, encoder/macroblock.c:116-150)
        intra_predict_4x4(best_mode, p, 32);
        {
            int d[16], co[16], r[16];
            load_residual(in.fenc + 64 * by + 4 * bx, 16, p, 32, d);
            dct4x4<1>(d, co);
            if (quant4x4<1>(co, in.quant_mf, in.quant_bias))
            {
                dequant4x4<1>(co, in.dequant_mf, in.qp);
                idct4x4(co, r);
                add_residual(p, 32, r);
            }
        }
        modes[idx] = (int8_t)best_mode;
    }
}
#endif

} // namespace pcamv
