// pcamv_rd.cuh — the distortion half of x264_rd_cost_mb (encoder/rdo.c:106-172): SSD of the reconstructed macroblock against the
// source, luma + both chroma planes, plus the psy-RD term of the luma plane ("absolute difference of complexities",
// rdo.c:97-131: Hadamard AC energy of the reconstruction against the cached AC energy of the source, encoder/analyse.c:522-549).
//
// STATUS: second parity-tested piece of RD mode decision (--subme 6 / 7), NOT on the product path.  It takes the
// macroblock exactly where the existing device code leaves it after reconstructing a candidate — source pixels in w.fenc_*,
// reconstruction in w.pred_* (recon_mb / encode_mb_inter) — so that together with csrc/pcamv_cavlc.cuh both halves of the RD cost
// exist; what is missing is the candidate driver and the wiring.  Written as straight scalar code (one lane): a kernel would spread
// the 4x4 transforms over the team like cand_cost does.  Checked on the CPU against every inter candidate the reference's RD
// mode decision costed (tests/emu/emu_rd_check.cpp, 'RDMB' records of oracle/_ref/x264_dump_rd).
//
// Arithmetic notes (all integer): the reference computes the Hadamard sums in packed 16-bit lanes (common/pixel.c:306-355);
// mathematically that is, per 8x8 block, sum4 = sum of |4x4 Hadamard coefficients| of its four 4x4 blocks minus the pixel sum,
// sum8 = sum of |8x8 Hadamard coefficients| minus the pixel sum; 16x16: (sum of sum4) >> 1 and (sum of sum8) >> 2.  The source side
// rounds per block: satd_4x4 = sum|H4| >> 1 minus (pixel sum >> 1); sa8d_8x8 = (sum|H8| + 2) >> 2 minus (pixel sum >> 2).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
  #define PCAMV_RD_HD __host__ __device__
#else
  #define PCAMV_RD_HD
#endif

namespace pcamv {

PCAMV_RD_HD static inline int rd_iabs(int v) { return v < 0 ? -v : v; }

// unnormalised 4x4 Hadamard transform of a 4x4 pixel block (rows `stride` apart) into t[16]
PCAMV_RD_HD static inline void rd_hadamard4x4(const uint8_t *p, int stride, int t[16])
{
    int r[16];
    for (int y = 0; y < 4; y++)
    {
        const int a = p[y * stride], b = p[y * stride + 1], c = p[y * stride + 2], d = p[y * stride + 3];
        const int s0 = a + b, d0 = a - b, s1 = c + d, d1 = c - d;
        r[4 * y] = s0 + s1; r[4 * y + 1] = d0 + d1; r[4 * y + 2] = s0 - s1; r[4 * y + 3] = d0 - d1;
    }
    for (int x = 0; x < 4; x++)
    {
        const int a = r[x], b = r[4 + x], c = r[8 + x], d = r[12 + x];
        const int s0 = a + b, d0 = a - b, s1 = c + d, d1 = c - d;
        t[x] = s0 + s1; t[4 + x] = d0 + d1; t[8 + x] = s0 - s1; t[12 + x] = d0 - d1;
    }
}

// one 8x8 block: *sum4 = sum over its four 4x4 blocks of sum|H4|, *sum8 = sum|H8| (the 8x8 transform = a 2x2 Hadamard across the
// four 4x4 transforms, coefficient by coefficient), *dc = pixel sum; *sum4_rounded = the source-side per-block rounding
PCAMV_RD_HD static inline void rd_hadamard_8x8(const uint8_t *p, int stride, int *sum4, int *sum8, int *dc, int *satd4_rounded)
{
    int t[4][16];
    int s4 = 0, s8 = 0, pix = 0, rounded = 0;
    for (int k = 0; k < 4; k++)
    {
        rd_hadamard4x4(p + 4 * (k & 1) + 4 * (k >> 1) * stride, stride, t[k]);
        int a = 0;
        for (int i = 0; i < 16; i++) a += rd_iabs(t[k][i]);
        s4 += a; pix += t[k][0];
        rounded += (a >> 1) - (t[k][0] >> 1);                 // satd_4x4( zero, blk ) - ( sad_4x4( zero, blk ) >> 1 )
    }
    for (int i = 0; i < 16; i++)
    {
        const int a = t[0][i], b = t[1][i], c = t[2][i], d = t[3][i];
        s8 += rd_iabs(a + b + c + d) + rd_iabs(a - b + c - d) + rd_iabs(a + b - c - d) + rd_iabs(a - b - c + d);
    }
    *sum4 = s4; *sum8 = s8; *dc = pix; *satd4_rounded = rounded;
}

// ssd_mb (rdo.c:133-138) with the psy term of ssd_plane (rdo.c:106-131) for the 16x16 luma plane.
// fenc_* / rec_*: the macroblock's source and reconstruction, luma 16x16 with row pitch 16, chroma 8x8 with row pitch 8 (MbWork's
// fenc_* and pred_* arrays).  psy_rd = h->mb.i_psy_rd (FIX8 of --psy-rd's first number, 256 by default), lambda = x264_lambda_tab[qp].
PCAMV_RD_HD static inline int rd_distortion_mb(const uint8_t *fenc_y, const uint8_t *rec_y, const uint8_t *fenc_u, const uint8_t *rec_u,
                                              const uint8_t *fenc_v, const uint8_t *rec_v, int psy_rd, int lambda)
{
    int ssd = 0;
    for (int i = 0; i < 256; i++) { const int d = fenc_y[i] - rec_y[i]; ssd += d * d; }
    for (int i = 0; i < 64; i++) { const int d = fenc_u[i] - rec_u[i]; ssd += d * d; }
    for (int i = 0; i < 64; i++) { const int d = fenc_v[i] - rec_v[i]; ssd += d * d; }
    if (!psy_rd) return ssd;
    int src_satd = 0, src_sa8d = 0, rec_sum4 = 0, rec_sum8 = 0;
    for (int k = 0; k < 4; k++)
    {
        const int off = 8 * (k & 1) + 128 * (k >> 1);
        int s4, s8, dc, rounded;
        rd_hadamard_8x8(fenc_y + off, 16, &s4, &s8, &dc, &rounded);      // x264_mb_cache_fenc_satd (analyse.c:522-549)
        src_satd += rounded;
        src_sa8d += ((s8 + 2) >> 2) - (dc >> 2);
        rd_hadamard_8x8(rec_y + off, 16, &s4, &s8, &dc, &rounded);       // x264_pixel_hadamard_ac_16x16 (pixel.c:306-355)
        rec_sum4 += s4 - dc; rec_sum8 += s8 - dc;
    }
    int psy = (rd_iabs((rec_sum4 >> 1) - src_satd) + rd_iabs((rec_sum8 >> 2) - src_sa8d)) >> 1;
    psy = (psy * psy_rd * lambda + 128) >> 8;
    return ssd + psy;
}

// ---- the quantised levels of a candidate, from the device's own residual path ---------------------------------------------------
// Only when the frame-level device code is in the translation unit (pcamv_frame.cuh / pcamv_cost.cuh included first).
#if defined(PCAMV_FN)
struct RdLevels                      // what x264_macroblock_encode leaves in h->dct / h->mb for the entropy coder, inter macroblock
{
    int16_t coef[24][16];            // h->dct.luma4x4: zigzag-scanned levels, 16 luma blocks (block_idx order) + 8 chroma AC blocks ([0] = 0)
    int16_t chroma_dc[2][4];         // h->dct.chroma_dc
    uint8_t coded[26];               // non_zero_count != 0 of the 16 + 8 + 2 blocks (after decimation)
    int cbp_luma, cbp_chroma;
};

// Call with the motion-compensated prediction staged in c.w.pred_* (before encode_mb_residual adds the residual to it).  Same
// arithmetic and the same decisions as encode_mb_residual (csrc/pcamv_cost.cuh = encoder/macroblock.c:690-755 luma, :277-372
// chroma), but it keeps what that function only passes through: the levels.  One lane (see the header comment).
PCAMV_DEV void rd_levels_mb(MbCtx &c, RdLevels &o)
{
    const DevTables &t = c.fc.tab;
    const unsigned long long zz = 0xfbeda7369c852140ull;      // zigzag position i lives at dct[x][y], flattened 4*x+y (decimate_score)
    int score[24], nz[24];
    int16_t dcs[8];
    for (int it = 0; it < 24; it++)
    {
        const int ch = it >= 16;
        int d[16], co[16];
        if (!ch)
        {
            const int bx = (it & 1) | ((it >> 1) & 2), by = ((it >> 1) & 1) | ((it >> 2) & 2);
            load_residual(c.w.fenc_y + 64 * by + 4 * bx, 16, c.w.pred_y + 64 * by + 4 * bx, 16, d);
        }
        else
        {
            const int pl = (it - 16) >> 2, blk = (it - 16) & 3;
            const uint8_t *fe = pl ? c.w.fenc_v : c.w.fenc_u, *pr = pl ? c.w.pred_v : c.w.pred_u;
            load_residual(fe + 32 * (blk >> 1) + 4 * (blk & 1), 8, pr + 32 * (blk >> 1) + 4 * (blk & 1), 8, d);
        }
        dct4x4<1>(d, co);
        if (ch) { dcs[it - 16] = (int16_t)co[0]; co[0] = 0; }
        nz[it] = quant4x4<1>(co, t.quant4_mf[ch], t.quant4_bias[ch]);
        int16_t q[16];
        for (int i = 0; i < 16; i++) q[i] = (int16_t)co[i];
        score[it] = nz[it] ? decimate_score(q, ch) : 0;
        for (int i = 0; i < 16; i++) o.coef[it][i] = q[(zz >> (4 * i)) & 15];
    }
    const int b_decimate = c.fc.b_dct_decimate;
    int s8[4], any8[4], mb = 0;
    for (int i = 0; i < 4; i++)
    {
        s8[i] = 0; any8[i] = 0;
        for (int j = 0; j < 4; j++) { any8[i] |= nz[4 * i + j]; s8[i] += score[4 * i + j]; }
        mb += s8[i];
    }
    o.cbp_luma = 0;
    for (int i = 0; i < 4; i++)
        if (b_decimate ? (mb >= 6 && s8[i] >= 4) : any8[i]) o.cbp_luma |= 1 << i;
    for (int it = 0; it < 16; it++) o.coded[it] = (uint8_t)(((o.cbp_luma >> (it >> 2)) & 1) && nz[it]);
    int any_ac = 0, any_dc = 0;
    for (int pl = 0; pl < 2; pl++)
    {
        int nz_ac = 0, sc = 0;
        for (int j = 0; j < 4; j++) { nz_ac |= nz[16 + 4 * pl + j]; sc += score[16 + 4 * pl + j]; }
        const int b0 = dcs[4 * pl], b1 = dcs[4 * pl + 1], b2 = dcs[4 * pl + 2], b3 = dcs[4 * pl + 3];
        const int D0 = b0 + b1, D1 = b2 + b3, D2 = b0 - b1, D3 = b2 - b3;
        const int mf = t.quant4_mf[1][0] >> 1, bias = t.quant4_bias[1][0] << 1;
        // dct2x2dc + quant_2x2_dc + zigzag_scan_2x2_dc (common/dct.c:87-105, quant.c:63-75); the reference's 2x2 array is indexed
        // [x][y] of the 4x4 block inside the plane, so its scan order is: sum, horizontal difference, vertical difference, diagonal
        o.chroma_dc[pl][0] = (int16_t)quant_one((int16_t)(D0 + D1), mf, bias);
        o.chroma_dc[pl][1] = (int16_t)quant_one((int16_t)(D2 + D3), mf, bias);
        o.chroma_dc[pl][2] = (int16_t)quant_one((int16_t)(D0 - D1), mf, bias);
        o.chroma_dc[pl][3] = (int16_t)quant_one((int16_t)(D2 - D3), mf, bias);
        const int nz_dc = (o.chroma_dc[pl][0] | o.chroma_dc[pl][1] | o.chroma_dc[pl][2] | o.chroma_dc[pl][3]) != 0;
        const int keep_ac = !((b_decimate && sc < 7) || !nz_ac);
        for (int j = 0; j < 4; j++) o.coded[16 + 4 * pl + j] = (uint8_t)(keep_ac && nz[16 + 4 * pl + j]);
        o.coded[24 + pl] = (uint8_t)nz_dc;
        any_ac |= keep_ac; any_dc |= nz_dc;
    }
    o.cbp_chroma = any_ac ? 2 : any_dc ? 1 : 0;
}

// the motion compensation of a candidate into c.w.pred_* (the first half of recon_mb, csrc/pcamv_recon.cuh, verbatim)
PCAMV_FN void rd_mc_inter(MbCtx &c, const MbResult &r)
{
    init_limits(c);
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
#endif

} // namespace pcamv
