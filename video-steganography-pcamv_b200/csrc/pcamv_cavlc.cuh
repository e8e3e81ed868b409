// pcamv_cavlc.cuh — CAVLC size of an inter macroblock: the bit count x264_macroblock_size_cavlc (encoder/rdo.c:41-48 =
// encoder/cavlc.c:61-198,285-600 compiled with RDO_SKIP_BS, "produces exactly the same bit count as a normal encode") returns
// for a P_L0 / P_8x8 macroblock, as a host/device function.
//
// STATUS: the first parity-tested piece of RD mode decision (the second is csrc/pcamv_rd.cuh) (--subme 6 / 7), NOT on the product path: no kernel
// calls it yet and the bound host still refuses --subme >= 6.  Why this piece first: with --no-cabac the size of a macroblock
// depends on nothing but the macroblock itself, the motion-vector predictors and the coefficient counts of its left / top
// neighbours — all of which the wavefront already orders — whereas CABAC sizes depend on the live arithmetic-coder state
// (encoder/rdo.c:153-157), a raster-serial chain.  Checked on the CPU (tests/emu/emu_cavlc_check.cpp, tests/test_emu_cavlc.py)
// against every macroblock the reference's x264_rd_cost_mb sized at --subme 6 --no-cabac ('RDMB' records of
// oracle/_ref/x264_dump_rd).
//
// Tables: the code LENGTHS of coeff_token / total_zeros / run_before (H.264 tables 9-5, 9-7 .. 9-10) are not restated here; they
// arrive in a CavlcSizes record filled from the encoder's own tables (common/vlc.c) — dumped as 'VLC0' by the oracle, uploaded
// by the host in a product, the way the lambda*bits tables are (pcamv_set_qp_tables).  Level codes, Exp-Golomb codes and the
// coded_block_pattern mapping (table 9-4) are arithmetic / H.264's own constants and are written out.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
  #define PCAMV_CAVLC_HD __host__ __device__
#else
  #define PCAMV_CAVLC_HD
#endif

namespace pcamv {

struct CavlcSizes                     // bit lengths only; layout == the 'VLC0' record
{
    uint8_t coeff0[5];                // x264_coeff0_token[t]
    uint8_t coeff_token[5][64];       // x264_coeff_token[t][4 * (total - 1) + trailing_ones]
    uint8_t total_zeros[15][16];      // x264_total_zeros[total - 1][total_zeros]
    uint8_t total_zeros_dc[3][4];     // x264_total_zeros_dc[total - 1][total_zeros]
    uint8_t run_before[7][16];        // x264_run_before[min(zeros_left - 1, 6)][run]
};

struct CavlcMb                        // one inter macroblock as x264_macroblock_size_cavlc sees it
{
    int type, partition;              // P_L0 = 4 / P_8x8 = 5; D_16x8 = 14, D_8x16 = 15, D_16x16 = 16
    int sub[4];                       // D_L0_4x4 = 0, D_L0_8x4 = 1, D_L0_4x8 = 2, D_L0_8x8 = 3
    int n_ref;                        // h->mb.pic.i_fref[0]
    int psub8x8;                      // X264_ANALYSE_PSUB8x8: sub_mb_types are written as ue(v), else as four 1-bit codes
    int cbp_luma, cbp_chroma, qp_delta;
    int8_t ref[4];                    // per 8x8 block
    int n_mvd;
    int16_t mvd[16][2];               // in the order the macroblock layer writes them
    const int16_t (*coef)[16];        // h->dct.luma4x4[24][16]: 16 luma blocks (block_idx order), 8 chroma AC blocks (read from [1])
    const int16_t (*chroma_dc)[4];    // h->dct.chroma_dc[2][4]
    uint8_t coded[26];                // h->mb.cache.non_zero_count of the macroblock's own blocks after x264_macroblock_encode: 16 luma,
                                      // 8 chroma AC, 2 chroma DC.  0 = the block is written as "no coefficients" WHATEVER its array holds:
                                      // decimation (encoder/macroblock.c:700-755) clears the counts and the cbp bits, not the arrays
    // coefficient counts of the neighbouring 4x4 blocks, 0x80 = not available (h->mb.cache.non_zero_count semantics):
    uint8_t nnz_left[4], nnz_top[4];                 // luma: rows 0..3 of the left macroblock's last column, columns 0..3 of the top's last row
    uint8_t nnz_left_c[2][2], nnz_top_c[2][2];       // chroma AC, per plane
};

PCAMV_CAVLC_HD static inline int cavlc_ue_bits(unsigned v) { int n = 0; for (v += 1; v > 1; v >>= 1) n++; return 2 * n + 1; }
PCAMV_CAVLC_HD static inline int cavlc_se_bits(int v) { return cavlc_ue_bits(v <= 0 ? (unsigned)(-v) * 2u : (unsigned)v * 2u - 1u); }
PCAMV_CAVLC_HD static inline int cavlc_te_bits(int max, int v) { return max == 1 ? 1 : cavlc_ue_bits((unsigned)v); }

// x264_mb_predict_non_zero_code (common/macroblock.h:445-457) + ct_index (encoder/cavlc.c:114)
PCAMV_CAVLC_HD static inline int cavlc_table_of(int na, int nb)
{
    int n = na + nb;
    if (n < 0x80) n = (n + 1) >> 1;
    n &= 0x7f;
    return n < 2 ? 0 : n < 4 ? 1 : n < 8 ? 2 : 3;
}

// bits of one level code (H.264 9.2.2.1 read forwards) and the suffix length that follows it
PCAMV_CAVLC_HD static inline int cavlc_level_bits(int level, int &suffix_length, int first_after_fewer_than_3_trailing)
{
    const int a = level < 0 ? -level : level;
    int code = 2 * a - 2 + (level < 0);
    int bits;
    if (first_after_fewer_than_3_trailing) code -= 2;
    if (suffix_length == 0)
        bits = code < 14 ? code + 1 : code < 30 ? 19 : 28;
    else
        bits = (code >> suffix_length) < 15 ? (code >> suffix_length) + 1 + suffix_length : 28;
    if (suffix_length == 0) suffix_length = 1;
    if (a > (3 << (suffix_length - 1)) && suffix_length < 6) suffix_length++;
    return bits;
}

// block_residual_write_cavlc (encoder/cavlc.c:112-198): `l` in zigzag order, n = 16 (luma 4x4), 15 (chroma AC) or 4 (chroma DC);
// t = table (0..3 from the neighbours' counts, 4 = chroma DC).  Returns the bits, *total receives the coefficient count.
PCAMV_CAVLC_HD static inline int cavlc_block_bits(const CavlcSizes &z, const int16_t *l, int n, int t, int coded, int *total)
{
    int last = -1, cnt = 0;
    for (int i = 0; coded && i < n; i++)
        if (l[i]) { last = i; cnt++; }
    *total = cnt;
    if (!cnt) return z.coeff0[t];
    int trailing = 0;
    for (int i = last, k = 0; i >= 0 && k < 3; i--)           // up to three +-1 at the high-frequency end
    {
        if (!l[i]) continue;
        if (l[i] != 1 && l[i] != -1) break;
        trailing++; k++;
    }
    int bits = z.coeff_token[t][4 * (cnt - 1) + trailing] + trailing;
    int suffix_length = cnt > 10 && trailing < 3;
    int seen = 0, zeros_left = last + 1 - cnt, prev = -1;
    const int total_zeros = zeros_left;
    for (int i = last; i >= 0; i--)
    {
        if (!l[i]) continue;
        if (seen >= trailing)
            bits += cavlc_level_bits(l[i], suffix_length, seen == trailing && trailing < 3);
        // run_before of the previous (higher-frequency) coefficient = zeros between it and this one
        if (prev >= 0 && zeros_left > 0)
        {
            const int run = prev - i - 1;
            bits += z.run_before[zeros_left - 1 < 6 ? zeros_left - 1 : 6][run];
            zeros_left -= run;
        }
        prev = i;
        seen++;
    }
    if (cnt < n)
        bits += t == 4 ? z.total_zeros_dc[cnt - 1][total_zeros] : z.total_zeros[cnt - 1][total_zeros];
    return bits;
}

// coded_block_pattern -> codeNum, inter column of H.264 table 9-4 (cbp = luma | chroma << 4)
PCAMV_CAVLC_HD static inline int cavlc_inter_cbp_code(int cbp)
{
    const uint8_t by_code[48] = { 0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44,
                                  33, 34, 36, 40, 39, 43, 45, 46, 17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41 };
    for (int c = 0; c < 48; c++)
        if (by_code[c] == cbp) return c;
    return 0;
}

// x264_macroblock_write_cavlc for P_L0 / P_8x8 (encoder/cavlc.c:285-600) as a bit count
PCAMV_CAVLC_HD static inline int cavlc_mb_inter_bits(const CavlcSizes &z, const CavlcMb &m)
{
    int bits = 0;
    if (m.type == 4)                                          // P_L0: mb_type, ref_idx per partition, the mvds
    {
        const int parts = m.partition == 16 ? 1 : 2;
        bits += cavlc_ue_bits(m.partition == 16 ? 0 : m.partition == 14 ? 1 : 2);
        if (m.n_ref > 1)
            for (int i = 0; i < parts; i++)
                bits += cavlc_te_bits(m.n_ref - 1, m.ref[m.partition == 14 ? 2 * i : i]);
    }
    else                                                      // P_8x8 (mb_type 3) / P_8x8ref0 (4): the same length
    {
        const int ref0 = (m.ref[0] | m.ref[1] | m.ref[2] | m.ref[3]) == 0;
        bits += cavlc_ue_bits(ref0 ? 4 : 3);
        if (m.psub8x8)
            for (int i = 0; i < 4; i++) bits += cavlc_ue_bits(m.sub[i] == 3 ? 0 : m.sub[i] == 1 ? 1 : m.sub[i] == 2 ? 2 : 3);
        else
            bits += 4;
        if (!ref0 && m.n_ref > 1)
            for (int i = 0; i < 4; i++) bits += cavlc_te_bits(m.n_ref - 1, m.ref[i]);
    }
    for (int i = 0; i < m.n_mvd; i++) bits += cavlc_se_bits(m.mvd[i][0]) + cavlc_se_bits(m.mvd[i][1]);
    bits += cavlc_ue_bits(cavlc_inter_cbp_code(m.cbp_luma | (m.cbp_chroma << 4)));
    if (m.cbp_luma | m.cbp_chroma)
    {
        bits += cavlc_se_bits(m.qp_delta);
        // luma: the count of every block is the context of its right / lower neighbours; blocks of an uncoded 8x8 count 0
        uint8_t cnt[4][4];                                    // [row][column] inside the macroblock
        for (int i = 0; i < 16; i++)
        {
            const int x = (i & 1) | ((i >> 1) & 2), y = ((i >> 1) & 1) | ((i >> 2) & 2);
            int total = 0;
            if (m.cbp_luma & (1 << (i >> 2)))
            {
                const int na = x ? cnt[y][x - 1] : m.nnz_left[y], nb = y ? cnt[y - 1][x] : m.nnz_top[x];
                bits += cavlc_block_bits(z, m.coef[i], 16, cavlc_table_of(na, nb), m.coded[i], &total);
            }
            cnt[y][x] = (uint8_t)total;
        }
    }
    if (m.cbp_chroma)
    {
        int total;
        bits += cavlc_block_bits(z, m.chroma_dc[0], 4, 4, m.coded[24], &total);
        bits += cavlc_block_bits(z, m.chroma_dc[1], 4, 4, m.coded[25], &total);
        if (m.cbp_chroma & 2)
            for (int pl = 0; pl < 2; pl++)
            {
                uint8_t cnt[2][2];
                for (int i = 0; i < 4; i++)
                {
                    const int x = i & 1, y = i >> 1;
                    const int na = x ? cnt[y][0] : m.nnz_left_c[pl][y], nb = y ? cnt[0][x] : m.nnz_top_c[pl][x];
                    bits += cavlc_block_bits(z, m.coef[16 + 4 * pl + i] + 1, 15, cavlc_table_of(na, nb), m.coded[16 + 4 * pl + i], &total);
                    cnt[y][x] = (uint8_t)total;
                }
            }
    }
    return bits;
}

} // namespace pcamv
