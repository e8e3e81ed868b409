// pcamv_frame.cuh — per-macroblock P-slice analysis: neighbour cache, MV prediction, P_SKIP probe,
// partition search order + mode decision, and the PCAMV candidate-MV cost table.
//
// Behavioural contract (bit-exact), reference file:line —
//   neighbour load            common/macroblock.c:914-1224 (x264_macroblock_cache_load, P part)
//   MV prediction             common/macroblock.c:28-163 (x264_mb_predict_mv, _16x16, _pskip)
//   extra 16x16 candidates    common/macroblock.c:388-470 (x264_mb_predict_mv_ref16x16)
//   MV limits                 encoder/analyse.c:271-318
//   P_SKIP probe              encoder/macroblock.c:809-895 (x264_macroblock_probe_skip)
//   search order / decision   encoder/analyse.c:1122-1205 (p16x16), :1371-1426 (p8x8), :1428-1533 (p16x8/p8x16),
//                             :2613-2810 (x264_macroblock_analyse, non-RD P path), :2658-2676 (pass-2 overrides)
//   MB reconstruction         encoder/macroblock.c:605-755 (inter luma), :277-372 (chroma), common/dct.c:122-262,364,
//                             common/quant.c:33-75,203-252, common/macroblock.c:483-508,630-650 (x264_mb_mc)
//   cost table                encoder/analyse.c:2364-2550 (MV_SATD_FDEC_IH, x264_ih_get_mv_cost), :3518-3689
// Every x264_me_search_ref / x264_me_refine_qpel / x264_ih_get_mv_cost result is appended to a per-MB
// log in call order; the host encoder replays that log instead of searching.
#pragma once
#include "pcamv_me.cuh"
#include "pcamv_device.h"
#include "pcamv_frame_types.h"

namespace pcamv {

// Per-macroblock state lives in SHARED memory, one copy per lane team: the control flow of the analysis is executed
// redundantly by every lane (uniform values), and keeping that state in per-thread local memory costs 32 copies of it
// in L1 per resident team (2 KB x 32 lanes) — with tens of encoder contexts resident the stacks alone evicted the
// reference windows from L1.
struct MeSlot            // the fields of x264_me_t that survive a search (per partition)
{
    MeResult r;
    int mvp[2];
    int i_ref, i_ref_cost, i_pixel, xoff, yoff;
};

struct MbAnalysis
{
    MeSlot me16x16, me8x8[4], me16x8[2], me8x16[2];
    int mvc[PCAMV_MAX_REFS][5][2];        // [ref][0] = 16x16 result, [ref][1..4] = 8x8 results
    int cost8x8, cost16x8, cost8x16;
    int8_t sub[4];                        // h->mb.i_sub_partition
};

// One search handed to a search team by the split wavefront (pcamv_split.cu): everything x264_me_search_ref /
// x264_me_refine_qpel read for one call (the fields of the reference's x264_me_t that are inputs, encoder/me.h:30-51)
struct SearchReq
{
    int32_t item;                // frame of the launch
    int16_t mb_x, mb_y;
    int8_t kind;                 // LOG_SEARCH / LOG_REFINE
    int8_t i_ref, i_pixel, i_mvc;
    int8_t xoff, yoff, has_thresh, pad;
    uint32_t mvp;                // packed x | y << 16
    int32_t thresh;              // *p_halfpel_thresh on entry
    uint32_t mv;                 // refine: the slot's vector, cost, cost_mv and reference cost on entry
    int32_t cost, cost_mv, i_ref_cost;
    int16_t lim[8];              // mv_min_fpel[2], mv_max_fpel[2], mv_min_spel[2], mv_max_spel[2]
    uint32_t mvc[PCAMV_MAX_MVC];
    int32_t pad2;                // split wavefront: sequence number of the row slot's request
};
static_assert(sizeof(SearchReq) <= 128 && sizeof(SearchReq) % 4 == 0, "a request is copied as up to 32 words");
struct alignas(16) SearchRes { uint32_t mv; int32_t cost, cost_mv, thresh; };
static_assert(sizeof(SearchRes) == 16, "a result is one 16-byte store");

// Where the analysis of a macroblock stands when it hands a search to somebody else and is resumed with the result (split
// wavefront).  The synchronous kernels run the same code and never leave it in the middle.
enum { PT_DONE = 0, PT_YIELD = 1 };
struct PtState
{
    int8_t stage;                // analyse_p_mb: 0 = not started
    int8_t in16, in8, in168;     // inside analyse_p16x16 / analyse_p8x8 / analyse_p16x8_8x16
    int8_t wait;                 // a search has been issued and its result not yet taken
    int8_t i_ref, i, j;          // loop positions of the function in progress
    int8_t type, partition, early_skip, b_try_pskip;
    int32_t i_cost, total;
};

// team-shared state of one macroblock.  Everything up to `fenc_y` is what has to survive while a search is out (the split
// wavefront parks exactly those bytes); the rest is scratch that is rebuilt or dead at those points.
struct alignas(16) MbWork
{
    int8_t ref[48];          // scan8-indexed neighbour cache, list 0
    uint32_t mv[48];
    MbAnalysis an;           // search results of the partitions tried so far
    MeSlot slot;             // the search in flight
    int mvc[PCAMV_MAX_MVC][2];
    int halfpel_thresh;
    PtState pt;
    SearchReq rq;            // the search handed out
    SearchRes rs;            // ... and its result
    alignas(16) uint8_t fenc_y[256];
    uint8_t fenc_u[64], fenc_v[64];
    uint8_t pred_y[256], pred_u[64], pred_v[64];     // MC / reconstruction staging
    int32_t scratch[32];
    int16_t coef[24][16];    // per 4x4 block (16 luma in block_idx order, 4 U, 4 V): quantised / dequantised coefficients
    MeBlock blk;             // block descriptor of the search in flight (synchronous kernels)
};

// Search results of the sub-8x8 partitions (reference x264_mb_analysis_t.l0.me4x4 / me8x4 / me4x8, encoder/analyse.c:66-75),
// 8 per 8x8 block: [0..3] 4x4, [4..5] 8x4, [6..7] 4x8.  They live in MbWork::coef (512 of its 768 bytes): the coefficient
// scratch is only in use inside the P_SKIP probe and the cost-table kernel, never while partitions are being searched.
struct SubSlot { int16_t mv[2]; int16_t mvp[2]; int32_t cost, cost_mv; };
enum { SUB_4x4 = 0, SUB_8x4 = 1, SUB_4x8 = 2, SUB_8x8 = 3 };       // D_L0_4x4 .. D_L0_8x8 (common/macroblock.h:115-118)
PCAMV_DEV SubSlot *sub_slots(MbWork &w, int i8) { return (SubSlot *)&w.coef[0][0] + 8 * i8; }
PCAMV_DEV int sub_first(int kind) { return kind == SUB_4x4 ? 0 : kind == SUB_8x4 ? 4 : 6; }
PCAMV_DEV int sub_count(int kind) { return kind == SUB_4x4 ? 4 : kind == SUB_8x8 ? 1 : 2; }
PCAMV_DEV int sub_pixel(int kind) { return kind == SUB_4x4 ? PIX_4x4 : kind == SUB_8x4 ? PIX_8x4 : kind == SUB_4x8 ? PIX_4x8 : PIX_8x8; }
// offset of sub-block j of an 8x8 block split as `kind`, in pixels
PCAMV_DEV int sub_xoff(int kind, int j) { return kind == SUB_4x4 ? 4 * (j & 1) : kind == SUB_4x8 ? 4 * j : 0; }
PCAMV_DEV int sub_yoff(int kind, int j) { return kind == SUB_4x4 ? 4 * (j >> 1) : kind == SUB_8x4 ? 4 * j : 0; }

// Loads of per-frame motion state written by OTHER macroblocks of the running wavefront (possibly on another SM):
// they bypass the non-coherent L1 (ld.global.cg).  Ordering against the writer is the row-progress flag
// (release/acquire, pcamv_frame_kernels.cu).
#if defined(PCAMV_EMU)
  #define PCAMV_LDV(p) (*(p))
#else
  #define PCAMV_LDV(p) __ldcg(p)
#endif

PCAMV_DEV int scan8(int idx)
{
    const int x = (idx & 1) | ((idx >> 1) & 2);
    const int y = ((idx >> 1) & 1) | ((idx >> 2) & 2);
    return 4 + x + 8 * (1 + y);
}
PCAMV_DEV uint32_t pack_mv(int x, int y) { return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16); }
PCAMV_DEV int mv_x(uint32_t p) { return (int)(int16_t)(p & 0xffffu); }
PCAMV_DEV int mv_y(uint32_t p) { return (int)(int16_t)(p >> 16); }

// everything the per-MB code needs, by reference
struct MbCtx
{
    const DevFrameCtx &fc;
    const FrameParams &fp;
    MbWork &w;
    int mb_x, mb_y, mb_xy;
    int type_left, type_top, type_topleft, type_topright;
    int partition;                        // h->mb.i_partition as seen by x264_mb_predict_mv
    int mv_min[2], mv_max[2];             // full-range qpel limits (MC clipping)
    MeEnv env;
    int n_log;
    int pskip_mv[2];
    PCAMV_MEM MbCtx(const DevFrameCtx &f, const FrameParams &p, MbWork &wk) : fc(f), fp(p), w(wk) {}
};

// generic team work split: on the GPU lane l does item l; in emulation the single lane loops over all items
#if defined(PCAMV_EMU)
  #define PCAMV_FOR_ITEMS(it, n) for (int it = 0; it < (n); it++)
#else
  #define PCAMV_FOR_ITEMS(it, n) for (int it = team_lane(); it < (n); it += 32)
#endif

PCAMV_FN void log_push(MbCtx &c, int kind, int i_pixel, int i_ref, int mvx, int mvy, int cost, int cost_mv)
{
    // the counter lives in the team-shared context: every lane reads it before lane 0 alone advances it
    const int n = c.n_log;
    team_sync();
    if (team_lane() == 0)
    {
        if (n < c.fp.log_stride)
        {
            LogEntry e;
            e.kind = (int8_t)kind; e.i_pixel = (int8_t)i_pixel; e.i_ref = (int8_t)i_ref; e.pad = 0;
            e.mv[0] = (int16_t)mvx; e.mv[1] = (int16_t)mvy; e.cost = cost; e.cost_mv = cost_mv;
            c.fp.log[(size_t)c.mb_xy * c.fp.log_stride + n] = e;
        }
        c.n_log = n + 1;
    }
    team_sync();
}

// ---- neighbour cache -------------------------------------------------------------------------------
// one cache entry per lane (the cache is team-shared: every lane reads all of it afterwards, hence the sync)
PCAMV_FN void cache_fill_rect(MbCtx &c, int x, int y, int wd, int ht, int ref, uint32_t mv, int set_ref, int set_mv)
{
    team_sync();                 // nobody may still be reading the entries about to change
    PCAMV_FOR_ITEMS(it, wd * ht)
    {
        const int j = it / wd, i = it - j * wd;
        const int k = 12 + x + i + 8 * (y + j);
        if (set_ref) c.w.ref[k] = (int8_t)ref;
        if (set_mv) c.w.mv[k] = mv;
    }
    team_sync();
}

PCAMV_FN void cache_load(MbCtx &c)
{
    const int mb_w = c.fc.mb_w, s8 = 2 * mb_w, s4 = 4 * mb_w;
    const int mb_x = c.mb_x, mb_y = c.mb_y;
    const FrameArrays &a = c.fp.cur;
    const bool top = mb_y > 0, left = mb_x > 0, topright = top && mb_x < mb_w - 1, topleft = top && left;
    const int top_xy = (mb_y - 1) * mb_w + mb_x;
    const int top8 = (2 * (mb_y - 1) + 1) * s8 + 2 * mb_x, top4 = (4 * (mb_y - 1) + 3) * s4 + 4 * mb_x;
    const int cur8 = 2 * mb_y * s8 + 2 * mb_x, cur4 = 4 * mb_y * s4 + 4 * mb_x;
    c.type_top = top ? PCAMV_LDV(a.type + top_xy) : -1;
    c.type_left = left ? PCAMV_LDV(a.type + c.mb_xy - 1) : -1;
    c.type_topright = topright ? PCAMV_LDV(a.type + top_xy + 1) : -1;
    c.type_topleft = topleft ? PCAMV_LDV(a.type + top_xy - 1) : -1;
    // positions never written for the current MB keep "unavailable" (the reference memsets the cache to -2 once)
    team_sync();
    PCAMV_FOR_ITEMS(k, 48) { c.w.ref[k] = -2; c.w.mv[k] = 0; }
    team_sync();
    if (topleft) { c.w.ref[3] = PCAMV_LDV(a.ref8 + top8 - 1); c.w.mv[3] = PCAMV_LDV(a.mv4 + top4 - 1); }
    if (top)
    {
        c.w.ref[4] = c.w.ref[5] = PCAMV_LDV(a.ref8 + top8); c.w.ref[6] = c.w.ref[7] = PCAMV_LDV(a.ref8 + top8 + 1);
        for (int k = 0; k < 4; k++) c.w.mv[4 + k] = PCAMV_LDV(a.mv4 + top4 + k);
    }
    if (topright) { c.w.ref[8] = PCAMV_LDV(a.ref8 + top8 + 2); c.w.mv[8] = PCAMV_LDV(a.mv4 + top4 + 4); }
    if (left)
    {
        c.w.ref[11] = c.w.ref[19] = PCAMV_LDV(a.ref8 + cur8 - 1);
        c.w.ref[27] = c.w.ref[35] = PCAMV_LDV(a.ref8 + cur8 - 1 + s8);
        for (int k = 0; k < 4; k++) c.w.mv[11 + 8 * k] = PCAMV_LDV(a.mv4 + cur4 - 1 + k * s4);
    }
}

PCAMV_DEV uint32_t median_mv(uint32_t a, uint32_t b, uint32_t cc)
{
    return pack_mv(median3(mv_x(a), mv_x(b), mv_x(cc)), median3(mv_y(a), mv_y(b), mv_y(cc)));
}

PCAMV_DEV uint32_t predict_from(int i_ref, int refa, uint32_t mva, int refb, uint32_t mvb, int refc, uint32_t mvc)
{
    const int count = (refa == i_ref) + (refb == i_ref) + (refc == i_ref);
    if (count > 1) return median_mv(mva, mvb, mvc);
    if (count == 1) return refa == i_ref ? mva : refb == i_ref ? mvb : mvc;
    if (refb == -2 && refc == -2 && refa != -2) return mva;
    return median_mv(mva, mvb, mvc);
}

PCAMV_FN uint32_t predict_mv_16x16(const MbCtx &c, int i_ref)
{
    int refc = c.w.ref[8]; uint32_t mvc = c.w.mv[8];
    if (refc == -2) { refc = c.w.ref[3]; mvc = c.w.mv[3]; }
    return predict_from(i_ref, c.w.ref[11], c.w.mv[11], c.w.ref[4], c.w.mv[4], refc, mvc);
}

PCAMV_FN uint32_t predict_mv(const MbCtx &c, int idx, int width)
{
    const int i8 = scan8(idx);
    const int i_ref = c.w.ref[i8];
    const int refa = c.w.ref[i8 - 1], refb = c.w.ref[i8 - 8];
    const uint32_t mva = c.w.mv[i8 - 1], mvb = c.w.mv[i8 - 8];
    int refc = c.w.ref[i8 - 8 + width]; uint32_t mvc = c.w.mv[i8 - 8 + width];
    if ((idx & 3) == 3 || (width == 2 && (idx & 3) == 2) || refc == -2)
    {
        refc = c.w.ref[i8 - 8 - 1]; mvc = c.w.mv[i8 - 8 - 1];
    }
    if (c.partition == PART_16x8)
    {
        if (idx == 0 && refb == i_ref) return mvb;
        if (idx != 0 && refa == i_ref) return mva;
    }
    else if (c.partition == PART_8x16)
    {
        if (idx == 0 && refa == i_ref) return mva;
        if (idx != 0 && refc == i_ref) return mvc;
    }
    return predict_from(i_ref, refa, mva, refb, mvb, refc, mvc);
}

PCAMV_DEV uint32_t predict_mv_pskip(const MbCtx &c)
{
    const int refa = c.w.ref[11], refb = c.w.ref[4];
    const uint32_t mva = c.w.mv[11], mvb = c.w.mv[4];
    if (refa == -2 || refb == -2 || !(refa | (int)mva) || !(refb | (int)mvb))
        return 0;
    return predict_mv_16x16(c, 0);
}

// candidate list of the 16x16 search: neighbours' 16x16 search results + temporally scaled co-located MVs
PCAMV_FN int predict_mv_ref16x16(const MbCtx &c, int i_ref, int (*mvc)[2])
{
    const int mb_w = c.fc.mb_w, n_mb = mb_w * c.fc.mb_h;
    const uint32_t *mvr = c.fp.cur.mvr + (size_t)i_ref * n_mb;
    const int8_t *type = c.fp.cur.type;
    int n = 0;
#define PCAMV_SET(p) { const uint32_t v_ = (p); mvc[n][0] = mv_x(v_); mvc[n][1] = mv_y(v_); n++; }
    if (c.mb_x > 0 && PCAMV_LDV(type + c.mb_xy - 1) != MB_P_SKIP) PCAMV_SET(PCAMV_LDV(mvr + c.mb_xy - 1));
    if (c.mb_y > 0)
    {
        const int t = c.mb_xy - mb_w;
        if (PCAMV_LDV(type + t) != MB_P_SKIP) PCAMV_SET(PCAMV_LDV(mvr + t));
        if (c.mb_x > 0 && PCAMV_LDV(type + t - 1) != MB_P_SKIP) PCAMV_SET(PCAMV_LDV(mvr + t - 1));
        if (c.mb_x < mb_w - 1 && PCAMV_LDV(type + t + 1) != MB_P_SKIP) PCAMV_SET(PCAMV_LDV(mvr + t + 1));
    }
#undef PCAMV_SET
    if (c.fp.col_n_ref > 0)
    {
        const int s8 = 2 * mb_w, s4 = 4 * mb_w;
        const int cur8 = 2 * c.mb_y * s8 + 2 * c.mb_x, cur4 = 4 * c.mb_y * s4 + 4 * c.mb_x;
#pragma unroll 1
        for (int k = 0; k < 3; k++)
        {
            const int dx = k == 1, dy = k == 2;
            if (k == 1 && !(c.mb_x < mb_w - 1)) continue;
            if (k == 2 && !(c.mb_y < c.fc.mb_h - 1)) continue;
            const int ref_col = c.fp.col_ref8[cur8 + dx * 2 + dy * 2 * s8];
            if (ref_col >= 0)
            {
                const int scale = (c.fp.cur_poc - c.fp.ref_poc[i_ref]) * c.fp.col_inv_ref_poc[ref_col];
                const uint32_t m = c.fp.col_mv4[cur4 + dx * 4 + dy * 4 * s4];
                mvc[n][0] = (int16_t)((mv_x(m) * scale + 128) >> 8);
                mvc[n][1] = (int16_t)((mv_y(m) * scale + 128) >> 8);
                n++;
            }
        }
    }
    return n;
}

// ---- MV limits (reference encoder/analyse.c:271-318; single thread, progressive) -------------------
PCAMV_FN void init_limits(MbCtx &c)
{
    const int fmv = 4 * c.fc.mv_range;
    c.mv_min[0] = 4 * (-16 * c.mb_x - 24);
    c.mv_max[0] = 4 * (16 * (c.fc.mb_w - c.mb_x - 1) + 24);
    c.mv_min[1] = 4 * (-16 * c.mb_y - 24);
    c.mv_max[1] = 4 * (16 * (c.fc.mb_h - c.mb_y - 1) + 24);
    c.env.mv_min_spel[0] = clip3(c.mv_min[0], -fmv, fmv - 1);
    c.env.mv_max_spel[0] = clip3(c.mv_max[0], -fmv, fmv - 1);
    c.env.mv_min_spel[1] = clip3(c.mv_min[1], imax(4 * (-512 + 8), -fmv), fmv);
    c.env.mv_max_spel[1] = imin(clip3(c.mv_max[1], -fmv, fmv - 1), fmv * 4);
    for (int k = 0; k < 2; k++)
    {
        c.env.mv_min_fpel[k] = (c.env.mv_min_spel[k] >> 2) + 5;
        c.env.mv_max_fpel[k] = (c.env.mv_max_spel[k] >> 2) - 5;
    }
}

// ---- 4x4 transform / quantisation (one 4x4 block per lane) -----------------------------------------
// d[] = fenc - pred residual in raster order (d[4*y+x]); out = coefficients in the reference's dct[i][j] layout
// ROLL = 1 keeps the 4- and 16-element loops of the transform / quantiser rolled (arrays in local memory): the wavefront
// kernel only needs them for the P_SKIP probe of some macroblocks, and there a small body beats a fast one — the launch
// is bound by the footprint of its code (DESIGN.md "What limits it").  The cost-table kernel uses the unrolled forms.
template <int ROLL>
PCAMV_DEV void dct4x4(const int d[16], int out[16])
{
    int tmp[16];
#pragma unroll (ROLL ? 1 : 4)
    for (int i = 0; i < 4; i++)
    {
        const int s03 = d[4 * i] + d[4 * i + 3], s12 = d[4 * i + 1] + d[4 * i + 2];
        const int d03 = d[4 * i] - d[4 * i + 3], d12 = d[4 * i + 1] - d[4 * i + 2];
        tmp[0 * 4 + i] = s03 + s12; tmp[1 * 4 + i] = 2 * d03 + d12; tmp[2 * 4 + i] = s03 - s12; tmp[3 * 4 + i] = d03 - 2 * d12;
    }
#pragma unroll (ROLL ? 1 : 4)
    for (int i = 0; i < 4; i++)
    {
        const int s03 = tmp[4 * i] + tmp[4 * i + 3], s12 = tmp[4 * i + 1] + tmp[4 * i + 2];
        const int d03 = tmp[4 * i] - tmp[4 * i + 3], d12 = tmp[4 * i + 1] - tmp[4 * i + 2];
        out[4 * i] = (int16_t)(s03 + s12); out[4 * i + 1] = (int16_t)(2 * d03 + d12);
        out[4 * i + 2] = (int16_t)(s03 - s12); out[4 * i + 3] = (int16_t)(d03 - 2 * d12);
    }
}
// in-place quantisation with dead-zone bias; returns nonzero flag
template <int ROLL>
PCAMV_DEV int quant4x4(int coef[16], const uint16_t *mf, const uint16_t *bias)
{
    int nz = 0;
#pragma unroll (ROLL ? 1 : 16)
    for (int i = 0; i < 16; i++)
    {
        const int v = coef[i];
        const int q = v > 0 ? ((bias[i] + v) * mf[i]) >> 16 : -(((bias[i] - v) * mf[i]) >> 16);
        coef[i] = (int16_t)q;
        nz |= q;
    }
    return nz != 0;
}
PCAMV_DEV int quant_one(int v, int mf, int bias)
{
    return (int16_t)(v > 0 ? ((bias + v) * mf) >> 16 : -(((bias - v) * mf) >> 16));
}
// decimation score of a quantised 4x4 block held in memory; first = 0 (all 16 coefficients) or 1 (AC only)
PCAMV_DEV int decimate_score(const int16_t *coef, int first)
{
    // coefficient i of the frame zigzag scan lives at dct[x][y] (flattened 4*x+y): 0 4 1 2 5 8 12 9 6 3 7 10 13 14 11 15
    const unsigned long long zz = 0xfbeda7369c852140ull;
    int idx = 15;
    while (idx >= first && coef[(zz >> (4 * idx)) & 15] == 0) idx--;
    int score = 0;
    while (idx >= first)
    {
        const int v = coef[(zz >> (4 * idx)) & 15];
        idx--;
        if ((unsigned)(v + 1) > 2) return 9;
        int run = 0;
        while (idx >= first && coef[(zz >> (4 * idx)) & 15] == 0) { idx--; run++; }
        score += run == 0 ? 3 : run <= 2 ? 2 : run <= 5 ? 1 : 0;        // x264_decimate_table4
    }
    return score;
}
template <int ROLL>
PCAMV_DEV void dequant4x4(int coef[16], const int32_t *dequant_mf, int qp)
{
    const int32_t *dm = dequant_mf + (qp % 6) * 16;
    const int qbits = qp / 6 - 4;
    if (qbits >= 0)
    {
#pragma unroll (ROLL ? 1 : 16)
        for (int i = 0; i < 16; i++) coef[i] = (int16_t)((coef[i] * dm[i]) << qbits);
    }
    else
    {
        const int f = 1 << (-qbits - 1);
#pragma unroll (ROLL ? 1 : 16)
        for (int i = 0; i < 16; i++) coef[i] = (int16_t)((coef[i] * dm[i] + f) >> (-qbits));
    }
}
// residual block r[4*y+x] to add to the prediction
PCAMV_DEV void idct4x4(const int coef[16], int r[16])
{
    int tmp[16];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int s02 = coef[0 * 4 + i] + coef[2 * 4 + i], d02 = coef[0 * 4 + i] - coef[2 * 4 + i];
        const int s13 = coef[1 * 4 + i] + (coef[3 * 4 + i] >> 1), d13 = (coef[1 * 4 + i] >> 1) - coef[3 * 4 + i];
        tmp[4 * i] = (int16_t)(s02 + s13); tmp[4 * i + 1] = (int16_t)(d02 + d13);
        tmp[4 * i + 2] = (int16_t)(d02 - d13); tmp[4 * i + 3] = (int16_t)(s02 - s13);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int s02 = tmp[0 * 4 + i] + tmp[2 * 4 + i], d02 = tmp[0 * 4 + i] - tmp[2 * 4 + i];
        const int s13 = tmp[1 * 4 + i] + (tmp[3 * 4 + i] >> 1), d13 = (tmp[1 * 4 + i] >> 1) - tmp[3 * 4 + i];
        r[0 * 4 + i] = (int16_t)((s02 + s13 + 32) >> 6); r[1 * 4 + i] = (int16_t)((d02 + d13 + 32) >> 6);
        r[2 * 4 + i] = (int16_t)((d02 - d13 + 32) >> 6); r[3 * 4 + i] = (int16_t)((s02 - s13 + 32) >> 6);
    }
}

PCAMV_DEV void load_residual(const uint8_t *fenc, int fs, const uint8_t *pred, int ps, int d[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        const uint32_t f = ld4a(fenc + y * fs), p = ld4a(pred + y * ps);
#pragma unroll
        for (int x = 0; x < 4; x++) d[4 * y + x] = px(f, x) - px(p, x);
    }
}
PCAMV_DEV void add_residual(uint8_t *pred, int ps, const int r[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        const uint32_t p = ld4a(pred + y * ps);
        uint32_t o = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) o |= (uint32_t)clip_u8(px(p, x) + r[4 * y + x]) << (8 * x);
#if defined(PCAMV_EMU)
        memcpy(pred + y * ps, &o, 4);
#else
        *(uint32_t *)(pred + y * ps) = o;
#endif
    }
}
PCAMV_DEV void st4a(uint8_t *p, uint32_t v)
{
#if defined(PCAMV_EMU)
    memcpy(p, &v, 4);
#else
    *(uint32_t *)p = v;
#endif
}

PCAMV_DEV int team_any(int v)
{
#if defined(PCAMV_EMU)
    return v != 0;
#else
    return __any_sync(0xffffffffu, v != 0);
#endif
}

// ---- motion compensation of a rectangle into the staging buffers -------------------------------------
// luma w x h at block-relative (x0,y0) with quarter-pel MV (already clipped); chroma analogously.
PCAMV_FN void mc_rect(MbCtx &c, int ref_slot, int x0, int y0, int wd, int ht, int qmx, int qmy)
{
    const DevRef &rf = c.fc.ref[ref_slot];
    MeBlock b;
    b.stride = c.fc.stride_y; b.stride_c = c.fc.stride_c;
    const ptrdiff_t off = (ptrdiff_t)(16 * c.mb_y + y0) * b.stride + 16 * c.mb_x + x0;
    for (int k = 0; k < 4; k++) b.ref[k] = rf.y[k] + off;
    const uint8_t *s1, *s2;
    qpel_sources(b, qmx, qmy, s1, s2);
#if defined(PCAMV_CHECKED)
    b.chk_lo = rf.base; b.chk_hi = rf.base + rf.bytes;
    PCAMV_CHK_RANGE(b, (s1 < s2 ? s1 : s2) - 3, (s1 > s2 ? s1 : s2) + (ht - 1) * b.stride + wd + 7, "mc_rect (luma)");
    {
        const uint8_t *cu = rf.u + (ptrdiff_t)(8 * c.mb_y + (y0 >> 1) + (qmy >> 3)) * b.stride_c + 8 * c.mb_x + (x0 >> 1) + (qmx >> 3);
        const uint8_t *cv = rf.v + (ptrdiff_t)(8 * c.mb_y + (y0 >> 1) + (qmy >> 3)) * b.stride_c + 8 * c.mb_x + (x0 >> 1) + (qmx >> 3);
        PCAMV_CHK_RANGE(b, cu - 3, cu + (ht >> 1) * b.stride_c + (wd >> 1) + 1 + 7, "mc_rect (U)");
        PCAMV_CHK_RANGE(b, cv - 3, cv + (ht >> 1) * b.stride_c + (wd >> 1) + 1 + 7, "mc_rect (V)");
    }
#endif
    const int w4 = wd >> 2;
    PCAMV_FOR_ITEMS(it, w4 * ht)
    {
        const int y = it / w4, x = (it - y * w4) << 2;
        st4a(c.w.pred_y + (y0 + y) * 16 + x0 + x, pred4(s1, s2, b.stride, x, y));
    }
    const int cw4 = wd >> 3, chh = ht >> 1;      // chroma words per row / rows
    const ptrdiff_t offc = (ptrdiff_t)(8 * c.mb_y + (y0 >> 1)) * b.stride_c + 8 * c.mb_x + (x0 >> 1);
    if (cw4 == 0)
    {
        // 4 luma pixels wide: two chroma pixels per row
        PCAMV_FOR_ITEMS(it, 2 * chh)
        {
            const int pl = it / chh, y = it - pl * chh;
            const uint8_t *src = (pl ? rf.v : rf.u) + offc;
            uint8_t *dst = (pl ? c.w.pred_v : c.w.pred_u) + ((y0 >> 1) + y) * 8 + (x0 >> 1);
            const uint32_t v = chroma4(src, b.stride_c, qmx, qmy, 0, y);
            dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8);
        }
        team_sync();
        return;
    }
    PCAMV_FOR_ITEMS(it, 2 * cw4 * chh)
    {
        const int pl = it / (cw4 * chh), r = it - pl * cw4 * chh;
        const int y = r / cw4, x = (r - y * cw4) << 2;
        const uint8_t *src = (pl ? rf.v : rf.u) + offc;
        uint8_t *dst = (pl ? c.w.pred_v : c.w.pred_u) + ((y0 >> 1) + y) * 8 + (x0 >> 1) + x;
        st4a(dst, chroma4(src, b.stride_c, qmx, qmy, x, y));
    }
    team_sync();
}

// ---- transform / quantisation of one 4x4 block of the macroblock (the one expanded copy of the 16-coefficient code) ----
// Block `it`: 0..15 luma (block_idx order), 16..19 U, 20..23 V.  Residual = fenc - staged prediction; DCT; for chroma
// the DC term is set aside (raw, in dcs[]) and zeroed; quantise; decimate score; dequantise.  The coefficients end up
// in w.coef[it] (dequantised when non-zero).  Returns score | nz << 8.
// (reference encoder/macroblock.c:605-755 luma, :277-372 chroma, :809-895 probe; common/dct.c:122-162; quant.c:33-75,203-252)
template <int ROLL>
PCAMV_FN int quant_block(MbCtx &c, int it, int16_t *dcs)
{
    const DevTables &t = c.fc.tab;
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
    dct4x4<ROLL>(d, co);
    if (ch)
    {
        dcs[it - 16] = (int16_t)co[0];
        co[0] = 0;
    }
    const int nz = quant4x4<ROLL>(co, t.quant4_mf[ch], t.quant4_bias[ch]);
    int16_t *out = c.w.coef[it];
    int score = 0;
    if (nz)
    {
#pragma unroll (ROLL ? 1 : 16)
        for (int i = 0; i < 16; i++) out[i] = (int16_t)co[i];
        score = decimate_score(out, ch);
        dequant4x4<ROLL>(co, t.dequant4_mf[ch], ch ? t.chroma_qp : t.qp);
#pragma unroll (ROLL ? 1 : 16)
        for (int i = 0; i < 16; i++) out[i] = (int16_t)co[i];
    }
    return score | (nz << 8);
}

// Reconstruction of block `it` into the staged prediction: mode 0 = IDCT of w.coef[it]; 1 = chroma DC only;
// 2 = chroma IDCT with the dequantised DC term `dc` put back  (common/dct.c:174-262,364-384)
PCAMV_FN void recon_block(MbCtx &c, int it, int mode, int dc)
{
    int co[16], r[16];
    uint8_t *pr; int ps;
    if (it < 16)
    {
        const int bx = (it & 1) | ((it >> 1) & 2), by = ((it >> 1) & 1) | ((it >> 2) & 2);
        pr = c.w.pred_y + 64 * by + 4 * bx; ps = 16;
    }
    else
    {
        const int pl = (it - 16) >> 2, blk = (it - 16) & 3;
        pr = (pl ? c.w.pred_v : c.w.pred_u) + 32 * (blk >> 1) + 4 * (blk & 1); ps = 8;
    }
    if (mode == 1)
    {
        const int v = (int16_t)((dc + 32) >> 6);
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = v;
    }
    else
    {
        const int16_t *in = c.w.coef[it];
#pragma unroll
        for (int i = 0; i < 16; i++) co[i] = in[i];
        if (mode == 2) co[0] = dc;
        idct4x4(co, r);
    }
    add_residual(pr, ps, r);
}

// ---- P_SKIP probe -----------------------------------------------------------------------------------------
PCAMV_FN int probe_pskip(MbCtx &c)
{
    const int mvx = clip3(c.pskip_mv[0], c.mv_min[0], c.mv_max[0]);
    const int mvy = clip3(c.pskip_mv[1], c.mv_min[1], c.mv_max[1]);
    mc_rect(c, c.fp.ref_slot[0], 0, 0, 16, 16, mvx, mvy);
    const DevTables &t = c.fc.tab;
    int16_t *dcs = (int16_t *)(c.w.scratch + 24);           // 8 x int16: raw chroma DC terms
    // luma: 16 blocks, total decimate score must stay below 6
    int score = 0;
    PCAMV_FOR_ITEMS(blk, 16)
    {
        const int v = quant_block<1>(c, blk, dcs);
        if (v >> 8) score += v & 0xff;
    }
    score = team_sum(score);
    if (score >= 6)
        return 0;
    // chroma: SSD gate, then DC must quantise to zero and the AC decimate score must stay below 7, per plane
    const int thresh = (t.lambda2_chroma + 32) >> 6;
    for (int pl = 0; pl < 2; pl++)
    {
        const uint8_t *fe = pl ? c.w.fenc_v : c.w.fenc_u, *pr = pl ? c.w.pred_v : c.w.pred_u;
        int ssd = 0;
        PCAMV_FOR_ITEMS(it, 16)
        {
            const uint32_t f = ld4a(fe + 4 * it), p = ld4a(pr + 4 * it);
            for (int k = 0; k < 4; k++) { const int dd = px(f, k) - px(p, k); ssd += dd * dd; }
        }
        ssd = team_sum(ssd);
        if (ssd < thresh) continue;
        int sc = 0;
        PCAMV_FOR_ITEMS(blk, 4)
        {
            const int v = quant_block<1>(c, 16 + 4 * pl + blk, dcs);
            if (v >> 8) sc += v & 0xff;
        }
        team_sync();
        const int b0 = dcs[4 * pl], b1 = dcs[4 * pl + 1], b2 = dcs[4 * pl + 2], b3 = dcs[4 * pl + 3];
        const int d0 = b0 + b1, d1 = b2 + b3, d2 = b0 - b1, d3 = b2 - b3;
        const int mf = t.quant4_mf[1][0] >> 1, bias = t.quant4_bias[1][0] << 1;
        if (quant_one((int16_t)(d0 + d1), mf, bias) | quant_one((int16_t)(d2 + d3), mf, bias) |
            quant_one((int16_t)(d0 - d1), mf, bias) | quant_one((int16_t)(d2 - d3), mf, bias))
            return 0;
        sc = team_sum(sc);
        if (sc >= 7)
            return 0;
        team_sync();
    }
    return 1;
}

// =======================================================================================================
// Search drivers
// =======================================================================================================
PCAMV_DEV void setup_block(const MbCtx &c, MeBlock &b, int i_ref, int i_pixel, int xoff, int yoff)
{
    const DevRef &rf = c.fc.ref[c.fp.ref_slot[i_ref]];
    b.i_pixel = i_pixel; b.bw = pix_w(i_pixel); b.bh = pix_h(i_pixel);
    b.fenc = c.w.fenc_y + yoff * 16 + xoff;
    b.fenc_u = c.w.fenc_u + (yoff >> 1) * 8 + (xoff >> 1);
    b.fenc_v = c.w.fenc_v + (yoff >> 1) * 8 + (xoff >> 1);
    b.stride = c.fc.stride_y; b.stride_c = c.fc.stride_c;
    const ptrdiff_t off = (ptrdiff_t)(16 * c.mb_y + yoff) * b.stride + 16 * c.mb_x + xoff;
    for (int k = 0; k < 4; k++) b.ref[k] = rf.y[k] + off;
    const ptrdiff_t offc = (ptrdiff_t)(8 * c.mb_y + (yoff >> 1)) * b.stride_c + 8 * c.mb_x + (xoff >> 1);
    b.ref_u = rf.u + offc; b.ref_v = rf.v + offc;
    b.integral = rf.integral ? rf.integral + off : nullptr;
    b.integral4 = rf.integral4 ? rf.integral4 + off : nullptr;
#if defined(PCAMV_CHECKED)
    b.chk_lo = rf.base; b.chk_hi = rf.base + rf.bytes;
#endif
}

template <int XS>
PCAMV_FN void run_search(MbCtx &c, MeSlot &s, uint32_t mvp, const int (*mvc)[2], int i_mvc, int *thresh)
{
    MeBlock &b = c.w.blk;
    setup_block(c, b, s.i_ref, s.i_pixel, s.xoff, s.yoff);
    s.mvp[0] = mv_x(mvp); s.mvp[1] = mv_y(mvp);
    block_set_mvp(b, c.env, s.mvp[0], s.mvp[1]);
    s.r.mv[0] = s.r.mv[1] = 0; s.r.cost = 0; s.r.cost_mv = 0;
    me_search_ref<XS>(c.env, b, mvc, i_mvc, thresh, s.r);
    log_push(c, LOG_SEARCH, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
}

PCAMV_FN void run_refine(MbCtx &c, MeSlot &s)
{
    MeBlock &b = c.w.blk;
    setup_block(c, b, s.i_ref, s.i_pixel, s.xoff, s.yoff);
    block_set_mvp(b, c.env, s.mvp[0], s.mvp[1]);
    me_refine_qpel(c.env, b, s.r, s.i_ref_cost);
    log_push(c, LOG_REFINE, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
}

// ---- resumable form (split wavefront, pcamv_split.cu) -------------------------------------------------------------
// A search is issued in two halves.  AS = 0 (synchronous kernels, emulation): run_search_begin does the whole search and
// returns 0.  AS = 1 (split wavefront): it only writes the request into w.rq and returns 1 — the caller records where it
// stands and unwinds with PT_YIELD; when the result is back in w.rs the same code runs again, skips to that point and
// takes it with run_search_end.
template <int XS, int AS>
PCAMV_FN int run_search_begin(MbCtx &c, MeSlot &s, uint32_t mvp, const int (*mvc)[2], int i_mvc, int *thresh)
{
    s.mvp[0] = mv_x(mvp); s.mvp[1] = mv_y(mvp);
    s.r.mv[0] = s.r.mv[1] = 0; s.r.cost = 0; s.r.cost_mv = 0;
    if (!AS)
    {
        MeBlock &b = c.w.blk;
        setup_block(c, b, s.i_ref, s.i_pixel, s.xoff, s.yoff);
        block_set_mvp(b, c.env, s.mvp[0], s.mvp[1]);
        me_search_ref<XS>(c.env, b, mvc, i_mvc, thresh, s.r);
        log_push(c, LOG_SEARCH, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
        return 0;
    }
    team_sync();
    if (team_lane() == 0)
    {
        SearchReq &q = c.w.rq;
        q.mb_x = (int16_t)c.mb_x; q.mb_y = (int16_t)c.mb_y;
        q.kind = LOG_SEARCH; q.i_ref = (int8_t)s.i_ref; q.i_pixel = (int8_t)s.i_pixel; q.i_mvc = (int8_t)i_mvc;
        q.xoff = (int8_t)s.xoff; q.yoff = (int8_t)s.yoff; q.has_thresh = thresh != nullptr; q.pad = 0;
        q.mvp = mvp; q.thresh = thresh ? *thresh : 0;
        q.mv = 0; q.cost = 0; q.cost_mv = 0; q.i_ref_cost = s.i_ref_cost;
        for (int k = 0; k < 2; k++)
        {
            q.lim[k] = (int16_t)c.env.mv_min_fpel[k]; q.lim[2 + k] = (int16_t)c.env.mv_max_fpel[k];
            q.lim[4 + k] = (int16_t)c.env.mv_min_spel[k]; q.lim[6 + k] = (int16_t)c.env.mv_max_spel[k];
        }
        for (int k = 0; k < i_mvc; k++) q.mvc[k] = pack_mv(mvc[k][0], mvc[k][1]);
    }
    team_sync();
    return 1;
}
template <int AS>
PCAMV_DEV void run_search_end(MbCtx &c, MeSlot &s, int *thresh)
{
    if (!AS)
        return;
    const SearchRes r = c.w.rs;
    team_sync();
    s.r.mv[0] = mv_x(r.mv); s.r.mv[1] = mv_y(r.mv); s.r.cost = r.cost; s.r.cost_mv = r.cost_mv;
    if (thresh) *thresh = r.thresh;
    team_sync();
    log_push(c, LOG_SEARCH, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
}
template <int AS>
PCAMV_FN int run_refine_begin(MbCtx &c, MeSlot &s)
{
    if (!AS)
    {
        MeBlock &b = c.w.blk;
        setup_block(c, b, s.i_ref, s.i_pixel, s.xoff, s.yoff);
        block_set_mvp(b, c.env, s.mvp[0], s.mvp[1]);
        me_refine_qpel(c.env, b, s.r, s.i_ref_cost);
        log_push(c, LOG_REFINE, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
        return 0;
    }
    team_sync();
    if (team_lane() == 0)
    {
        SearchReq &q = c.w.rq;
        q.mb_x = (int16_t)c.mb_x; q.mb_y = (int16_t)c.mb_y;
        q.kind = LOG_REFINE; q.i_ref = (int8_t)s.i_ref; q.i_pixel = (int8_t)s.i_pixel; q.i_mvc = 0;
        q.xoff = (int8_t)s.xoff; q.yoff = (int8_t)s.yoff; q.has_thresh = 0; q.pad = 0;
        q.mvp = pack_mv(s.mvp[0], s.mvp[1]); q.thresh = 0;
        q.mv = pack_mv(s.r.mv[0], s.r.mv[1]); q.cost = s.r.cost; q.cost_mv = s.r.cost_mv; q.i_ref_cost = s.i_ref_cost;
        for (int k = 0; k < 2; k++)
        {
            q.lim[k] = (int16_t)c.env.mv_min_fpel[k]; q.lim[2 + k] = (int16_t)c.env.mv_max_fpel[k];
            q.lim[4 + k] = (int16_t)c.env.mv_min_spel[k]; q.lim[6 + k] = (int16_t)c.env.mv_max_spel[k];
        }
    }
    team_sync();
    return 1;
}
template <int AS>
PCAMV_DEV void run_refine_end(MbCtx &c, MeSlot &s)
{
    if (!AS)
        return;
    const SearchRes r = c.w.rs;
    team_sync();
    s.r.mv[0] = mv_x(r.mv); s.r.mv[1] = mv_y(r.mv); s.r.cost = r.cost; s.r.cost_mv = r.cost_mv;
    team_sync();
    log_push(c, LOG_REFINE, s.i_pixel, s.i_ref, s.r.mv[0], s.r.mv[1], s.r.cost, s.r.cost_mv);
}

// What a search team does with a request (split wavefront; the emulation checker serves its own requests with it): the block
// descriptor and search limits are rebuilt from the request, then the same me_search_ref / me_refine_qpel run.
// `w` only lends its source-pixel staging (fenc_*, already holding the macroblock) and block descriptor.
template <int XS>
PCAMV_FN void serve_request(const DevFrameCtx &fc, const FrameParams &fp, MbWork &w, const SearchReq &q, SearchRes &out)
{
    MbCtx c(fc, fp, w);
    c.mb_x = q.mb_x; c.mb_y = q.mb_y; c.mb_xy = q.mb_y * fc.mb_w + q.mb_x;
    MeEnv &env = c.env;
    env.cost_mv = fc.tab.cost_mv;
    env.cost_mv_fpel[0] = env.cost_mv_fpel[1] = env.cost_mv_fpel[2] = env.cost_mv_fpel[3] = nullptr;
    env.me_method = fc.me_method; env.me_range = fc.me_range; env.subme = fc.subme;
    env.chroma_me = fc.chroma_me && fc.subme >= 5;
    env.mbcmp_satd = fc.subme > 1;
    env.mvsads = fp.mvsads ? fp.mvsads + (size_t)q.mb_y * fp.mvsads_cap : nullptr;
    for (int k = 0; k < 2; k++)
    {
        env.mv_min_fpel[k] = q.lim[k]; env.mv_max_fpel[k] = q.lim[2 + k];
        env.mv_min_spel[k] = q.lim[4 + k]; env.mv_max_spel[k] = q.lim[6 + k];
    }
    MeBlock &b = w.blk;
    setup_block(c, b, q.i_ref, q.i_pixel, q.xoff, q.yoff);
    block_set_mvp(b, env, mv_x(q.mvp), mv_y(q.mvp));
    MeResult m;
    int thresh = q.thresh;
    if (q.kind == LOG_SEARCH)
    {
        int mvc[PCAMV_MAX_MVC][2];
        for (int k = 0; k < q.i_mvc; k++) { mvc[k][0] = mv_x(q.mvc[k]); mvc[k][1] = mv_y(q.mvc[k]); }
        m.mv[0] = m.mv[1] = 0; m.cost = 0; m.cost_mv = 0;
        me_search_ref<XS>(env, b, mvc, q.i_mvc, q.has_thresh ? &thresh : nullptr, m);
    }
    else
    {
        m.mv[0] = mv_x(q.mv); m.mv[1] = mv_y(q.mv); m.cost = q.cost; m.cost_mv = q.cost_mv;
        me_refine_qpel(env, b, m, q.i_ref_cost);
    }
    out.mv = pack_mv(m.mv[0], m.mv[1]); out.cost = m.cost; out.cost_mv = m.cost_mv; out.thresh = thresh;
}

PCAMV_DEV int ref_cost(const MbCtx &c, int i_ref)
{
    return c.fc.tab.cost_ref[clip3(c.fp.n_ref - 1, 0, 2) * 33 + i_ref];
}

// 16x16 search over all references; returns 1 when the early P_SKIP termination fired
template <int XS>
PCAMV_FN int analyse_p16x16(MbCtx &c, MbAnalysis &a, int allow_skip, int b_try_pskip)
{
    const int lambda = c.fc.tab.lambda;
    int &halfpel_thresh = c.w.halfpel_thresh;
    halfpel_thresh = 0x7fffffff;
    int *p_thresh = c.fp.n_ref > 1 ? &halfpel_thresh : nullptr;
    a.me16x16.r.cost = 0x7fffffff;
#pragma unroll 1
    for (int i_ref = 0; i_ref < c.fp.n_ref; i_ref++)
    {
        MeSlot &m = c.w.slot;
        const int rc = ref_cost(c, i_ref);
        halfpel_thresh -= rc;
        m.i_ref = i_ref; m.i_ref_cost = rc; m.i_pixel = PIX_16x16; m.xoff = 0; m.yoff = 0;
        int (*mvc)[2] = c.w.mvc;
        const uint32_t mvp = predict_mv_16x16(c, i_ref);
        const int i_mvc = predict_mv_ref16x16(c, i_ref, mvc);
        run_search<XS>(c, m, mvp, mvc, i_mvc, p_thresh);
        if (allow_skip && i_ref == 0 && b_try_pskip && m.r.cost - m.r.cost_mv < 300 * lambda &&
            iabs(m.r.mv[0] - c.pskip_mv[0]) + iabs(m.r.mv[1] - c.pskip_mv[1]) <= 1 && probe_pskip(c))
            return 1;
        m.r.cost += rc;
        halfpel_thresh += rc;
        if (m.r.cost < a.me16x16.r.cost)
            a.me16x16 = m;
        a.mvc[i_ref][0][0] = m.r.mv[0]; a.mvc[i_ref][0][1] = m.r.mv[1];
        if (team_lane() == 0)
            c.fp.cur.mvr[(size_t)i_ref * c.fc.mb_w * c.fc.mb_h + c.mb_xy] = pack_mv(m.r.mv[0], m.r.mv[1]);
    }
    cache_fill_rect(c, 0, 0, 4, 4, a.me16x16.i_ref, 0, 1, 0);
    return 0;
}

template <int XS>
PCAMV_FN void analyse_p8x8(MbCtx &c, MbAnalysis &a)
{
    const int i_ref = a.me16x16.i_ref;
    const int rc = (c.fc.b_cabac || i_ref) ? ref_cost(c, i_ref) : 0;
    int (*mvc)[2] = a.mvc[i_ref];
    c.partition = PART_8x8;
    int i_mvc = 1;
    mvc[0][0] = a.me16x16.r.mv[0]; mvc[0][1] = a.me16x16.r.mv[1];
#pragma unroll 1
    for (int i = 0; i < 4; i++)
    {
        MeSlot &m = a.me8x8[i];
        const int x8 = i & 1, y8 = i >> 1;
        m.i_ref = i_ref; m.i_ref_cost = rc; m.i_pixel = PIX_8x8; m.xoff = 8 * x8; m.yoff = 8 * y8;
        run_search<XS>(c, m, predict_mv(c, 4 * i, 2), mvc, i_mvc, nullptr);
        cache_fill_rect(c, 2 * x8, 2 * y8, 2, 2, 0, pack_mv(m.r.mv[0], m.r.mv[1]), 0, 1);
        mvc[i_mvc][0] = m.r.mv[0]; mvc[i_mvc][1] = m.r.mv[1];
        i_mvc++;
        m.r.cost += rc;
        m.r.cost += c.fc.tab.lambda * 1;          // sub-partition type cost of an unsplit 8x8
    }
    a.cost8x8 = a.me8x8[0].r.cost + a.me8x8[1].r.cost + a.me8x8[2].r.cost + a.me8x8[3].r.cost;
    if (c.fc.b_cabac)
        a.cost8x8 -= rc;
    a.sub[0] = a.sub[1] = a.sub[2] = a.sub[3] = SUB_8x8;
}

// 16x8 (dir = 0) or 8x16 (dir = 1)
template <int XS>
PCAMV_FN void analyse_p16x8_8x16(MbCtx &c, MbAnalysis &a, int dir)
{
    c.partition = dir ? PART_8x16 : PART_16x8;
    int total = 0;
#pragma unroll 1
    for (int i = 0; i < 2; i++)
    {
        MeSlot &best = dir ? a.me8x16[i] : a.me16x8[i];
        const int r0 = dir ? a.me8x8[i].i_ref : a.me8x8[2 * i].i_ref;
        const int r1 = dir ? a.me8x8[i + 2].i_ref : a.me8x8[2 * i + 1].i_ref;
        const int nrefs = r0 == r1 ? 1 : 2;
        best.r.cost = 0x7fffffff;
#pragma unroll 1
        for (int j = 0; j < nrefs; j++)
        {
            const int i_ref = j ? r1 : r0;
            MeSlot &m = c.w.slot;
            m.i_ref = i_ref; m.i_ref_cost = ref_cost(c, i_ref);
            m.i_pixel = dir ? PIX_8x16 : PIX_16x8;
            m.xoff = dir ? 8 * i : 0; m.yoff = dir ? 0 : 8 * i;
            int (*mvc)[2] = c.w.mvc;
            const int k1 = dir ? i + 1 : 2 * i + 1, k2 = dir ? i + 3 : 2 * i + 2;
            mvc[0][0] = a.mvc[i_ref][0][0]; mvc[0][1] = a.mvc[i_ref][0][1];
            mvc[1][0] = a.mvc[i_ref][k1][0]; mvc[1][1] = a.mvc[i_ref][k1][1];
            mvc[2][0] = a.mvc[i_ref][k2][0]; mvc[2][1] = a.mvc[i_ref][k2][1];
            if (dir) cache_fill_rect(c, 2 * i, 0, 2, 4, i_ref, 0, 1, 0);
            else     cache_fill_rect(c, 0, 2 * i, 4, 2, i_ref, 0, 1, 0);
            run_search<XS>(c, m, dir ? predict_mv(c, 4 * i, 2) : predict_mv(c, 8 * i, 4), mvc, 3, nullptr);
            m.r.cost += m.i_ref_cost;
            if (m.r.cost < best.r.cost)
                best = m;
        }
        const uint32_t mv = pack_mv(best.r.mv[0], best.r.mv[1]);
        if (dir) cache_fill_rect(c, 2 * i, 0, 2, 4, best.i_ref, mv, 1, 1);
        else     cache_fill_rect(c, 0, 2 * i, 4, 2, best.i_ref, mv, 1, 1);
        total += best.r.cost;
    }
    if (dir) a.cost8x16 = total; else a.cost16x8 = total;
}

// resumable forms of the three drivers above: same logic, loop positions and the search in flight kept in w.pt
// 16x16 search over all references; returns 1 when the early P_SKIP termination fired, 2 when a search is out (AS)
template <int XS, int AS>
PCAMV_FN int analyse_p16x16_rs(MbCtx &c, MbAnalysis &a, int allow_skip, int b_try_pskip)
{
    const int lambda = c.fc.tab.lambda;
    PtState &pt = c.w.pt;
    int &halfpel_thresh = c.w.halfpel_thresh;
    int *p_thresh = c.fp.n_ref > 1 ? &halfpel_thresh : nullptr;
    if (!pt.in16)
    {
        halfpel_thresh = 0x7fffffff;
        a.me16x16.r.cost = 0x7fffffff;
        pt.in16 = 1; pt.i_ref = 0; pt.wait = 0;
    }
#pragma unroll 1
    for (; pt.i_ref < c.fp.n_ref; pt.i_ref++)
    {
        const int i_ref = pt.i_ref;
        MeSlot &m = c.w.slot;
        const int rc = ref_cost(c, i_ref);
        if (!pt.wait)
        {
            halfpel_thresh -= rc;
            m.i_ref = i_ref; m.i_ref_cost = rc; m.i_pixel = PIX_16x16; m.xoff = 0; m.yoff = 0;
            int (*mvc)[2] = c.w.mvc;
            const uint32_t mvp = predict_mv_16x16(c, i_ref);
            const int i_mvc = predict_mv_ref16x16(c, i_ref, mvc);
            if (run_search_begin<XS, AS>(c, m, mvp, mvc, i_mvc, p_thresh)) { pt.wait = 1; return 2; }
        }
        pt.wait = 0;
        run_search_end<AS>(c, m, p_thresh);
        if (allow_skip && i_ref == 0 && b_try_pskip && m.r.cost - m.r.cost_mv < 300 * lambda &&
            iabs(m.r.mv[0] - c.pskip_mv[0]) + iabs(m.r.mv[1] - c.pskip_mv[1]) <= 1 && probe_pskip(c))
        {
            pt.in16 = 0;
            return 1;
        }
        m.r.cost += rc;
        halfpel_thresh += rc;
        if (m.r.cost < a.me16x16.r.cost)
            a.me16x16 = m;
        a.mvc[i_ref][0][0] = m.r.mv[0]; a.mvc[i_ref][0][1] = m.r.mv[1];
        if (team_lane() == 0)
            c.fp.cur.mvr[(size_t)i_ref * c.fc.mb_w * c.fc.mb_h + c.mb_xy] = pack_mv(m.r.mv[0], m.r.mv[1]);
    }
    pt.in16 = 0;
    cache_fill_rect(c, 0, 0, 4, 4, a.me16x16.i_ref, 0, 1, 0);
    return 0;
}

template <int XS, int AS>
PCAMV_FN int analyse_p8x8_rs(MbCtx &c, MbAnalysis &a)
{
    PtState &pt = c.w.pt;
    const int i_ref = a.me16x16.i_ref;
    const int rc = (c.fc.b_cabac || i_ref) ? ref_cost(c, i_ref) : 0;
    int (*mvc)[2] = a.mvc[i_ref];
    if (!pt.in8)
    {
        c.partition = PART_8x8;
        mvc[0][0] = a.me16x16.r.mv[0]; mvc[0][1] = a.me16x16.r.mv[1];
        pt.in8 = 1; pt.i = 0; pt.wait = 0;
    }
#pragma unroll 1
    for (; pt.i < 4; pt.i++)
    {
        const int i = pt.i;
        MeSlot &m = a.me8x8[i];
        const int x8 = i & 1, y8 = i >> 1;
        if (!pt.wait)
        {
            m.i_ref = i_ref; m.i_ref_cost = rc; m.i_pixel = PIX_8x8; m.xoff = 8 * x8; m.yoff = 8 * y8;
            if (run_search_begin<XS, AS>(c, m, predict_mv(c, 4 * i, 2), mvc, i + 1, nullptr)) { pt.wait = 1; return PT_YIELD; }
        }
        pt.wait = 0;
        run_search_end<AS>(c, m, nullptr);
        cache_fill_rect(c, 2 * x8, 2 * y8, 2, 2, 0, pack_mv(m.r.mv[0], m.r.mv[1]), 0, 1);
        mvc[i + 1][0] = m.r.mv[0]; mvc[i + 1][1] = m.r.mv[1];
        m.r.cost += rc;
        m.r.cost += c.fc.tab.lambda * 1;          // sub-partition type cost of an unsplit 8x8
    }
    pt.in8 = 0;
    a.cost8x8 = a.me8x8[0].r.cost + a.me8x8[1].r.cost + a.me8x8[2].r.cost + a.me8x8[3].r.cost;
    if (c.fc.b_cabac)
        a.cost8x8 -= rc;
    a.sub[0] = a.sub[1] = a.sub[2] = a.sub[3] = SUB_8x8;
    return PT_DONE;
}

// 16x8 (dir = 0) or 8x16 (dir = 1)
template <int XS, int AS>
PCAMV_FN int analyse_p16x8_8x16_rs(MbCtx &c, MbAnalysis &a, int dir)
{
    PtState &pt = c.w.pt;
    if (!pt.in168)
    {
        c.partition = dir ? PART_8x16 : PART_16x8;
        pt.total = 0;
        pt.in168 = 1; pt.i = 0; pt.j = 0; pt.wait = 0;
    }
#pragma unroll 1
    for (; pt.i < 2; pt.i++)
    {
        const int i = pt.i;
        MeSlot &best = dir ? a.me8x16[i] : a.me16x8[i];
        const int r0 = dir ? a.me8x8[i].i_ref : a.me8x8[2 * i].i_ref;
        const int r1 = dir ? a.me8x8[i + 2].i_ref : a.me8x8[2 * i + 1].i_ref;
        const int nrefs = r0 == r1 ? 1 : 2;
        if (pt.j == 0 && !pt.wait)
            best.r.cost = 0x7fffffff;
#pragma unroll 1
        for (; pt.j < nrefs; pt.j++)
        {
            const int i_ref = pt.j ? r1 : r0;
            MeSlot &m = c.w.slot;
            if (!pt.wait)
            {
                m.i_ref = i_ref; m.i_ref_cost = ref_cost(c, i_ref);
                m.i_pixel = dir ? PIX_8x16 : PIX_16x8;
                m.xoff = dir ? 8 * i : 0; m.yoff = dir ? 0 : 8 * i;
                int (*mvc)[2] = c.w.mvc;
                const int k1 = dir ? i + 1 : 2 * i + 1, k2 = dir ? i + 3 : 2 * i + 2;
                mvc[0][0] = a.mvc[i_ref][0][0]; mvc[0][1] = a.mvc[i_ref][0][1];
                mvc[1][0] = a.mvc[i_ref][k1][0]; mvc[1][1] = a.mvc[i_ref][k1][1];
                mvc[2][0] = a.mvc[i_ref][k2][0]; mvc[2][1] = a.mvc[i_ref][k2][1];
                if (dir) cache_fill_rect(c, 2 * i, 0, 2, 4, i_ref, 0, 1, 0);
                else     cache_fill_rect(c, 0, 2 * i, 4, 2, i_ref, 0, 1, 0);
                if (run_search_begin<XS, AS>(c, m, dir ? predict_mv(c, 4 * i, 2) : predict_mv(c, 8 * i, 4), mvc, 3, nullptr))
                {
                    pt.wait = 1;
                    return PT_YIELD;
                }
            }
            pt.wait = 0;
            run_search_end<AS>(c, m, nullptr);
            m.r.cost += m.i_ref_cost;
            if (m.r.cost < best.r.cost)
                best = m;
        }
        pt.j = 0;
        const uint32_t mv = pack_mv(best.r.mv[0], best.r.mv[1]);
        if (dir) cache_fill_rect(c, 2 * i, 0, 2, 4, best.i_ref, mv, 1, 1);
        else     cache_fill_rect(c, 0, 2 * i, 4, 2, best.i_ref, mv, 1, 1);
        pt.total += best.r.cost;
    }
    pt.in168 = 0;
    if (dir) a.cost8x16 = pt.total; else a.cost16x8 = pt.total;
    return PT_DONE;
}

// ---- sub-8x8 partitions of one 8x8 block (reference encoder/analyse.c:1569-1693) --------------------------------
// chroma part of a sub-partition cost (x264_mb_analyse_inter_p4x4_chroma, analyse.c:1535-1567): the 4x4 chroma block of
// each plane is predicted piecewise with the sub-blocks' own vectors (one pixel per lane) and compared with mbcmp 4x4
template <int XS>
PCAMV_FN int sub_chroma_cost(MbCtx &c, MbAnalysis &a, int i8, int kind)
{
    const DevRef &rf = c.fc.ref[c.fp.ref_slot[a.me8x8[i8].i_ref]];
    const SubSlot *ss = sub_slots(c.w, i8) + sub_first(kind);
    const int stride_c = c.fc.stride_c;
    const int cx0 = 4 * (i8 & 1), cy0 = 4 * (i8 >> 1);
    const ptrdiff_t offc = (ptrdiff_t)(8 * c.mb_y + cy0) * stride_c + 8 * c.mb_x + cx0;
    team_sync();
    PCAMV_FOR_ITEMS(it, 32)
    {
        const int pl = it >> 4, px_ = it & 3, py_ = (it >> 2) & 3;
        const int j = kind == SUB_4x4 ? (px_ >> 1) + 2 * (py_ >> 1) : kind == SUB_8x4 ? py_ >> 1 : px_ >> 1;
        const int mvx = ss[j].mv[0], mvy = ss[j].mv[1];
        const int dx = mvx & 7, dy = mvy & 7;
        const uint8_t *s = (pl ? rf.v : rf.u) + offc + (py_ + (mvy >> 3)) * stride_c + px_ + (mvx >> 3);
        const int v = ((8 - dx) * (8 - dy) * s[0] + dx * (8 - dy) * s[1] + (8 - dx) * dy * s[stride_c] + dx * dy * s[stride_c + 1] + 32) >> 6;
        (pl ? c.w.pred_v : c.w.pred_u)[(cy0 + py_) * 8 + cx0 + px_] = (uint8_t)v;
    }
    team_sync();
    int cost = 0;
    PCAMV_FOR_ITEMS(pl, 2)
    {
        const uint8_t *fe = (pl ? c.w.fenc_v : c.w.fenc_u) + cy0 * 8 + cx0, *pr = (pl ? c.w.pred_v : c.w.pred_u) + cy0 * 8 + cx0;
        uint32_t f[4], p4[4];
#pragma unroll
        for (int r = 0; r < 4; r++) { f[r] = ld4a(fe + 8 * r); p4[r] = ld4a(pr + 8 * r); }
        if (c.env.mbcmp_satd)
            cost += (int)(hadamard_4x4_sum(f, p4) >> 1);
        else
            cost += sad4(f[0], p4[0]) + sad4(f[1], p4[1]) + sad4(f[2], p4[2]) + sad4(f[3], p4[3]);
    }
    return team_sum(cost);
}

// searches of 8x8 block i8 split as `kind` (p4x4 / p8x4 / p4x8); returns the cost of that split
template <int XS>
PCAMV_FN int analyse_sub8x8(MbCtx &c, MbAnalysis &a, int i8, int kind)
{
    const int i_ref = a.me8x8[i8].i_ref;
    const int n = sub_count(kind);
    SubSlot *ss = sub_slots(c.w, i8) + sub_first(kind);
    c.partition = PART_8x8;
    int total = 0;
#pragma unroll 1
    for (int j = 0; j < n; j++)
    {
        // idx of the sub-block's first 4x4 in block_idx order: 4x4 -> j, 8x4 -> 2j, 4x8 -> j
        const int idx = 4 * i8 + (kind == SUB_8x4 ? 2 * j : j);
        const int x4 = (idx & 1) | ((idx >> 1) & 2), y4 = ((idx >> 1) & 1) | ((idx >> 2) & 2);
        MeSlot &m = c.w.slot;
        m.i_ref = i_ref; m.i_ref_cost = 0; m.i_pixel = sub_pixel(kind); m.xoff = 4 * x4; m.yoff = 4 * y4;
        int (*mvc)[2] = c.w.mvc;
        if (kind == SUB_4x4) { mvc[0][0] = a.me8x8[i8].r.mv[0]; mvc[0][1] = a.me8x8[i8].r.mv[1]; }
        else { const SubSlot &f = sub_slots(c.w, i8)[0]; mvc[0][0] = f.mv[0]; mvc[0][1] = f.mv[1]; }
        const uint32_t mvp = predict_mv(c, idx, kind == SUB_8x4 ? 2 : 1);
        run_search<XS>(c, m, mvp, mvc, j == 0, nullptr);
        ss[j].mv[0] = (int16_t)m.r.mv[0]; ss[j].mv[1] = (int16_t)m.r.mv[1];
        ss[j].mvp[0] = (int16_t)m.mvp[0]; ss[j].mvp[1] = (int16_t)m.mvp[1];
        ss[j].cost = m.r.cost; ss[j].cost_mv = m.r.cost_mv;
        cache_fill_rect(c, x4, y4, kind == SUB_4x8 || kind == SUB_4x4 ? 1 : 2, kind == SUB_8x4 || kind == SUB_4x4 ? 1 : 2, 0,
                        pack_mv(m.r.mv[0], m.r.mv[1]), 0, 1);
        total += m.r.cost;
    }
    // i_sub_mb_p_cost_table = { 5, 3, 3, 1 } for 4x4, 8x4, 4x8, 8x8 (analyse.c:179-181)
    total += ref_cost(c, i_ref) + c.fc.tab.lambda * (kind == SUB_4x4 ? 5 : 3);
    if (c.env.chroma_me)
        total += sub_chroma_cost<XS>(c, a, i8, kind);
    return total;
}

// x264_mb_cache_mv_p8x8 (analyse.c:1821-1849): the chosen split's vectors of 8x8 block i8 into the MV cache
PCAMV_FN void cache_mv_p8x8(MbCtx &c, const MbAnalysis &a, int i8)
{
    const int x = 2 * (i8 & 1), y = 2 * (i8 >> 1), kind = a.sub[i8];
    if (kind == SUB_8x8)
    {
        cache_fill_rect(c, x, y, 2, 2, 0, pack_mv(a.me8x8[i8].r.mv[0], a.me8x8[i8].r.mv[1]), 0, 1);
        return;
    }
    const SubSlot *ss = sub_slots(c.w, i8) + sub_first(kind);
#pragma unroll 1
    for (int j = 0; j < sub_count(kind); j++)
        cache_fill_rect(c, x + (sub_xoff(kind, j) >> 2), y + (sub_yoff(kind, j) >> 2), kind == SUB_8x4 ? 2 : 1, kind == SUB_4x8 ? 2 : 1, 0,
                        pack_mv(ss[j].mv[0], ss[j].mv[1]), 0, 1);
}

// x264_me_refine_qpel on the sub-blocks of a split 8x8 block; returns the summed cost
PCAMV_FN int refine_sub8x8(MbCtx &c, MbAnalysis &a, int i8)
{
    const int kind = a.sub[i8];
    if (kind == SUB_8x8)
    {
        run_refine(c, a.me8x8[i8]);
        return a.me8x8[i8].r.cost;
    }
    SubSlot *ss = sub_slots(c.w, i8) + sub_first(kind);
    int total = 0;
#pragma unroll 1
    for (int j = 0; j < sub_count(kind); j++)
    {
        MeSlot &m = c.w.slot;
        m.i_ref = a.me8x8[i8].i_ref; m.i_ref_cost = 0; m.i_pixel = sub_pixel(kind);
        m.xoff = 8 * (i8 & 1) + sub_xoff(kind, j); m.yoff = 8 * (i8 >> 1) + sub_yoff(kind, j);
        m.mvp[0] = ss[j].mvp[0]; m.mvp[1] = ss[j].mvp[1];
        m.r.mv[0] = ss[j].mv[0]; m.r.mv[1] = ss[j].mv[1]; m.r.cost = ss[j].cost; m.r.cost_mv = ss[j].cost_mv;
        run_refine(c, m);
        ss[j].mv[0] = (int16_t)m.r.mv[0]; ss[j].mv[1] = (int16_t)m.r.mv[1]; ss[j].cost = m.r.cost; ss[j].cost_mv = m.r.cost_mv;
        total += m.r.cost;
    }
    return total;
}

// write the decided mode into the neighbour cache (x264_analyse_update_cache, P part)
// SUB8: the sub-8x8 partition code is compiled in (X264_ANALYSE_PSUB8x8); kernels for the default partition set leave it out
template <int SUB8>
PCAMV_FN void update_cache(MbCtx &c, const MbAnalysis &a, int type, int partition)
{
    if (type == MB_P_SKIP)
    {
        cache_fill_rect(c, 0, 0, 4, 4, 0, pack_mv(c.pskip_mv[0], c.pskip_mv[1]), 1, 1);
        return;
    }
    if (partition == PART_16x16)
        cache_fill_rect(c, 0, 0, 4, 4, a.me16x16.i_ref, pack_mv(a.me16x16.r.mv[0], a.me16x16.r.mv[1]), 1, 1);
    else if (partition == PART_16x8)
#pragma unroll 1
        for (int i = 0; i < 2; i++)
            cache_fill_rect(c, 0, 2 * i, 4, 2, a.me16x8[i].i_ref, pack_mv(a.me16x8[i].r.mv[0], a.me16x8[i].r.mv[1]), 1, 1);
    else if (partition == PART_8x16)
#pragma unroll 1
        for (int i = 0; i < 2; i++)
            cache_fill_rect(c, 2 * i, 0, 2, 4, a.me8x16[i].i_ref, pack_mv(a.me8x16[i].r.mv[0], a.me8x16[i].r.mv[1]), 1, 1);
    else
#pragma unroll 1
        for (int i = 0; i < 4; i++)
        {
            if (SUB8)
            {
                cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, a.me8x8[i].i_ref, 0, 1, 0);
                cache_mv_p8x8(c, a, i);
            }
            else
                cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, a.me8x8[i].i_ref, pack_mv(a.me8x8[i].r.mv[0], a.me8x8[i].r.mv[1]), 1, 1);
        }
}

// store the MB's final state to the frame arrays (x264_macroblock_cache_save, inter part) and the result record
template <int SUB8>
PCAMV_FN void finalize_mb(MbCtx &c, const MbAnalysis &a, int type, int partition, int early_skip)
{
    const int no_parts = partition < 0;        // elided pass-2 macroblock: the partition slots were never searched
    if (no_parts) partition = -partition;
    const int mb_w = c.fc.mb_w, s8 = 2 * mb_w, s4 = 4 * mb_w;
    const int cur8 = 2 * c.mb_y * s8 + 2 * c.mb_x, cur4 = 4 * c.mb_y * s4 + 4 * c.mb_x;
    if (team_lane() == 0)
    {
        const FrameArrays &fa = c.fp.cur;
        fa.type[c.mb_xy] = (int8_t)type;
        fa.ref8[cur8] = c.w.ref[12]; fa.ref8[cur8 + 1] = c.w.ref[14];
        fa.ref8[cur8 + s8] = c.w.ref[28]; fa.ref8[cur8 + s8 + 1] = c.w.ref[30];
#pragma unroll 1
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++)
                fa.mv4[cur4 + y * s4 + x] = c.w.mv[12 + x + 8 * y];
        MbResult r = MbResult();      // unused partition slots stay zero: records are comparable byte for byte
        r.type = (int8_t)type; r.partition = (int8_t)partition; r.early_skip = (int8_t)early_skip;
        r.ref[0] = c.w.ref[12]; r.ref[1] = c.w.ref[14]; r.ref[2] = c.w.ref[28]; r.ref[3] = c.w.ref[30];
#pragma unroll 1
        for (int i = 0; i < 16; i++) r.mv[i] = c.w.mv[scan8(i)];
        r.n_part = 0;
        if (type != MB_P_SKIP && !no_parts && partition != PART_8x8)
        {
            const int np = partition == PART_16x16 ? 1 : 2;
            r.n_part = (int8_t)np;
#pragma unroll 1
            for (int i = 0; i < np; i++)
            {
                const MeSlot &m = partition == PART_16x16 ? a.me16x16 : partition == PART_16x8 ? a.me16x8[i] : a.me8x16[i];
                r.part[i].mv[0] = (int16_t)m.r.mv[0]; r.part[i].mv[1] = (int16_t)m.r.mv[1];
                r.part[i].mvp[0] = (int16_t)m.mvp[0]; r.part[i].mvp[1] = (int16_t)m.mvp[1];
                r.part[i].ref = (int8_t)m.i_ref; r.part[i].i_pixel = (int8_t)m.i_pixel;
                r.part[i].xoff = (int8_t)m.xoff; r.part[i].yoff = (int8_t)m.yoff;
            }
        }
        else if (SUB8 && type != MB_P_SKIP && !no_parts)
        {
            // P_8x8: up to 16 MV-carrying blocks, in the order the reference costs them (analyse.c:3546-3606); they go to the
            // side array (the first four are mirrored into the record)
            int np = 0;
            PartInfo *sp = c.fp.subparts ? c.fp.subparts + (size_t)16 * c.mb_xy : nullptr;
#pragma unroll 1
            for (int i = 0; i < 4; i++)
            {
                const int kind = a.sub[i];
                const SubSlot *ss = sub_slots(c.w, i) + sub_first(kind);
#pragma unroll 1
                for (int j = 0; j < sub_count(kind); j++)
                {
                    PartInfo pi;
                    if (kind == SUB_8x8)
                    {
                        const MeSlot &m = a.me8x8[i];
                        pi.mv[0] = (int16_t)m.r.mv[0]; pi.mv[1] = (int16_t)m.r.mv[1];
                        pi.mvp[0] = (int16_t)m.mvp[0]; pi.mvp[1] = (int16_t)m.mvp[1];
                    }
                    else
                    {
                        pi.mv[0] = ss[j].mv[0]; pi.mv[1] = ss[j].mv[1]; pi.mvp[0] = ss[j].mvp[0]; pi.mvp[1] = ss[j].mvp[1];
                    }
                    pi.ref = (int8_t)a.me8x8[i].i_ref; pi.i_pixel = (int8_t)sub_pixel(kind);
                    pi.xoff = (int8_t)(8 * (i & 1) + sub_xoff(kind, j)); pi.yoff = (int8_t)(8 * (i >> 1) + sub_yoff(kind, j));
                    if (np < 4) r.part[np] = pi;
                    if (sp) sp[np] = pi;
                    np++;
                }
            }
            r.n_part = (int8_t)np;
        }
        r.n_log = c.n_log;
        r.pskip_mv[0] = (int16_t)c.pskip_mv[0]; r.pskip_mv[1] = (int16_t)c.pskip_mv[1];
        c.fp.results[c.mb_xy] = r;
    }
}

// Quirk q2 support: the wavefront only orders a macroblock after its left / top / top-right neighbours, but the
// stale cache of a forced skip is what the previous MB IN RASTER ORDER left behind; for mb_x == 0 that is the last
// MB of the row above, so wait for that whole row (it never depends on this one, hence no deadlock).
PCAMV_DEV void wait_prev_raster(const MbCtx &c)
{
#if !defined(PCAMV_EMU)
    if (c.mb_x == 0 && c.mb_y > 0)
    {
        const int *flag = c.fp.row_progress + c.mb_y - 1;
        int v;
        unsigned ns = 64;
        for (;;)
        {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v >= c.fc.mb_w) break;
            __nanosleep(ns);                       // back off: the row above still has macroblocks to go
            if (ns < 4096) ns <<= 1;
        }
    }
#else
    (void)c;
#endif
}

// One macroblock of a P slice.  `prev_mv` = the 16 cache MVs left behind by the previous MB in raster order
// (needed only for the pass-2 "forced skip without cache update" quirk, analyse.c:2668-2676).
// F: feature mask of the instantiation — bit 0 = exhaustive searches (--me esa / tesa), bit 1 = sub-8x8 partitions
template <int F>
PCAMV_FN void analyse_p_mb(MbCtx &c, const uint32_t *prev_mv)
{
    constexpr int XS = F & 1, SUB8 = (F >> 1) & 1;
    const DevFrameCtx &fc = c.fc;
    MbAnalysis &a = c.w.an;
    c.n_log = 0;
    c.partition = PART_16x16;
    cache_load(c);
    const uint32_t ps = predict_mv_pskip(c);
    c.pskip_mv[0] = mv_x(ps); c.pskip_mv[1] = mv_y(ps);
    init_limits(c);
    c.env.cost_mv = fc.tab.cost_mv;
    c.env.me_method = fc.me_method; c.env.me_range = fc.me_range; c.env.subme = fc.subme;
    c.env.chroma_me = fc.chroma_me && fc.subme >= 5;
    c.env.mbcmp_satd = fc.subme > 1;
    c.env.mvsads = c.fp.mvsads ? c.fp.mvsads + (size_t)c.mb_y * c.fp.mvsads_cap : nullptr;

    int b_try_pskip = 0, b_skip = 0;
    if (fc.b_fast_pskip)
    {
        if (fc.subme >= 3) b_try_pskip = 1;
        else if (c.type_left == MB_P_SKIP || c.type_top == MB_P_SKIP || c.type_topleft == MB_P_SKIP || c.type_topright == MB_P_SKIP)
            b_skip = probe_pskip(c);
    }
    const ForcedMb *forced = c.fp.pass == 2 ? &c.fp.forced[c.mb_xy] : nullptr;
    int type = MB_P_L0, partition = PART_16x16, early_skip = 0;
    if (b_skip)
    {
        // (subme < 3 only) the reference takes this MB as P_SKIP before any search; pass 2 cannot override it
        update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);
        finalize_mb<SUB8>(c, a, MB_P_SKIP, PART_16x16, 1);
        return;
    }
    early_skip = analyse_p16x16<XS>(c, a, 1, b_try_pskip);
    if (early_skip)
    {
        type = MB_P_SKIP;
        update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);
    }
    if (forced)
    {
        if (forced->type != MB_P_SKIP && type == MB_P_SKIP)
            analyse_p16x16<XS>(c, a, 0, b_try_pskip);       // the reference re-runs the 16x16 search without the skip exit
        type = forced->type;
    }
    if (type == MB_P_SKIP)
    {
        if (!early_skip && fc.conformant)
            update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);    // pcamv_set_conformant: a forced skip leaves like every other skip
        else if (!early_skip)
        {
            // forced to P_SKIP without x264_analyse_update_cache: the MV cache still holds the previous MB's vectors
            // and the refs are what the 16x16 search left (its best reference)
            wait_prev_raster(c);
#pragma unroll 1
            for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = PCAMV_LDV(prev_mv + i);
        }
        finalize_mb<SUB8>(c, a, MB_P_SKIP, PART_16x16, early_skip);
        return;
    }

    // Elision has one exception.  When this pass's probe found the macroblock skippable but pass 1 did not (the re-run above),
    // the host keeps b_skip_mc set (quirk q1, analyse.c:2663-2668 / encoder/macroblock.c:611-612): x264_macroblock_encode then
    // takes the residual against whatever the analysis left in fdec, and that is the host's INTRA analysis, whose early-outs
    // compare against the inter cost of the "dead" searches and refinement.  Such macroblocks get the full analysis.
    if (forced && forced->used && fc.pass2_elide && !early_skip)
    {
        // pass 2, decision forced from pass 1: nothing the remaining searches produce survives analyse.c:2868-2991
        // (info.cache[].i_partition is only written for P_L0, analyse.c:3612: a forced P_8x8 is 8x8 by construction)
        type = forced->type; partition = forced->type == MB_P_8x8 ? PART_8x8 : forced->partition;
#pragma unroll 1
        for (int i = 0; i < 4; i++)
            cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, forced->ref[i], 0, 1, 0);
#pragma unroll 1
        for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = forced->mv[i];
        finalize_mb<SUB8>(c, a, type, -partition, 0);
        return;
    }

    const int flags = fc.analyse_inter;
    const int psub16 = (flags & 0x10) != 0;
    if (psub16)
        analyse_p8x8<XS>(c, a);
    int i_cost = a.me16x16.r.cost;
    if (SUB8 && psub16 && (flags & 0x20) && a.cost8x8 < a.me16x16.r.cost)
    {
        // X264_ANALYSE_PSUB8x8: P_8x8 becomes the incumbent and every 8x8 block may split further (analyse.c:2697-2731)
        type = MB_P_8x8; partition = PART_8x8;
        i_cost = a.cost8x8;
#pragma unroll 1
        for (int i = 0; i < 4; i++)
        {
            const int c4x4 = analyse_sub8x8<XS>(c, a, i, SUB_4x4);
            if (c4x4 < a.me8x8[i].r.cost)
            {
                int best8 = c4x4;
                a.sub[i] = SUB_4x4;
                const int c8x4 = analyse_sub8x8<XS>(c, a, i, SUB_8x4);
                if (c8x4 < best8) { best8 = c8x4; a.sub[i] = SUB_8x4; }
                const int c4x8 = analyse_sub8x8<XS>(c, a, i, SUB_4x8);
                if (c4x8 < best8) { best8 = c4x8; a.sub[i] = SUB_4x8; }
                i_cost += best8 - a.me8x8[i].r.cost;
            }
            cache_mv_p8x8(c, a, i);
        }
        a.cost8x8 = i_cost;
    }
    if (psub16)
    {
        const int thresh16x8 = a.me8x8[1].r.cost_mv + a.me8x8[2].r.cost_mv;
        if (a.cost8x8 < a.me16x16.r.cost + thresh16x8)
        {
            analyse_p16x8_8x16<XS>(c, a, 0);
            if (a.cost16x8 < i_cost) { i_cost = a.cost16x8; type = MB_P_L0; partition = PART_16x8; }
            analyse_p16x8_8x16<XS>(c, a, 1);
            if (a.cost8x16 < i_cost) { i_cost = a.cost8x16; type = MB_P_L0; partition = PART_8x16; }
        }
    }
    c.partition = partition;
    if (partition == PART_16x16) run_refine(c, a.me16x16);
    else if (partition == PART_16x8) { run_refine(c, a.me16x8[0]); run_refine(c, a.me16x8[1]); }
    else if (partition == PART_8x16) { run_refine(c, a.me8x16[0]); run_refine(c, a.me8x16[1]); }
    else if (SUB8)
#pragma unroll 1
        for (int i = 0; i < 4; i++) refine_sub8x8(c, a, i);

    if (forced && forced->used)
    {
        // pass 2: type / partition / refs / MVs come from pass 1 with the embedding flips applied; for P_8x8 the reference
        // forces the sub-partition types and leaves h->mb.i_partition as this pass decided it (analyse.c:2872-2890)
        type = forced->type;
        if (type != MB_P_8x8) partition = forced->partition;
        else if (fc.conformant) partition = PART_8x8;       // pcamv_set_conformant: a forced P_8x8 is 8x8 in h->mb.i_partition too
#pragma unroll 1
        for (int i = 0; i < 4; i++)
            cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, forced->ref[i], 0, 1, 0);
#pragma unroll 1
        for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = forced->mv[i];
        // the analysis slots keep the searched values; neighbours only ever see the cache
    }
    else
        update_cache<SUB8>(c, a, type, partition);
    // a forced decision names partitions this pass may never have searched: its record carries no partition slots
    // (pass 2 has no cost table; the slots would be whatever the team's scratch held)
    // early_skip = 2: this pass's probe found the macroblock skippable, the decision forced from pass 1 codes it anyway (quirk q1:
    // the host keeps b_skip_mc set and reconstructs it from what its intra analysis left in fdec — pcamv_recon.cuh needs to know)
    finalize_mb<SUB8>(c, a, type, (forced && forced->used) ? -partition : partition, early_skip ? 2 : 0);
}

// Resumable form of analyse_p_mb (same decisions, written as stages that can be left and re-entered).
// AS = 1 (split wavefront): every search is handed out — the function returns PT_YIELD with the request in w.rq and is
// called again, with w.pt and everything before MbWork::fenc_y as it left them and the result in w.rs, until it returns
// PT_DONE.  The caller zeroes w.pt.stage before the first call of a macroblock.  AS = 0 never yields.
template <int F, int AS>
PCAMV_FN int analyse_p_mb_rs(MbCtx &c, const uint32_t *prev_mv)
{
    constexpr int XS = F & 1, SUB8 = (F >> 1) & 1;
    static_assert(!(AS && SUB8), "the sub-8x8 partition searches are synchronous");
    const DevFrameCtx &fc = c.fc;
    MbAnalysis &a = c.w.an;
    PtState &pt = c.w.pt;
    const ForcedMb *forced = c.fp.pass == 2 ? &c.fp.forced[c.mb_xy] : nullptr;
    if (pt.stage == 0)
    {
        c.n_log = 0;
        c.partition = PART_16x16;
        cache_load(c);
        const uint32_t ps = predict_mv_pskip(c);
        c.pskip_mv[0] = mv_x(ps); c.pskip_mv[1] = mv_y(ps);
        init_limits(c);
        c.env.cost_mv = fc.tab.cost_mv;
        c.env.me_method = fc.me_method; c.env.me_range = fc.me_range; c.env.subme = fc.subme;
        c.env.chroma_me = fc.chroma_me && fc.subme >= 5;
        c.env.mbcmp_satd = fc.subme > 1;
        c.env.mvsads = c.fp.mvsads ? c.fp.mvsads + (size_t)c.mb_y * c.fp.mvsads_cap : nullptr;

        int b_try_pskip = 0, b_skip = 0;
        if (fc.b_fast_pskip)
        {
            if (fc.subme >= 3) b_try_pskip = 1;
            else if (c.type_left == MB_P_SKIP || c.type_top == MB_P_SKIP || c.type_topleft == MB_P_SKIP || c.type_topright == MB_P_SKIP)
                b_skip = probe_pskip(c);
        }
        pt.b_try_pskip = (int8_t)b_try_pskip; pt.type = MB_P_L0; pt.partition = PART_16x16; pt.early_skip = 0;
        pt.in16 = pt.in8 = pt.in168 = pt.wait = 0;
        if (b_skip)
        {
            // (subme < 3 only) the reference takes this MB as P_SKIP before any search; pass 2 cannot override it
            update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);
            finalize_mb<SUB8>(c, a, MB_P_SKIP, PART_16x16, 1);
            return PT_DONE;
        }
        pt.stage = 1;
    }
    if (pt.stage == 1)
    {
        const int r = analyse_p16x16_rs<XS, AS>(c, a, 1, pt.b_try_pskip);
        if (r == 2) return PT_YIELD;
        pt.early_skip = (int8_t)r;
        if (r)
        {
            pt.type = MB_P_SKIP;
            update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);
        }
        // the reference re-runs the 16x16 search without the skip exit when pass 1 coded a macroblock this pass would skip
        pt.stage = (forced && forced->type != MB_P_SKIP && pt.type == MB_P_SKIP) ? 2 : 3;
    }
    if (pt.stage == 2)
    {
        if (analyse_p16x16_rs<XS, AS>(c, a, 0, pt.b_try_pskip) == 2) return PT_YIELD;
        pt.stage = 3;
    }
    if (pt.stage == 3)
    {
        const int early_skip = pt.early_skip;
        if (forced)
            pt.type = forced->type;
        if (pt.type == MB_P_SKIP)
        {
            if (!early_skip && fc.conformant)
                update_cache<SUB8>(c, a, MB_P_SKIP, PART_16x16);
            else if (!early_skip)
            {
                // forced to P_SKIP without x264_analyse_update_cache: the MV cache still holds the previous MB's vectors
                // and the refs are what the 16x16 search left (its best reference)
                wait_prev_raster(c);
#pragma unroll 1
                for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = PCAMV_LDV(prev_mv + i);
            }
            finalize_mb<SUB8>(c, a, MB_P_SKIP, PART_16x16, early_skip);
            return PT_DONE;
        }

        // Elision has one exception.  When this pass's probe found the macroblock skippable but pass 1 did not (the re-run above),
        // the host keeps b_skip_mc set (quirk q1, analyse.c:2663-2668 / encoder/macroblock.c:611-612): x264_macroblock_encode then
        // takes the residual against whatever the analysis left in fdec, and that is the host's INTRA analysis, whose early-outs
        // compare against the inter cost of the "dead" searches and refinement.  Such macroblocks get the full analysis.
        if (forced && forced->used && fc.pass2_elide && !early_skip)
        {
            // pass 2, decision forced from pass 1: nothing the remaining searches produce survives analyse.c:2868-2991
            // (info.cache[].i_partition is only written for P_L0, analyse.c:3612: a forced P_8x8 is 8x8 by construction)
            const int type = forced->type, partition = forced->type == MB_P_8x8 ? PART_8x8 : forced->partition;
#pragma unroll 1
            for (int i = 0; i < 4; i++)
                cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, forced->ref[i], 0, 1, 0);
#pragma unroll 1
            for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = forced->mv[i];
            finalize_mb<SUB8>(c, a, type, -partition, 0);
            return PT_DONE;
        }
        pt.type = MB_P_L0; pt.partition = PART_16x16;
        pt.stage = 4;
    }

    const int flags = fc.analyse_inter;
    const int psub16 = (flags & 0x10) != 0;
    if (pt.stage == 4)
    {
        if (psub16 && analyse_p8x8_rs<XS, AS>(c, a) == PT_YIELD)
            return PT_YIELD;
        pt.i_cost = a.me16x16.r.cost;
        if (SUB8 && psub16 && (flags & 0x20) && a.cost8x8 < a.me16x16.r.cost)
        {
            // X264_ANALYSE_PSUB8x8: P_8x8 becomes the incumbent and every 8x8 block may split further (analyse.c:2697-2731)
            pt.type = MB_P_8x8; pt.partition = PART_8x8;
            int i_cost = a.cost8x8;
#pragma unroll 1
            for (int i = 0; i < 4; i++)
            {
                const int c4x4 = analyse_sub8x8<XS>(c, a, i, SUB_4x4);
                if (c4x4 < a.me8x8[i].r.cost)
                {
                    int best8 = c4x4;
                    a.sub[i] = SUB_4x4;
                    const int c8x4 = analyse_sub8x8<XS>(c, a, i, SUB_8x4);
                    if (c8x4 < best8) { best8 = c8x4; a.sub[i] = SUB_8x4; }
                    const int c4x8 = analyse_sub8x8<XS>(c, a, i, SUB_4x8);
                    if (c4x8 < best8) { best8 = c4x8; a.sub[i] = SUB_4x8; }
                    i_cost += best8 - a.me8x8[i].r.cost;
                }
                cache_mv_p8x8(c, a, i);
            }
            a.cost8x8 = i_cost;
            pt.i_cost = i_cost;
        }
        pt.stage = 7;
        if (psub16)
        {
            const int thresh16x8 = a.me8x8[1].r.cost_mv + a.me8x8[2].r.cost_mv;
            if (a.cost8x8 < a.me16x16.r.cost + thresh16x8)
                pt.stage = 5;
        }
    }
    if (pt.stage == 5)
    {
        if (analyse_p16x8_8x16_rs<XS, AS>(c, a, 0) == PT_YIELD) return PT_YIELD;
        if (a.cost16x8 < pt.i_cost) { pt.i_cost = a.cost16x8; pt.type = MB_P_L0; pt.partition = PART_16x8; }
        pt.stage = 6;
    }
    if (pt.stage == 6)
    {
        if (analyse_p16x8_8x16_rs<XS, AS>(c, a, 1) == PT_YIELD) return PT_YIELD;
        if (a.cost8x16 < pt.i_cost) { pt.i_cost = a.cost8x16; pt.type = MB_P_L0; pt.partition = PART_8x16; }
        pt.stage = 7;
    }
    if (pt.stage == 7)
    {
        c.partition = pt.partition;
        pt.i = 0; pt.wait = 0;
        pt.stage = 8;
    }
    if (pt.stage == 8)
    {
        // quarter-pel refinement of the winner's partitions
        if (pt.partition != PART_8x8)
        {
            const int np = pt.partition == PART_16x16 ? 1 : 2;
#pragma unroll 1
            for (; pt.i < np; pt.i++)
            {
                MeSlot &m = pt.partition == PART_16x16 ? a.me16x16 : pt.partition == PART_16x8 ? a.me16x8[pt.i] : a.me8x16[pt.i];
                if (!pt.wait && run_refine_begin<AS>(c, m)) { pt.wait = 1; return PT_YIELD; }
                pt.wait = 0;
                run_refine_end<AS>(c, m);
            }
        }
        else if (SUB8)
#pragma unroll 1
            for (int i = 0; i < 4; i++) refine_sub8x8(c, a, i);
        pt.stage = 9;
    }

    int type = pt.type, partition = pt.partition;
    if (forced && forced->used)
    {
        // pass 2: type / partition / refs / MVs come from pass 1 with the embedding flips applied; for P_8x8 the reference
        // forces the sub-partition types and leaves h->mb.i_partition as this pass decided it (analyse.c:2872-2890)
        type = forced->type;
        if (type != MB_P_8x8) partition = forced->partition;
        else if (fc.conformant) partition = PART_8x8;       // pcamv_set_conformant: a forced P_8x8 is 8x8 in h->mb.i_partition too
#pragma unroll 1
        for (int i = 0; i < 4; i++)
            cache_fill_rect(c, 2 * (i & 1), 2 * (i >> 1), 2, 2, forced->ref[i], 0, 1, 0);
#pragma unroll 1
        for (int i = 0; i < 16; i++) c.w.mv[scan8(i)] = forced->mv[i];
        // the analysis slots keep the searched values; neighbours only ever see the cache
    }
    else
        update_cache<SUB8>(c, a, type, partition);
    // a forced decision names partitions this pass may never have searched: its record carries no partition slots
    // (pass 2 has no cost table; the slots would be whatever the team's scratch held)
    finalize_mb<SUB8>(c, a, type, (forced && forced->used) ? -partition : partition, pt.early_skip ? 2 : 0);
    return PT_DONE;
}
} // namespace pcamv
