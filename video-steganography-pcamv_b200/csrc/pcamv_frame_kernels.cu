// pcamv_frame_kernels.cu — frame-level kernels of the P-slice analysis.
//
//   k_analyse_p     macroblock wavefront over one P slice: neighbour cache, MV prediction, P_SKIP probe,
//                   16x16 / 8x8 / 16x8 / 8x16 searches, mode decision, qpel refinement, pass-2 forcing
//                   (reference encoder/analyse.c:2613-3172 non-RD P path; device code in pcamv_frame.cuh)
//   k_cost_table    the PCAMV candidate-MV cost table of every non-skip macroblock
//                   (reference encoder/analyse.c:2391-2550, 3518-3689; device code in pcamv_cost.cuh)
//
// Wavefront: a macroblock needs its left, top-left, top and top-right neighbours (MV prediction,
// common/macroblock.c:28-163,388-470,1085-1224).  One lane team (warp) owns one macroblock ROW and walks it
// left to right; before macroblock x it waits until the row above has finished x+1 (its top-right).  Groups of
// AP_WARPS consecutive rows are claimed by a CTA from an atomic counter in increasing order, so every row a claimed
// row waits on is owned by a team that is already running: the scheme cannot deadlock for any grid size.  Progress is published with
// st.release.gpu after the macroblock's state is in HBM and consumed with ld.acquire.gpu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <new>
#include "pcamv_device.h"
#include "pcamv_cost.cuh"

namespace pcamv {

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// source pixels of one macroblock -> team scratch (Y 16x16 stride 16, U/V 8x8 stride 8)
__device__ __forceinline__ void stage_fenc(const DevFrameCtx &fc, int mb_x, int mb_y, MbWork &w)
{
    const int lane = threadIdx.x & 31;
    if (lane < 16)
        *(uint4 *)(w.fenc_y + 16 * lane) = *(const uint4 *)(fc.fenc_y + (size_t)(16 * mb_y + lane) * fc.stride_y + 16 * mb_x);
    else if (lane < 24)
        *(uint2 *)(w.fenc_u + 8 * (lane - 16)) = *(const uint2 *)(fc.fenc_u + (size_t)(8 * mb_y + lane - 16) * fc.stride_c + 8 * mb_x);
    else
        *(uint2 *)(w.fenc_v + 8 * (lane - 24)) = *(const uint2 *)(fc.fenc_v + (size_t)(8 * mb_y + lane - 24) * fc.stride_c + 8 * mb_x);
    __syncwarp();
}


// one macroblock row, left to right, behind the row above (the caller has claimed `row` in increasing order)
template <int XS>
__device__ __forceinline__ void analyse_row(const DevFrameCtx &fc, const FrameParams &fp, MbCtx &c, MbWork &work, int row)
{
    const int lane = threadIdx.x & 31;
    const int mb_w = fc.mb_w;
    for (int x = 0; x < mb_w; x++)
    {
        if (row > 0)
        {
            const int need = min(x + 2, mb_w);
            // back off while waiting: with many encoder contexts resident, spinning warps would otherwise take
            // issue slots from the ones doing the work
            // (a macroblock takes ~150 us, so polling every few microseconds costs no latency worth the name, while
            // tight polling by the waiting half of the resident warps was measured at two thirds of all issued instructions)
#ifndef PCAMV_POLL_NS_MIN
#define PCAMV_POLL_NS_MIN 64
#define PCAMV_POLL_NS_MAX 4096
#endif
            unsigned ns = PCAMV_POLL_NS_MIN;
            while (ld_acquire(fp.row_progress + row - 1) < need)
            {
                __nanosleep(ns);
                if (ns < PCAMV_POLL_NS_MAX) ns <<= 1;
            }
        }
        c.mb_x = x; c.mb_y = row; c.mb_xy = row * mb_w + x;
        if (fp.trace && lane == 0) fp.trace[2 * c.mb_xy] = globaltimer_ns();
        stage_fenc(fc, x, row, work);
        analyse_p_mb<XS>(c, c.mb_xy ? fp.results[c.mb_xy - 1].mv : fp.stale_mv);
        __syncwarp();
        if (lane == 0)
        {
            if (fp.trace) fp.trace[2 * c.mb_xy + 1] = globaltimer_ns();
            __threadfence();
            st_release(fp.row_progress + row, x + 1);
        }
    }
}

template <int AP_WARPS, int XS>
__global__ void __launch_bounds__(AP_WARPS * 32) k_analyse_p(const __grid_constant__ DevFrameCtx fc,
                                                            const __grid_constant__ FrameParams fp, int *row_claim)
{
    __shared__ MbWork s_work[AP_WARPS];
    __shared__ __align__(16) unsigned char s_ctx[AP_WARPS][sizeof(MbCtx)];
    MbWork &work = s_work[threadIdx.x >> 5];
    MbCtx &c = *new (s_ctx[threadIdx.x >> 5]) MbCtx(fc, fp, work);      // every lane writes the same values
    const int warp = threadIdx.x >> 5;
    const int mb_h = fc.mb_h;
    __shared__ int s_group;
    for (;;)
    {
        // a CTA takes AP_WARPS consecutive rows: their reference windows overlap by 3/4 vertically and the rows run two
        // macroblocks apart, so the warps of a CTA share most of their L1 lines
        __syncthreads();
        if (threadIdx.x == 0)
            s_group = atomicAdd(row_claim, 1);
        __syncthreads();
        const int row = s_group * AP_WARPS + warp;
        if (s_group * AP_WARPS >= mb_h)
            return;
        if (row < mb_h)
            analyse_row<XS>(fc, fp, c, work, row);
    }
}

// Several frames (independent encoder contexts: GOP shards / streams of equal geometry) in ONE launch.  Work unit = a
// group of AP_WARPS consecutive rows of one frame.  A persistent CTA looking for work scans the frames (each thread
// looks at one, starting from a CTA-specific offset) for one whose NEXT group is ready — it is the first group, or the
// last row of the previous group is already a few macroblocks in — and claims it with a compare-and-swap on that
// frame's counter.  So a claimed group starts working at once instead of spinning through the wavefront's ramp (with
// a static assignment about half of the resident warps were waiting at any time), the resident warps are spread
// over as many frames as it takes to keep them busy, and every awaited row belongs to a group that was claimed before
// the waiter's — the no-deadlock argument of the single-frame kernel, for any grid size.
#ifndef PCAMV_BATCH_MIN_CTAS
#define PCAMV_BATCH_MIN_CTAS 6      // CTAs of 4 warps per SM the register budget must allow (resident warps are what hides latency here)
#endif
#define PCAMV_GROUP_READY 4      // macroblocks the previous group's last row must have finished before the next group is handed out
template <int AP_WARPS, int XS>
__global__ void __launch_bounds__(AP_WARPS * 32, PCAMV_BATCH_MIN_CTAS * 4 / AP_WARPS) k_analyse_p_batch(const BatchItem *__restrict__ items, int n_items, int *next_group)
{
    __shared__ MbWork s_work[AP_WARPS];
    __shared__ __align__(16) unsigned char s_ctx[AP_WARPS][sizeof(MbCtx)];
    __shared__ int s_pick, s_left, s_frame, s_grp;
    const int warp = threadIdx.x >> 5;
    MbWork &work = s_work[warp];
    const int mb_h = items[0].fc.mb_h;
    const int n_groups = (mb_h + AP_WARPS - 1) / AP_WARPS;
    unsigned rot = blockIdx.x * 37u;
    for (;;)
    {
        __syncthreads();
        if (threadIdx.x == 0) { s_pick = 0x7fffffff; s_left = 0; s_frame = -1; }
        __syncthreads();
        // every thread inspects frames rot + t, rot + t + blockDim, ...: lowest ready candidate wins
        int mine = 0x7fffffff, left = 0;
        for (int t = threadIdx.x; t < n_items; t += blockDim.x)
        {
            const int f = (int)((rot + (unsigned)t) % (unsigned)n_items);
            const int g = ld_acquire(next_group + f);
            if (g >= n_groups) continue;
            left = 1;
            if (mine == 0x7fffffff &&
                (g == 0 || ld_acquire(items[f].fp.row_progress + g * AP_WARPS - 1) >= min(PCAMV_GROUP_READY, items[f].fc.mb_w)))
                mine = t;
        }
        if (left) s_left = 1;                       // (benign race: everybody writes 1)
        if (mine != 0x7fffffff) atomicMin(&s_pick, mine);
        __syncthreads();
        if (!s_left)
            return;                                 // every group of every frame has been handed out
        if (threadIdx.x == 0 && s_pick != 0x7fffffff)
        {
            const int f = (int)((rot + (unsigned)s_pick) % (unsigned)n_items);
            const int g = ld_acquire(next_group + f);
            if (g < n_groups && atomicCAS(next_group + f, g, g + 1) == g) { s_frame = f; s_grp = g; }
        }
        __syncthreads();
        rot += 1u;
        if (s_frame < 0)
        {
            if (s_pick == 0x7fffffff) __nanosleep(2000);      // nothing ready anywhere: the running groups will get there
            continue;
        }
        const int row = s_grp * AP_WARPS + warp;
        if (row < mb_h)
        {
            const BatchItem &it = items[s_frame];
            MbCtx &c = *new (s_ctx[warp]) MbCtx(it.fc, it.fp, work);
            analyse_row<XS>(it.fc, it.fp, c, work, row);
        }
    }
}

// ---- row pool: the multi-context wavefront without standing waits ---------------------------------------------------
// In the row-group kernel above a team that reaches a macroblock whose top-right neighbour is not finished sleeps until it
// is; with minimal lag between consecutive rows that is the common case, and the profile of the 128-context launch shows
// ~45 % of the resident warp time asleep or polling.  Here a ROW is a resumable task, not a team's property: a team
// claims any row whose next macroblock is ready (compare-and-swap on the row's owner word), analyses macroblocks while
// they stay ready, then publishes the row's progress, releases it and looks for the next one — a blocked row waits in
// memory, not in a warp.  Nothing the analysis of a macroblock needs lives in the team between macroblocks (MbCtx / MbWork
// are rebuilt from the frame's records), so any team can resume any row.  Progress: the topmost unfinished row of a frame
// never waits on anything, so some row is always claimable until all are done — for any grid size.
// The pass-2 "forced skip" quirk (q2, analyse.c:2668-2676: the stale MV cache of the previous macroblock in raster
// order) makes macroblock 0 of a row depend on the whole previous row; the readiness test asks for that whenever the
// pass-1 record says P_SKIP, so wait_prev_raster() finds its condition already true and never spins here.
__device__ __forceinline__ bool row_ready(const DevFrameCtx &fc, const FrameParams &fp, int row, int x)
{
    if (row == 0)
        return true;
    int need = min(x + 2, fc.mb_w);
    if (x == 0 && fp.pass == 2 && fp.forced[row * fc.mb_w].type == MB_P_SKIP)
        need = fc.mb_w;
    return ld_acquire(fp.row_progress + row - 1) >= need;
}

template <int F>
__global__ void __launch_bounds__(128, PCAMV_BATCH_MIN_CTAS) k_analyse_p_pool(const BatchItem *__restrict__ items, int n_items, int *rows_done)
{
    __shared__ MbWork s_work[4];
    __shared__ __align__(16) unsigned char s_ctx[4][sizeof(MbCtx)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    MbWork &work = s_work[warp];
    const int mb_h = items[0].fc.mb_h, mb_w = items[0].fc.mb_w;
    const int total = n_items * mb_h;
    // teams start their searches spread over the frames; a team stays near the row it last ran (its reference window is warm)
    unsigned pos = (unsigned)(((unsigned long long)(blockIdx.x * 4 + warp) * (unsigned)total) / (gridDim.x * 4u));
    unsigned backoff = 256;
    for (;;)
    {
        if (ld_acquire(rows_done) >= total)
            return;
        // one sweep over the rows, 32 per step, from `pos`: first claimable row wins
        int got = -1;
        for (int step = 0; step * 32 < total && got < 0; step++)
        {
            const int g = (int)((pos + (unsigned)(step * 32 + lane)) % (unsigned)total);
            const int f = g / mb_h, r = g - f * mb_h;
            const FrameParams &fp = items[f].fp;
            const int *owner = fp.row_progress + mb_h + 1;
            bool ok = false;
            if (step * 32 + lane < total)
            {
                const int x = ld_acquire(fp.row_progress + r);
                ok = x < mb_w && ld_acquire(owner + r) == 0 && row_ready(items[f].fc, fp, r, x);
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            while (m && got < 0)
            {
                const int l = __ffs((int)m) - 1;
                m &= m - 1;
                int won = 0;
                if (lane == l)
                    won = atomicCAS((int *)owner + r, 0, 1) == 0;
                won = __shfl_sync(0xffffffffu, won, l);
                if (won)
                    got = __shfl_sync(0xffffffffu, g, l);
            }
        }
        if (got < 0)
        {
            __nanosleep(backoff);
            if (backoff < 8192) backoff <<= 1;
            continue;
        }
        backoff = 256;
        __threadfence();                                    // the previous owner's writes (this row's left neighbours) are visible
        const int f = got / mb_h, row = got - f * mb_h;
        const BatchItem &it = items[f];
        int *owner = it.fp.row_progress + mb_h + 1;
        MbCtx &c = *new (s_ctx[warp]) MbCtx(it.fc, it.fp, work);
        int x = ld_acquire(it.fp.row_progress + row);       // (may have moved between the look and the claim)
        bool finished = false;
        while (x < mb_w && row_ready(it.fc, it.fp, row, x))
        {
            c.mb_x = x; c.mb_y = row; c.mb_xy = row * mb_w + x;
            if (it.fp.trace && lane == 0) it.fp.trace[2 * c.mb_xy] = globaltimer_ns();
            stage_fenc(it.fc, x, row, work);
            analyse_p_mb<F>(c, c.mb_xy ? it.fp.results[c.mb_xy - 1].mv : it.fp.stale_mv);
            __syncwarp();
            if (lane == 0)
            {
                if (it.fp.trace) it.fp.trace[2 * c.mb_xy + 1] = globaltimer_ns();
                __threadfence();
                st_release(it.fp.row_progress + row, x + 1);
            }
            x++;
            finished = x == mb_w;
        }
        if (lane == 0)
        {
            __threadfence();
            st_release(owner + row, 0);
            if (finished)
                atomicAdd(rows_done, 1);
        }
        __syncwarp();
        pos = (unsigned)got + 1u;                           // the row below is the likeliest to have become ready
    }
}

// rows_per_cta: 1 = every row on its own SM (lowest latency for a single encoder), 4 = four consecutive rows share a
// CTA and its L1 (higher throughput when many encoder contexts run concurrently)
// feature mask of the instantiation: bit 0 = exhaustive searches (--me esa / tesa), bit 1 = sub-8x8 partitions; the default
// configuration runs the leanest kernel (code that is merely present in the call graph costs registers and layout)
#define PCAMV_DISPATCH_F(f, CALL) do { switch (f) { case 0: { constexpr int F_ = 0; CALL; } break; case 1: { constexpr int F_ = 1; CALL; } break; \
                                                     case 2: { constexpr int F_ = 2; CALL; } break; default: { constexpr int F_ = 3; CALL; } break; } } while (0)
static int feature_mask(const DevFrameCtx &fc) { return (fc.me_method >= ME_ESA ? 1 : 0) | ((fc.analyse_inter & 0x20) ? 2 : 0); }
void launch_analyse_p(const DevFrameCtx &fc, const FrameParams &fp, int *row_claim, int n_rows, int rows_per_cta, void *stream)
{
    const cudaStream_t st = (cudaStream_t)stream;
    const int f = feature_mask(fc);
    if (rows_per_cta >= 4)      PCAMV_DISPATCH_F(f, (k_analyse_p<4, F_><<<(n_rows + 3) / 4, 128, 0, st>>>(fc, fp, row_claim)));
    else if (rows_per_cta >= 2) PCAMV_DISPATCH_F(f, (k_analyse_p<2, F_><<<(n_rows + 1) / 2, 64, 0, st>>>(fc, fp, row_claim)));
    else                        PCAMV_DISPATCH_F(f, (k_analyse_p<1, F_><<<n_rows, 32, 0, st>>>(fc, fp, row_claim)));
}

// next_group: n_items counters, zeroed by the caller; all items share one configuration (checked by the caller)
void launch_analyse_p_batch(const BatchItem *items, int n_items, int *row_claim, int n_rows, int rows_per_cta, int max_ctas, const DevFrameCtx &fc, void *stream)
{
    const cudaStream_t st = (cudaStream_t)stream;
    const int f = feature_mask(fc);
    if (rows_per_cta < 0)
    {
        // row pool (rows_per_cta = -1): persistent grid of 4-team CTAs, row_claim[0] counts finished rows
        int ctas = max_ctas > 0 ? max_ctas : (n_items * n_rows + 3) / 4;
        if (ctas > (n_items * n_rows + 3) / 4) ctas = (n_items * n_rows + 3) / 4;
        PCAMV_DISPATCH_F(f, (k_analyse_p_pool<F_><<<ctas, 128, 0, st>>>(items, n_items, row_claim)));
        return;
    }
    const int w = rows_per_cta >= 4 ? 4 : rows_per_cta >= 2 ? 2 : 1;
    int ctas = n_items * ((n_rows + w - 1) / w);
    if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
    if (w == 4)      PCAMV_DISPATCH_F(f, (k_analyse_p_batch<4, F_><<<ctas, 128, 0, st>>>(items, n_items, row_claim)));
    else if (w == 2) PCAMV_DISPATCH_F(f, (k_analyse_p_batch<2, F_><<<ctas, 64, 0, st>>>(items, n_items, row_claim)));
    else             PCAMV_DISPATCH_F(f, (k_analyse_p_batch<1, F_><<<ctas, 32, 0, st>>>(items, n_items, row_claim)));
}

// ---- cost table: macroblocks are independent; one CTA of CTP_TEAMS teams per macroblock, one candidate vector per team ----
// x264_ih_get_mv_cost always costs the original vector and the first four replacement candidates, and the other eight only
// when none of those four qualified (analyse.c:2443-2449); each needs its own whole-macroblock reconstruction and 9-point
// ring, independent of the others — the "stop after four" rule is a selection made afterwards.  So team w of the CTA takes
// vector w - 1 of the round (the original, then candidates 0..3; second and third round: candidates 4..7 and 8..11 on teams
// 0..3) and the fold runs over the teams' results in the reference's order.  Besides cutting the latency of a vector five
// ways, the teams of a CTA walk the same code at the same time, which is what the instruction caches reward: the
// one-team-per-macroblock kernel of round 1 kept ~60 KB of code hot under 24 unrelated teams per SM and stalled on instruction
// fetch (profiles/r02_cost_table_base_ncu_summary.txt: no_instruction 4.6 warp-cycles per issue, the largest stall).  Measured
// gain of this layout at 128 frames: 57.9 -> 56.1 ms (profiles/r02_cost_table_ab.txt) — the fetch stalls are per team, not
// per distinct code stream, so most of the win is the latency of one vector.
#define CTP_TEAMS 5
#ifndef PCAMV_CTP_MIN_CTAS
#define PCAMV_CTP_MIN_CTAS 6
#endif
__device__ __forceinline__ void cost_table_cta(const DevFrameCtx &fc, const FrameParams &fp, int mb)
{
    __shared__ MbWork s_work[CTP_TEAMS];
    __shared__ MbResult s_res;
    __shared__ IhCand s_cand[CTP_TEAMS];
    __shared__ __align__(16) unsigned char s_ctx[CTP_TEAMS][sizeof(MbCtx)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const uint32_t *src = (const uint32_t *)(fp.results + mb);
        uint32_t *dst = (uint32_t *)&s_res;
        for (int i = threadIdx.x; i < (int)(sizeof(MbResult) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const MbResult &res = s_res;
    if (res.type == MB_P_SKIP)
        return;
    MbWork &work = s_work[warp];
    MbCtx &c = *new (s_ctx[warp]) MbCtx(fc, fp, work);
    c.mb_x = mb % fc.mb_w; c.mb_y = mb / fc.mb_w; c.mb_xy = mb;
    c.partition = res.partition;
    stage_fenc(fc, c.mb_x, c.mb_y, work);
    c.n_log = res.n_log;
    init_limits(c);
    c.env.cost_mv = fc.tab.cost_mv;
    c.env.me_method = fc.me_method; c.env.me_range = fc.me_range; c.env.subme = fc.subme;
    c.env.chroma_me = fc.chroma_me && fc.subme >= 5;
    c.env.mbcmp_satd = fc.subme > 1;
    c.env.mvsads = nullptr;
    const int n_part = res.n_part;
    for (int k = 0; k < n_part; k++)
    {
        const PartInfo &pi = mb_part(c, res, k);
        const int bmx = pi.mv[0], bmy = pi.mv[1];
        const int kind = ih_setup_block(c, res, k);
        IhFold f;
        int stop = 0;
        // round 0: original + candidates 0..3; rounds 1, 2: candidates 4..7, 8..11 (team 4 has nothing to do there)
        for (int round = 0; round < 3 && !stop; round++)
        {
            const int ii = round == 0 ? warp - 1 : 4 * round + warp;
            if (round == 0 || warp < 4)
            {
                IhCand o;
                ih_eval(c, res, k, ii, kind, bmx, bmy, o);
                if (lane == 0) s_cand[warp] = o;
            }
            __syncthreads();
            if (round == 0)
            {
                ih_fold_orig(f, s_cand[0]);
                for (int j = 0; j < 4 && !stop; j++) stop = ih_fold_cand(f, j, s_cand[1 + j]);
            }
            else
                for (int j = 0; j < 4 && !stop; j++) stop = ih_fold_cand(f, 4 * round + j, s_cand[j]);
            __syncthreads();                       // everybody has read the results before the next round overwrites them
        }
        const int cost_opt = ih_fold_finish(f);
        if (warp == 0)
            log_push(c, LOG_IHCOST, pi.i_pixel, pi.ref, f.m_x, f.m_y, cost_opt, 0);
    }
    if (threadIdx.x == 0)
        fp.results[mb].n_log = res.n_log + n_part;
}

__global__ void __launch_bounds__(CTP_TEAMS * 32, PCAMV_CTP_MIN_CTAS) k_cost_table_cta(const __grid_constant__ DevFrameCtx fc,
                                                                                    const __grid_constant__ FrameParams fp)
{
    cost_table_cta(fc, fp, blockIdx.x);
}
// blockIdx.y = frame of the batch
__global__ void __launch_bounds__(CTP_TEAMS * 32, PCAMV_CTP_MIN_CTAS) k_cost_table_cta_batch(const BatchItem *__restrict__ items)
{
    const BatchItem &it = items[blockIdx.y];
    cost_table_cta(it.fc, it.fp, blockIdx.x);
}

void launch_cost_table(const DevFrameCtx &fc, const FrameParams &fp, int n_mb, void *stream)
{
    k_cost_table_cta<<<n_mb, CTP_TEAMS * 32, 0, (cudaStream_t)stream>>>(fc, fp);
}

} // namespace pcamv

namespace pcamv {
void launch_cost_table_batch(const BatchItem *items, int n_items, int n_mb, void *stream)
{
    k_cost_table_cta_batch<<<dim3(n_mb, n_items), CTP_TEAMS * 32, 0, (cudaStream_t)stream>>>(items);
}
}
