// pcamv_split.cu — the multi-context macroblock wavefront with the motion SEARCHES and the per-macroblock CONTROL code on
// different SMs ("split wavefront", pcamv_cfg.rows_per_cta = -2).
//
// Why.  Measured on B200 (profiles/r02_icache_probe.txt, r02_search_seam_ncu_summary.txt): an SM whose resident warps walk
// more than ~32 KB of code at uncorrelated positions fetches instructions at 0.1-0.4 per clock (22 KB: 2.9, 45 KB: 1.0,
// 90 KB: 0.37, 180 KB: 0.15 — tools/probes/gen_icache_probe.py).  The monolithic wavefront kernel keeps ~100 KB hot (42 KB of
// search code + ~60 KB of neighbour cache / MV prediction / mode decision / P_SKIP probe code that every macroblock runs
// once) and issues on 25 % of the cycles with "no instruction" as the dominant stall; the very same searches, run by
// k_search_batch where nothing else competes for the instruction caches, take 7.8 ns each instead of 20.8.
//
// What.  One persistent kernel, every SM takes ONE role for the whole launch (decided when its first CTA arrives):
//   control SMs  run analyse_p_mb in its resumable form (pcamv_frame.cuh, AS = 1): a team owns up to `rows_per_team` macroblock
//                ROWS, each a parked task (its persistent bytes live in HBM / L2 between steps).  A step = take a row whose
//                awaited event has happened (wavefront dependency satisfied, or search result back), run the analysis until
//                it needs the next search, park it, publish the request.  Nothing ever blocks inside a team: waiting rows
//                wait in memory.  Reference for the logic: encoder/analyse.c:2613-3172, common/macroblock.c:28-470,914-1224.
//   search SMs   take requests from ticket rings, rebuild the block descriptor and run x264_me_search_ref /
//                x264_me_refine_qpel (encoder/me.c:158-843) exactly as k_search_batch does, publish the result.
// Ordering is release/acquire at GPU scope on the ring entries, the per-slot result sequence numbers and the per-row
// progress counters; data that crosses SMs is read with ld.global.cg.
//
// Progress.  Rows of a frame are claimed in increasing order and only once their first macroblock can start, every claimed
// row sits in a resident team's table, the topmost unfinished row of a frame never waits on anything, and search teams never
// wait on control teams except for work: no deadlock for any grid that has at least one SM in each role.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <new>
#include "pcamv_device.h"
#include "pcamv_cost.cuh"
#include "pcamv_split.h"

namespace pcamv {

__device__ __forceinline__ int sp_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sp_st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sp_ld_acquire64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sp_st_release64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sp_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned sp_smid()
{
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

__device__ __forceinline__ void sp_stage_fenc(const DevFrameCtx &fc, int mb_x, int mb_y, MbWork &w)
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

// bytes of a parked row: the persistent head of MbWork, then the scalars of MbCtx (everything behind its three references)
#define SP_KEEP_WORK ((int)offsetof(MbWork, fenc_y))
__device__ __forceinline__ int sp_ctx_off(const MbCtx &c) { return (int)((const char *)&c.mb_x - (const char *)&c); }

enum { SLOT_FREE = 0, SLOT_WAIT_DEP = 1, SLOT_WAIT_SEARCH = 2 };

struct TeamTable            // per control team, in shared memory
{
    int state[SPLIT_MAX_ROWS], item[SPLIT_MAX_ROWS], row[SPLIT_MAX_ROWS], x[SPLIT_MAX_ROWS], seq[SPLIT_MAX_ROWS];
    int kind[SPLIT_MAX_ROWS];        // which piece of the analysis the row's next step runs (step_kind)
};

// The next step of a parked macroblock, by the code it will run: 0 = start of a macroblock (neighbour cache, predictors, first
// 16x16 search out), 1 = a 16x16 result (P_SKIP probe, on to the 8x8 searches), 2 = an 8x8 result, 3 = a 16x8 / 8x16 result,
// 4 = a refinement result (decision, records).  Kind-affine scheduling (PCAMV_SPLIT_PHASE_NS > 0): all control teams of the
// GPU prefer, at any time, the steps of ONE kind — the kind is a function of %globaltimer, so nobody has to agree on it — and a
// control SM's instruction caches see one ~10 KB piece of code for a while instead of all 60 KB at once.
__device__ __forceinline__ int step_kind(const PtState &pt)
{
    return pt.in16 ? 1 : pt.in8 ? 2 : pt.in168 ? 3 : 4;
}
// phase schedule: the kinds in the order and proportion a macroblock needs them (1 start, 1 x 16x16, 4 x 8x8, 3 x 16x8 / 8x16, 1 refine)
__device__ __forceinline__ int phase_kind(unsigned long long now, unsigned phase_ns)
{
    const unsigned slot = (unsigned)((now / phase_ns) % 10ull);
    return slot == 0 ? 0 : slot == 1 ? 1 : slot <= 5 ? 2 : slot <= 8 ? 3 : 4;
}

__device__ __forceinline__ bool sp_dep_ready(const DevFrameCtx &fc, const FrameParams &fp, int row, int x)
{
    if (row == 0)
        return true;
    int need = min(x + 2, fc.mb_w);
    // quirk q2 (analyse.c:2668-2676): a macroblock forced to P_SKIP in pass 2 may keep the MV cache of the previous macroblock
    // in raster order; for x == 0 that is the end of the row above
    if (x == 0 && fp.pass == 2 && fp.forced[row * fc.mb_w].type == MB_P_SKIP)
        need = fc.mb_w;
    return sp_ld_acquire(fp.row_progress + row - 1) >= need;
}

// ---- search team ------------------------------------------------------------------------------------------------------
template <int XS>
__device__ __forceinline__ void search_team(const BatchItem *__restrict__ items, const SplitBufs sb, MbWork &work)
{
    const int lane = threadIdx.x & 31;
    int wid = 0;
    if (lane == 0) wid = atomicAdd(sb.hdr + SPH_WORKERS, 1);
    wid = __shfl_sync(0xffffffffu, wid, 0);
    const int q = wid % SPLIT_NQ;
    const unsigned long long t_start = sp_globaltimer();
    unsigned long long *ring = sb.ring + (size_t)q * SPLIT_RING_CAP;
    unsigned long long *head = sb.heads + (size_t)q * 16;          // one 128-byte line per counter
    unsigned long long *stats = (unsigned long long *)(sb.hdr + SPH_STATS);
    unsigned long long s_wait = 0, s_serve = 0, s_n = 0, tm = t_start;
#define SP_WORKER_EXIT() do { if (lane == 0) { atomicAdd(stats + 0, s_wait); atomicAdd(stats + 1, s_serve); atomicAdd(stats + 2, s_n); atomicAdd(stats + 3, 1ull); } } while (0)
    for (;;)
    {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(head, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        const unsigned long long *e = ring + (t & (SPLIT_RING_CAP - 1));
        unsigned long long v;
        unsigned ns = 32;
        for (;;)
        {
            v = sp_ld_acquire64(e);
            if ((v >> 32) == t + 1)
                break;
            if (sp_ld_acquire(sb.hdr + SPH_DONE))
            {
                s_wait += sp_globaltimer() - tm;
                SP_WORKER_EXIT();
                return;                                            // every row of every frame is finished: no request will come
            }
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
            else if (sp_globaltimer() - t_start > SPLIT_WATCHDOG_NS)
            {
                if (lane == 0) { sp_st_release(sb.hdr + SPH_ABORT, 1); sp_st_release(sb.hdr + SPH_DONE, 1); }
                return;
            }
        }
        { const unsigned long long now = sp_globaltimer(); s_wait += now - tm; tm = now; }
        const unsigned gslot = (unsigned)v;
        // the request: 32 words, one per lane
        {
            const uint32_t *src = (const uint32_t *)(sb.reqs + (size_t)gslot * SPLIT_REQ_STRIDE);
            uint32_t *dst = (uint32_t *)&work.rq;
            if (lane < (int)(sizeof(SearchReq) / 4)) dst[lane] = __ldcg(src + lane);
            __syncwarp();
        }
        const BatchItem &it = items[work.rq.item];
        sp_stage_fenc(it.fc, work.rq.mb_x, work.rq.mb_y, work);
        const int seq = work.rq.pad2;
        SearchRes out;
        serve_request<XS>(it.fc, it.fp, work, work.rq, out);
        __syncwarp();
        if (lane == 0)
        {
            SearchRes *dst = (SearchRes *)(sb.res + (size_t)gslot * sizeof(SearchRes));
            *(uint4 *)dst = *(const uint4 *)&out;
            __threadfence();
            sp_st_release(sb.ready + gslot, seq);
        }
        { const unsigned long long now = sp_globaltimer(); s_serve += now - tm; tm = now; s_n++; }
    }
}

// ---- control team ----------------------------------------------------------------------------------------------------
// claim the next row of some frame whose first macroblock can start now; -1 = none right now; *left = rows remain unclaimed
__device__ __forceinline__ int claim_row(const BatchItem *__restrict__ items, int n_items, int *next_row, unsigned rot, int mb_h, int mb_w,
                                         int *out_row, bool *left)
{
    const int lane = threadIdx.x & 31;
    bool any_left = false;
    for (int base = 0; base < n_items; base += 32)
    {
        const int t = base + lane;
        int f = 0, r = mb_h;
        bool ok = false;
        if (t < n_items)
        {
            f = (int)((rot + (unsigned)t) % (unsigned)n_items);
            r = sp_ld_acquire(next_row + f);
            if (r < mb_h)
            {
                any_left = true;
                ok = r == 0 || sp_ld_acquire(items[f].fp.row_progress + r - 1) >= min(2, mb_w);
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        while (m)
        {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            int won = 0;
            if (lane == l)
                won = atomicCAS(next_row + f, r, r + 1) == r;
            won = __shfl_sync(0xffffffffu, won, l);
            if (won)
            {
                *out_row = __shfl_sync(0xffffffffu, r, l);
                *left = true;
                return __shfl_sync(0xffffffffu, f, l);
            }
        }
    }
    *left = __any_sync(0xffffffffu, any_left);
    return -1;
}

template <int F>
__device__ __forceinline__ void control_team(const BatchItem *__restrict__ items, int n_items, int *next_row, const SplitBufs sb,
                                             MbWork &work, unsigned char *ctx_mem, TeamTable &tt, int rows_per_team, unsigned phase_ns,
                                             unsigned phase_patience_ns)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned gslot0 = (unsigned)gwarp * SPLIT_MAX_ROWS;
    const int mb_h = items[0].fc.mb_h, mb_w = items[0].fc.mb_w;
    const int total_rows = n_items * mb_h;
    if (lane < SPLIT_MAX_ROWS) { tt.state[lane] = SLOT_FREE; tt.seq[lane] = 0; }
    __syncwarp();
    unsigned rot = (unsigned)gwarp * 7u;
    unsigned rr = 0;                     // round-robin start of the slot scan
    bool exhausted = false;
    int claim_holdoff = 0;
    unsigned backoff = 64;
    const unsigned long long t_start = sp_globaltimer();
    unsigned long long *stats = (unsigned long long *)(sb.hdr + SPH_STATS);
    unsigned long long s_idle = 0, s_claim = 0, s_in = 0, s_run = 0, s_out = 0, s_steps = 0, tm = t_start, wait0 = 0, s_match = 0;
#define SP_TICK(acc) do { const unsigned long long now_ = sp_globaltimer(); acc += now_ - tm; tm = now_; } while (0)
#define SP_CTRL_EXIT() do { if (lane == 0) { atomicAdd(stats + 4, s_idle); atomicAdd(stats + 5, s_claim); atomicAdd(stats + 6, s_in); atomicAdd(stats + 7, s_run); \
        atomicAdd(stats + 8, s_out); atomicAdd(stats + 9, s_steps); atomicAdd(stats + 10, 1ull); atomicAdd(stats + 11, sp_globaltimer() - t_start); \
        atomicAdd(stats + 12, s_match); } } while (0)
    for (;;)
    {
        // ---- fill a free slot with a fresh row --------------------------------------------------------------------------
        const unsigned free_m = __ballot_sync(0xffffffffu, lane < rows_per_team && tt.state[lane] == SLOT_FREE);
        if (free_m && !exhausted && claim_holdoff == 0)
        {
            int row = 0;
            bool left = false;
            const int f = claim_row(items, n_items, next_row, rot, mb_h, mb_w, &row, &left);
            rot += 5u;
            if (f >= 0)
            {
                const int k = __ffs((int)free_m) - 1;
                if (lane == 0) { tt.state[k] = SLOT_WAIT_DEP; tt.item[k] = f; tt.row[k] = row; tt.x[k] = 0; tt.kind[k] = 0; }
                __syncwarp();
            }
            else
            {
                exhausted = !left;
                claim_holdoff = 4;
            }
        }
        else if (claim_holdoff)
            claim_holdoff--;
        SP_TICK(s_claim);

        // ---- which rows can take a step? ---------------------------------------------------------------------------------
        bool ok = false;
        if (lane < rows_per_team)
        {
            const int st = tt.state[lane];
            if (st == SLOT_WAIT_SEARCH)
                ok = sp_ld_acquire(sb.ready + gslot0 + lane) == tt.seq[lane];
            else if (st == SLOT_WAIT_DEP)
                ok = sp_dep_ready(items[tt.item[lane]].fc, items[tt.item[lane]].fp, tt.row[lane], tt.x[lane]);
        }
        const unsigned ready_m = __ballot_sync(0xffffffffu, ok);
        if (!ready_m)
        {
            const unsigned busy_m = __ballot_sync(0xffffffffu, lane < rows_per_team && tt.state[lane] != SLOT_FREE);
            if (!busy_m && exhausted)
            {
                SP_CTRL_EXIT();
                return;
            }
            __nanosleep(backoff);
            SP_TICK(s_idle);
            if (backoff < 2048) backoff <<= 1;
            else if (sp_ld_acquire(sb.hdr + SPH_ABORT) || sp_globaltimer() - t_start > SPLIT_WATCHDOG_NS)
            {
                if (lane == 0) { sp_st_release(sb.hdr + SPH_ABORT, 1); sp_st_release(sb.hdr + SPH_DONE, 1); }
                return;
            }
            continue;
        }
        backoff = 64;
        // kind-affine pick: a ready step of the current phase's kind if there is one; otherwise wait a little for one to turn up
        // (other teams of this SM are running that kind right now), then take whatever is ready
        unsigned pick_m = ready_m;
        if (phase_ns)
        {
            const unsigned long long now = sp_globaltimer();
            const int want = phase_kind(now, phase_ns);
            const unsigned match_m = __ballot_sync(0xffffffffu, ok && tt.kind[lane] == want);
            if (match_m) { pick_m = match_m; wait0 = 0; s_match++; }
            else
            {
                if (!wait0) wait0 = now;
                if (now - wait0 < phase_patience_ns) { __nanosleep(200); SP_TICK(s_idle); continue; }
                wait0 = 0;
            }
        }
        // first picked slot at or after rr
        const unsigned rotm = (pick_m >> rr) | (rr ? pick_m << (32 - rr) : 0u);
        const int k = (int)((rr + (unsigned)(__ffs((int)rotm) - 1)) & 31u);
        rr = (unsigned)(k + 1) % (unsigned)rows_per_team;

        // ---- one step of row slot k -------------------------------------------------------------------------------------------
        const int item = tt.item[k], row = tt.row[k], x = tt.x[k];
        const bool resume = tt.state[k] == SLOT_WAIT_SEARCH;
        const BatchItem &it = items[item];
        const unsigned gslot = gslot0 + (unsigned)k;
        MbCtx &c = *new (ctx_mem) MbCtx(it.fc, it.fp, work);      // every lane writes the same values
        const int ctx_off = sp_ctx_off(c), ctx_words = ((int)sizeof(MbCtx) - ctx_off) / 4;
        unsigned char *park = sb.park + (size_t)gslot * sb.park_stride;
        if (resume)
        {
            const uint4 *src = (const uint4 *)park;
            uint4 *dst = (uint4 *)&work;
            for (int i = lane; i < SP_KEEP_WORK / 16; i += 32) dst[i] = __ldcg(src + i);
            const uint32_t *src2 = (const uint32_t *)(park + SP_KEEP_WORK);
            uint32_t *dst2 = (uint32_t *)((char *)&c + ctx_off);
            for (int i = lane; i < ctx_words; i += 32) dst2[i] = __ldcg(src2 + i);
            if (lane < 4) ((uint32_t *)&work.rs)[lane] = __ldcg((const uint32_t *)(sb.res + (size_t)gslot * sizeof(SearchRes)) + lane);
        }
        else
        {
            c.mb_x = x; c.mb_y = row; c.mb_xy = row * mb_w + x;
            work.pt.stage = 0;
        }
        __syncwarp();
        if (it.fp.trace && lane == 0 && !resume) it.fp.trace[2 * c.mb_xy] = sp_globaltimer();
        sp_stage_fenc(it.fc, x, row, work);
        SP_TICK(s_in);
        const int r = analyse_p_mb_rs<F, 1>(c, c.mb_xy ? it.fp.results[c.mb_xy - 1].mv : it.fp.stale_mv);
        __syncwarp();
        SP_TICK(s_run);
        s_steps++;
        if (r == PT_YIELD)
        {
            const int seq = tt.seq[k] + 1;
            if (lane == 0) { work.rq.item = item; work.rq.pad2 = seq; tt.seq[k] = seq; tt.state[k] = SLOT_WAIT_SEARCH; tt.kind[k] = step_kind(work.pt); }
            __syncwarp();
            {
                uint4 *dst = (uint4 *)park;
                const uint4 *src = (const uint4 *)&work;
                for (int i = lane; i < SP_KEEP_WORK / 16; i += 32) __stcg(dst + i, src[i]);
                uint32_t *dst2 = (uint32_t *)(park + SP_KEEP_WORK);
                const uint32_t *src2 = (const uint32_t *)((const char *)&c + ctx_off);
                for (int i = lane; i < ctx_words; i += 32) __stcg(dst2 + i, src2[i]);
                uint32_t *rq = (uint32_t *)(sb.reqs + (size_t)gslot * SPLIT_REQ_STRIDE);
                if (lane < (int)(sizeof(SearchReq) / 4)) __stcg(rq + lane, ((const uint32_t *)&work.rq)[lane]);
            }
            __syncwarp();
            if (lane == 0)
            {
                __threadfence();
                const int q = (int)((gslot + (unsigned)seq) % SPLIT_NQ);
                const unsigned long long t = atomicAdd(sb.tails + (size_t)q * 16, 1ull);
                sp_st_release64(sb.ring + (size_t)q * SPLIT_RING_CAP + (t & (SPLIT_RING_CAP - 1)), ((t + 1) << 32) | gslot);
            }
        }
        else
        {
            if (lane == 0)
            {
                if (it.fp.trace) it.fp.trace[2 * c.mb_xy + 1] = sp_globaltimer();
                __threadfence();
                sp_st_release(it.fp.row_progress + row, x + 1);
                if (x + 1 == mb_w)
                {
                    tt.state[k] = SLOT_FREE;
                    if (atomicAdd(sb.hdr + SPH_ROWS_DONE, 1) + 1 == total_rows)
                        sp_st_release(sb.hdr + SPH_DONE, 1);
                }
                else
                {
                    tt.x[k] = x + 1;
                    tt.state[k] = SLOT_WAIT_DEP;
                    tt.kind[k] = 0;
                }
            }
        }
        __syncwarp();
        SP_TICK(s_out);
    }
}

template <int F>
__global__ void __launch_bounds__(128, PCAMV_SPLIT_MIN_CTAS) k_analyse_p_split(const BatchItem *__restrict__ items, int n_items, int *next_row,
                                                                              const SplitBufs sb, int n_ctrl_sms, int n_sms, int rows_per_team,
                                                                              unsigned phase_ns, unsigned phase_patience_ns)
{
    __shared__ MbWork s_work[4];
    __shared__ __align__(16) unsigned char s_ctx[4][sizeof(MbCtx)];
    __shared__ TeamTable s_tt[4];
    __shared__ int s_role;
    const int warp = threadIdx.x >> 5;
    // ---- the SM's role: decided by the first CTA that arrives on it ---------------------------------------------------------
    if (threadIdx.x == 0)
    {
        int *slot = sb.hdr + SPH_SM_ROLE + (sp_smid() & (SPLIT_MAX_SMS - 1));
        int role = sp_ld_acquire(slot);
        if (role == 0 && atomicCAS(slot, 0, 1) == 0)
        {
            // k-th SM to register: control SMs are spread evenly over the arrival order (the very first one is a control SM)
            const int k = atomicAdd(sb.hdr + SPH_SMS, 1);
            role = (int)(((long long)(k + 1) * n_ctrl_sms) / n_sms) > (int)(((long long)k * n_ctrl_sms) / n_sms) || k == 0 ? 2 : 3;
            if (k == 1 && n_ctrl_sms < n_sms) role = 3;            // ... and the second one a search SM, whatever the ratio
            sp_st_release(slot, role);
        }
        else
            while ((role = sp_ld_acquire(slot)) < 2) __nanosleep(64);
        s_role = role;
    }
    __syncthreads();
    if (s_role == 2)
        control_team<F>(items, n_items, next_row, sb, s_work[warp], s_ctx[warp], s_tt[warp], rows_per_team, phase_ns, phase_patience_ns);
    else
        search_team<(F & 1)>(items, sb, s_work[warp]);
}

size_t split_bytes(int total_warps, size_t *park_stride)
{
    const size_t stride = ((size_t)SP_KEEP_WORK + sizeof(MbCtx) + 127) & ~(size_t)127;
    if (park_stride) *park_stride = stride;
    const size_t n_slots = (size_t)total_warps * SPLIT_MAX_ROWS;
    size_t b = SPLIT_HDR_INTS * sizeof(int);
    b += 2 * (size_t)SPLIT_NQ * 16 * sizeof(unsigned long long);                // heads, tails (a 128-byte line each)
    b += (size_t)SPLIT_NQ * SPLIT_RING_CAP * sizeof(unsigned long long);
    b += n_slots * sizeof(int);                                                 // ready
    b = (b + 127) & ~(size_t)127;
    b += n_slots * SPLIT_REQ_STRIDE + n_slots * sizeof(SearchRes) + n_slots * stride;
    return b + 256;
}

// carves the buffers out of one allocation; the part that must be zero at every launch comes first (zero_bytes)
SplitBufs split_carve(unsigned char *base, int total_warps, size_t *zero_bytes)
{
    SplitBufs sb;
    size_t stride;
    split_bytes(total_warps, &stride);
    const size_t n_slots = (size_t)total_warps * SPLIT_MAX_ROWS;
    unsigned char *p = base;
    sb.hdr = (int *)p; p += SPLIT_HDR_INTS * sizeof(int);
    sb.heads = (unsigned long long *)p; p += (size_t)SPLIT_NQ * 16 * sizeof(unsigned long long);
    sb.tails = (unsigned long long *)p; p += (size_t)SPLIT_NQ * 16 * sizeof(unsigned long long);
    sb.ring = (unsigned long long *)p; p += (size_t)SPLIT_NQ * SPLIT_RING_CAP * sizeof(unsigned long long);
    sb.ready = (int *)p; p += n_slots * sizeof(int);
    p = base + (((size_t)(p - base) + 127) & ~(size_t)127);
    if (zero_bytes) *zero_bytes = (size_t)(p - base);
    sb.reqs = p; p += n_slots * SPLIT_REQ_STRIDE;
    sb.res = p; p += n_slots * sizeof(SearchRes);
    sb.park = p;
    sb.park_stride = (unsigned)stride;
    return sb;
}

void launch_analyse_p_split(const BatchItem *items, int n_items, int *next_row, const SplitBufs &sb, int ctas, int n_ctrl_sms, int n_sms,
                            int rows_per_team, int feature, void *stream)
{
    const cudaStream_t st = (cudaStream_t)stream;
    // kind-affine control scheduling: PCAMV_SPLIT_PHASE_NS = length of a phase (0 / unset = off), PCAMV_SPLIT_PATIENCE_NS = how long a
    // team with only other kinds ready waits for one of the phase's kind before it takes what it has
    unsigned phase_ns = 0, patience_ns = 2000;
    if (const char *e = getenv("PCAMV_SPLIT_PHASE_NS")) phase_ns = (unsigned)atoi(e);
    if (const char *e = getenv("PCAMV_SPLIT_PATIENCE_NS")) patience_ns = (unsigned)atoi(e);
    if (feature & 1) k_analyse_p_split<1><<<ctas, 128, 0, st>>>(items, n_items, next_row, sb, n_ctrl_sms, n_sms, rows_per_team, phase_ns, patience_ns);
    else             k_analyse_p_split<0><<<ctas, 128, 0, st>>>(items, n_items, next_row, sb, n_ctrl_sms, n_sms, rows_per_team, phase_ns, patience_ns);
}

} // namespace pcamv
