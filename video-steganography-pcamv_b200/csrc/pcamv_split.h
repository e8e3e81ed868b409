// pcamv_split.h — buffers and constants of the split wavefront (pcamv_split.cu), shared with the C-ABI layer.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace pcamv {

#define SPLIT_NQ 8                    // request rings (a search team serves one; spreads the ticket atomics over 8 lines)
#define SPLIT_RING_CAP 65536          // entries per ring, >= row slots of the largest grid: a ring can never wrap onto a pending entry
#define SPLIT_MAX_ROWS 16             // row slots per control team
#define SPLIT_REQ_STRIDE 128          // bytes per request record
#define SPLIT_MAX_SMS 512
#ifndef PCAMV_SPLIT_MIN_CTAS
#define PCAMV_SPLIT_MIN_CTAS 6        // CTAs of 4 teams per SM the register budget must allow
#endif
// header words
enum { SPH_ROWS_DONE = 0, SPH_DONE = 1, SPH_WORKERS = 2, SPH_SMS = 3, SPH_ABORT = 4, SPH_STATS = 8 /* 13 x u64: ints 8 .. 33 */, SPH_SM_ROLE = 40 };
// SPH_STATS (ns summed over teams, globaltimer): search teams 0 waiting for a request, 1 serving, 2 requests served, 3 teams;
// control teams 4 idle (nothing runnable), 5 claiming rows, 6 restore + stage, 7 analysis, 8 park + publish, 9 steps, 10 teams, 11 lifetime, 12 steps that matched the phase's kind
#define SPLIT_WATCHDOG_NS 20000000000ull    // a team that has waited this long for anything gives up and stops the launch (reported by the host)
#define SPLIT_HDR_INTS (SPH_SM_ROLE + SPLIT_MAX_SMS)

struct SplitBufs
{
    int *hdr;                         // [SPLIT_HDR_INTS]: counters, "all rows done" flag, role of every SM
    unsigned long long *heads, *tails;    // per ring, one 128-byte line each
    unsigned long long *ring;         // [SPLIT_NQ][SPLIT_RING_CAP]: (ticket + 1) << 32 | row slot
    int *ready;                       // per row slot: sequence number of the last search served
    unsigned char *reqs;              // per row slot: SearchReq
    unsigned char *res;               // per row slot: SearchRes
    unsigned char *park;              // per row slot: the parked macroblock state
    unsigned park_stride;
};

struct BatchItem;
size_t split_bytes(int total_warps, size_t *park_stride);
SplitBufs split_carve(unsigned char *base, int total_warps, size_t *zero_bytes);
void launch_analyse_p_split(const BatchItem *items, int n_items, int *next_row, const SplitBufs &sb, int ctas, int n_ctrl_sms, int n_sms,
                            int rows_per_team, int feature, void *stream);

} // namespace pcamv
