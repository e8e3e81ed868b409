// pcamv_frame_types.h — plain-data records of the frame-level analysis, shared by the kernels (pcamv_frame.cuh),
// the C-ABI layer and the CPU emulation checkers.  The ABI-visible ones are mirrored in include/pcamv.h
// (pcamv_log_entry == LogEntry, pcamv_mb_out == MbResult; static_asserts in pcamv_frame_api.cu).
#pragma once
#include <stdint.h>
#include "../../include/pcamv.h"
#include "pcamv_device.h"

namespace pcamv {

enum { MB_P_L0 = 4, MB_P_8x8 = 5, MB_P_SKIP = 6 };
enum { PART_8x8 = 13, PART_16x8 = 14, PART_8x16 = 15, PART_16x16 = 16 };
enum { LOG_SEARCH = 0, LOG_REFINE = 1, LOG_IHCOST = 2 };

struct LogEntry              // 16 bytes
{
    int8_t kind, i_pixel, i_ref, pad;
    int16_t mv[2];           // search/refine: resulting mv; ih-cost: chosen delta (m_x, m_y)
    int32_t cost;            // search/refine: m->cost; ih-cost: cost_opt
    int32_t cost_mv;         // search/refine: m->cost_mv (thresh_out in the upper use is not needed by the host)
};

struct ForcedMb              // pass-2 input per MB: the pass-1 decision with the STC flips applied (host glue)
{
    int8_t type, used, partition, pad;
    int8_t ref[4];           // per 8x8 block
    uint32_t mv[16];         // packed (x & 0xffff) | (y << 16), block_idx order
};

struct PartInfo { int16_t mv[2]; int16_t mvp[2]; int8_t ref, i_pixel, xoff, yoff; };

struct MbResult              // what the analysis leaves for the cost-table kernel and the tests
{
    int8_t type, partition, n_part, early_skip;
    int8_t ref[4];
    uint32_t mv[16];         // final cache MVs, block_idx order
    PartInfo part[4];        // partitions of the final mode that carry an MV
    int32_t n_log;
    int16_t pskip_mv[2];
};

struct FrameArrays           // per-frame motion state in HBM (h->mb.type / ref / mv / mvr of the reference)
{
    int8_t *type;            // [mb_h * mb_w]
    int8_t *ref8;            // [2*mb_h][2*mb_w]
    uint32_t *mv4;           // [4*mb_h][4*mb_w] packed
    uint32_t *mvr;           // [max_refs][mb_h*mb_w] packed: 16x16 search result per reference
};

struct FrameParams
{
    int pass;                // 0 = no embedding, 1 = pre-encode, 2 = final encode (decisions forced from pass 1)
    int n_ref;
    int ref_slot[PCAMV_MAX_REFS];
    int ref_poc[PCAMV_MAX_REFS];
    int cur_poc;
    int col_n_ref;           // fref0[0]->i_ref[0]; > 0 enables the temporal candidates
    int col_inv_ref_poc[PCAMV_MAX_REFS];
    const int8_t *col_ref8;
    const uint32_t *col_mv4;
    const ForcedMb *forced;  // pass 2 only
    uint32_t stale_mv[16];   // what the MV cache held before MB 0 of this pass (quirk q2)
    FrameArrays cur;
    LogEntry *log;           // [n_mb][log_stride]
    int log_stride;          // entries per macroblock (pcamv_log_stride)
    MbResult *results;       // [n_mb]
    PartInfo *subparts;      // [n_mb][16], P_8x8 macroblocks only (X264_ANALYSE_PSUB8x8), else null: the up to 16 MV-carrying
                             // blocks of the final mode in the reference's cost-table order (device-internal, not in the ABI)
    int *row_progress;       // [2*mb_h + 2]: [0, mb_h) wavefront counters (macroblocks finished per row) | [mb_h] group claim counter | [mb_h+1, 2*mb_h+1] row owners (row pool)
    unsigned long long *mvsads;  // --me tesa: [mb_h][mvsads_cap] candidate lists, one per macroblock row (= per lane team), else null
    int mvsads_cap;
    unsigned long long *trace;   // optional [n_mb][2]: globaltimer ns at the start / end of each macroblock (profiling aid)
};

struct BatchItem             // one frame of a multi-context launch (pcamv_frame_run_batch): lives in device memory
{
    DevFrameCtx fc;
    FrameParams fp;
};

} // namespace pcamv
