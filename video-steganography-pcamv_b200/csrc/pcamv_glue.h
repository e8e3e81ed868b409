// pcamv_glue.h — host-side restatement of how pass 2 of a P frame forces the pass-1 decisions
// (reference encoder/analyse.c:2870-2991) and applies the STC flips (analyse.c:3001-3107), so that the
// GPU wavefront of pass 2 knows every macroblock's final vectors up front.  Plain C++, no CUDA.
//
// Input per MB is the reference's h->info.cache[] entry verbatim (common/common.h:585-603), including
// the two quirks of how it is filled (analyse.c:3526-3632): ref[] is the 4x4 ref cache in RASTER order,
// mv[] is the copy produced by the unsequenced `idx++` loops.
#pragma once
#include <stdint.h>

namespace pcamv {

struct Pass1Mb               // mirrors the fields of h->info.cache[mb] that pass 2 reads
{
    int32_t type;            // P_L0 = 4, P_8x8 = 5, P_SKIP = 6
    int32_t partition;       // D_16x8 = 14, D_8x16 = 15, D_16x16 = 16
    uint8_t used;
    uint8_t sub[4];
    int8_t ref[16];
    int16_t mv[16][2];
    int16_t mv_stego[16][2];
};

struct ForcedOut             // == ForcedMb of pcamv_frame.cuh
{
    int8_t type, used, partition, pad;
    int8_t ref[4];
    uint32_t mv[16];
};

#if defined(__CUDACC__)
  #define PCAMV_HD __host__ __device__
#else
  #define PCAMV_HD
#endif

PCAMV_HD static inline uint32_t glue_pack(const int16_t v[2]) { return ((uint32_t)(uint16_t)v[0]) | ((uint32_t)(uint16_t)v[1] << 16); }
// block_idx -> (x,y) of the 4x4 block
PCAMV_HD static inline int glue_bx(int idx) { return (idx & 1) | ((idx >> 1) & 2); }
PCAMV_HD static inline int glue_by(int idx) { return ((idx >> 1) & 1) | ((idx >> 2) & 2); }

// One macroblock: fills `o` from the pass-1 record and the flips of ITS carriers (filp = first of them, in cover order);
// returns how many it consumed.  Runs on the host (build_forced) and on the device (k_embed_forced, pcamv_embed.cu).
PCAMV_HD static inline int build_forced_mb(const Pass1Mb &p, const int8_t *filp, ForcedOut &o)
{
    int n = 0;
    {
        {
        o.type = (int8_t)p.type; o.used = (int8_t)p.used; o.partition = (int8_t)p.partition; o.pad = 0;
        for (int i = 0; i < 4; i++) o.ref[i] = 0;
        for (int i = 0; i < 16; i++) o.mv[i] = 0;
        if (!p.used)
            return 0;
        if (p.type == 4)
        {
            if (p.partition == 16)
            {
                const int16_t *mv = filp[n++] == 1 ? p.mv_stego[0] : p.mv[0];
                for (int i = 0; i < 4; i++) o.ref[i] = p.ref[0];
                for (int i = 0; i < 16; i++) o.mv[i] = glue_pack(mv);
            }
            else if (p.partition == 14)         // 16x8: slots 0 and 8 (analyse.c:2911-2916, 3079-3087)
            {
                for (int j = 0; j < 2; j++)
                {
                    const int16_t *mv = filp[n++] == 1 ? p.mv_stego[8 * j] : p.mv[8 * j];
                    o.ref[2 * j] = o.ref[2 * j + 1] = p.ref[8 * j];
                    for (int i = 0; i < 16; i++) if ((glue_by(i) >> 1) == j) o.mv[i] = glue_pack(mv);
                }
            }
            else                                 // 8x16: slots 0 and 4 (analyse.c:2923-2928, 3069-3077)
            {
                for (int j = 0; j < 2; j++)
                {
                    const int16_t *mv = filp[n++] == 1 ? p.mv_stego[4 * j] : p.mv[4 * j];
                    o.ref[j] = o.ref[j + 2] = p.ref[4 * j];
                    for (int i = 0; i < 16; i++) if ((glue_bx(i) >> 1) == j) o.mv[i] = glue_pack(mv);
                }
            }
        }
        else if (p.type == 5)                    // P_8x8: per 8x8 block by its split (analyse.c:2944-2978, 3005-3052)
        {
            for (int i8 = 0; i8 < 4; i8++)
            {
                o.ref[i8] = p.ref[4 * i8];
                const int kind = p.sub[i8];      // D_L0_4x4 = 0, D_L0_8x4 = 1, D_L0_4x8 = 2, D_L0_8x8 = 3
                if (kind == 3)
                {
                    const int16_t *mv = filp[n++] == 1 ? p.mv_stego[4 * i8] : p.mv[4 * i8];
                    for (int i = 0; i < 4; i++) o.mv[4 * i8 + i] = glue_pack(mv);
                }
                else if (kind == 2)              // 4x8: slots 4i, 4i+1 = left / right column
                {
                    for (int j = 0; j < 2; j++)
                    {
                        const int16_t *mv = filp[n++] == 1 ? p.mv_stego[4 * i8 + j] : p.mv[4 * i8 + j];
                        o.mv[4 * i8 + j] = o.mv[4 * i8 + j + 2] = glue_pack(mv);
                    }
                }
                else if (kind == 1)              // 8x4: slots 4i, 4i+2 = top / bottom row
                {
                    for (int j = 0; j < 2; j++)
                    {
                        const int16_t *mv = filp[n++] == 1 ? p.mv_stego[4 * i8 + 2 * j] : p.mv[4 * i8 + 2 * j];
                        o.mv[4 * i8 + 2 * j] = o.mv[4 * i8 + 2 * j + 1] = glue_pack(mv);
                    }
                }
                else                             // 4x4: slots 4i .. 4i+3
                {
                    for (int j = 0; j < 4; j++)
                    {
                        const int16_t *mv = filp[n++] == 1 ? p.mv_stego[4 * i8 + j] : p.mv[4 * i8 + j];
                        o.mv[4 * i8 + j] = glue_pack(mv);
                    }
                }
            }
        }
        }
    }
    return n;
}

// Fills out[0..n_mb) and returns the number of filp[] entries consumed (== number of carrier MVs).
static inline int build_forced(int n_mb, const Pass1Mb *in, const int8_t *filp, ForcedOut *out)
{
    int n = 0;
    for (int mb = 0; mb < n_mb; mb++)
        n += build_forced_mb(in[mb], filp + n, out[mb]);
    return n;
}

} // namespace pcamv
