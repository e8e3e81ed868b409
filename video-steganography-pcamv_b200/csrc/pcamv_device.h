// pcamv_device.h — device-resident state shared by the kernels and the C-ABI layer.
#pragma once
#include <stdint.h>
#include "../../include/pcamv.h"

namespace pcamv {

#define PCAMV_PADH 32     // luma border in pixels (reference common/frame.h:28-29); chroma uses half
#define PCAMV_PADV 32
#define PCAMV_SLOTS (PCAMV_MAX_REFS + 2)

// One reference frame in HBM.  Layout mirrors the host encoder's x264_frame_t planes so that whole
// buffers can be compared byte for byte: four luma planes (integer, H, V, HV) of
// stride_y x (height + 64) bytes, two chroma planes of stride_c x (height/2 + 32) bytes.
// y[k] / u / v point at pixel (0,0), i.e. PADV rows and PADH columns into the padded buffer.
struct DevRef
{
    uint8_t *y[4];
    uint8_t *u, *v;
    uint16_t *integral;     // --me esa / tesa only, else null: 8x8 box sums (k_box_sum), same geometry as y[0], at pixel (0,0)
    uint16_t *integral4;    // 4x4 box sums, only when sub-8x8 partitions are searched exhaustively
    int poc;
    int valid;
    const uint8_t *base;    // the slot's allocation (4 luma planes | U | V | slack) and its size: what the PCAMV_CHECKED build checks loads against
    size_t bytes;
};

struct DevTables
{
    const int16_t *cost_mv;          // centre pointer (entry 0 <-> mv difference 0)
    const uint16_t *cost_ref;        // 3*33
    const uint16_t *quant4_mf[2];    // 16 each (inter luma, inter chroma) for the current qp
    const uint16_t *quant4_bias[2];
    const int32_t *dequant4_mf[2];   // 6*16 each
    int qp, lambda, lambda2_chroma, chroma_qp;
};

// Everything a kernel needs, passed by value as a __grid_constant__ parameter.
struct DevFrameCtx
{
    int width, height, mb_w, mb_h;
    int stride_y, stride_c;          // bytes per row of the padded luma / chroma planes
    int me_method, me_range, subme, chroma_me, mv_range;
    int max_refs, b_cabac, b_fast_pskip, b_dct_decimate, analyse_inter, pass2_elide;
    int conformant;                  // pcamv_set_conformant: pass 2 without quirk q2, embed stage with a straight vector copy
    const uint8_t *fenc_y, *fenc_u, *fenc_v;   // source frame, pixel (0,0); strides = stride_y / stride_c
    DevRef ref[PCAMV_SLOTS];
    DevTables tab;
};

// kernel launchers (pcamv_kernels.cu); all asynchronous on `stream`
void launch_expand_border(uint8_t *p0, uint8_t *p1, uint8_t *p2, int nplanes, int stride,
                          int keep_x0, int keep_x1, int keep_y0, int keep_y1,
                          int pad_x0, int pad_x1, int pad_y0, int pad_y1, void *stream);
void launch_hpel_filter(const uint8_t *src, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                        int stride, int width, int height, void *stream);
void launch_box_sum(const uint8_t *src_padded, uint16_t *dst_padded, int stride, int rows, int box, void *stream);
void launch_search_batch(const DevFrameCtx &fc, const pcamv_me_call *calls, int n, pcamv_me_result *results,
                         unsigned long long *mvsads, int mvsads_cap, int chunk, void *stream);

void launch_int_peak(uint32_t *out, int blocks, int iters, void *stream);

} // namespace pcamv
