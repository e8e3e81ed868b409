/* TEST INFRASTRUCTURE ONLY (oracle/): link shim for the un-vendored S-UNIWARD.lib symbol.
 *
 * The reference declares `extern * get_cost_lib_for_x264(int w,int h,int*mv_h,int*mv_v)`
 * (reference encoder/encoder.c:38) and calls it at encoder/encoder.c:1441.  The library
 * is not in the tree (build/win32/x264_vs2008.vcxproj:87,126).  Its result is weighted by
 * alpha_com = 0 (encoder/encoder.c:1651-1652) and free()d (encoder/encoder.c:1823), so a
 * zero-filled malloc-family buffer is result-neutral. */
#include <stdlib.h>

float *get_cost_lib_for_x264(int w, int h, int *mv_h, int *mv_v)
{
    (void)mv_h; (void)mv_v;
    return (float *)calloc((size_t)w * (size_t)h, sizeof(float));
}
