/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C restatement of the leaf arithmetic of the motion-estimation path.
 *
 * Scalar, loop-per-pixel code written from the reference's definitions (file:line of /root/reference); nothing here
 * is shipped or called by the product — tests/ load it (oracle/_ref/libpcamv_oracle.so) as the checker for the CUDA
 * kernels on inputs where no reference dump exists (random and saturating buffers, tools/checkasm.c:246-256 style).
 * Pinned two ways: tests/test_oracle_leaf.py compares it with the reference's own function tables
 * (x264_pixel_init / x264_mc_init of oracle/_ref/libx264_wide.a, built from the reference sources) and with the
 * reference encoder's planes in the golden dumps (tests/golden).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }

/* common/pixel.c:40-65  x264_pixel_sad_WxH */
int pcamv_oracle_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int s = 0, x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            s += abs(a[y * sa + x] - b[y * sb + x]);
    return s;
}

/* common/pixel.c:187-253  x264_pixel_satd_WxH: sum over 4x4 blocks of |H4 * D * H4| / 2.  The reference packs two 4x4
 * blocks per 8x4 call and halves their joint sum; every coefficient of one block has the parity of the block's pixel
 * sum, so each block's absolute sum is even and halving per block is the same number. */
int pcamv_oracle_satd(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int total = 0, bx, by, i, j;
    for (by = 0; by < h; by += 4)
        for (bx = 0; bx < w; bx += 4)
        {
            int d[4][4], t[4][4], s = 0;
            for (i = 0; i < 4; i++)
                for (j = 0; j < 4; j++)
                    d[i][j] = a[(by + i) * sa + bx + j] - b[(by + i) * sb + bx + j];
            for (i = 0; i < 4; i++)
            {
                const int s01 = d[i][0] + d[i][1], d01 = d[i][0] - d[i][1], s23 = d[i][2] + d[i][3], d23 = d[i][2] - d[i][3];
                t[i][0] = s01 + s23; t[i][1] = s01 - s23; t[i][2] = d01 + d23; t[i][3] = d01 - d23;
            }
            for (j = 0; j < 4; j++)
            {
                const int s01 = t[0][j] + t[1][j], d01 = t[0][j] - t[1][j], s23 = t[2][j] + t[3][j], d23 = t[2][j] - t[3][j];
                s += abs(s01 + s23) + abs(s01 - s23) + abs(d01 + d23) + abs(d01 - d23);
            }
            total += s >> 1;
        }
    return total;
}

/* common/frame.c:224-244 plane_expand_border, closed form: pixel (x,y) outside the kept rectangle takes the value of the
 * nearest kept pixel (left/right bands first, then whole rows up and down — i.e. coordinate clamping). */
static void expand(uint8_t *p, int stride, int kx0, int kx1, int ky0, int ky1, int px0, int px1, int py0, int py1)
{
    int x, y;
    for (y = py0; y < py1; y++)
    {
        const int cy = y < ky0 ? ky0 : y >= ky1 ? ky1 - 1 : y;
        for (x = px0; x < px1; x++)
        {
            const int cx = x < kx0 ? kx0 : x >= kx1 ? kx1 - 1 : x;
            if (cx != x || cy != y)
                p[y * stride + x] = p[cy * stride + cx];
        }
    }
}

/* chroma plane of a reference frame: 16-pixel border (common/frame.c:246-273, PADH/PADV halved) */
void pcamv_oracle_chroma_border(uint8_t *p, int stride, int w, int h)
{
    expand(p, stride, 0, w, 0, h, -16, w + 16, -16, h + 16);
}

#define TAP6(p, x, d) ((p)[(x) - 2 * (d)] + (p)[(x) + 3 * (d)] - 5 * ((p)[(x) - (d)] + (p)[(x) + 2 * (d)]) + 20 * ((p)[(x)] + (p)[(x) + (d)]))

/* The four luma planes of a reference frame from its reconstructed interior, as x264_fdec_filter_row leaves them:
 * integer plane with 32-pixel border (common/frame.c:246-273); H, V, HV by the 6-tap filter over [-8,W+8) x [-8,H+8)
 * (common/mc.c:134-190 hpel_filter called with offs = -8 rows - 8 columns, width W+16, mc.c:453-475; the vertical
 * intermediate is kept as int16 for the centre plane); filtered planes keep x in [-4,W+4), y in [-8,H+8) and are
 * replicated outward from there (common/frame.c:275-301: padh 28, padv 24).
 * planes[k] point at pixel (0,0) of buffers with 32 pixels of border on every side. */
void pcamv_oracle_frame_planes(uint8_t *const planes[4], int stride, int w, int h)
{
    uint8_t *src = planes[0], *dsth = planes[1], *dstv = planes[2], *dstc = planes[3];
    int16_t *buf = (int16_t *)malloc((size_t)(w + 16 + 5) * sizeof(int16_t));
    int x, y;
    expand(src, stride, 0, w, 0, h, -32, w + 32, -32, h + 32);
    for (y = -8; y < h + 8; y++)
    {
        const uint8_t *s = src + y * stride;
        for (x = -10; x < w + 11; x++)
        {
            const int v = TAP6(s, x, stride);
            dstv[y * stride + x] = (uint8_t)clip_u8((v + 16) >> 5);
            buf[x + 10] = (int16_t)v;
        }
        for (x = -8; x < w + 8; x++)
        {
            dstc[y * stride + x] = (uint8_t)clip_u8((TAP6(buf + 10, x, 1) + 512) >> 10);
            dsth[y * stride + x] = (uint8_t)clip_u8((TAP6(s, x, 1) + 16) >> 5);
        }
    }
    free(buf);
    for (x = 1; x < 4; x++)
        expand(planes[x], stride, -4, w + 4, -8, h + 8, -32, w + 32, -32, h + 32);
}

/* common/mc.c:192-243 mc_luma / get_ref: quarter-pel sample = rounded average of two of the four planes */
void pcamv_oracle_mc_luma(uint8_t *dst, int dst_stride, uint8_t *const src[4], int stride, int mvx, int mvy, int w, int h)
{
    static const int ref0[16] = { 0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1 };
    static const int ref1[16] = { 0, 0, 0, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2 };
    const int idx = ((mvy & 3) << 2) + (mvx & 3);
    const int off = (mvy >> 2) * stride + (mvx >> 2);
    const uint8_t *s1 = src[ref0[idx]] + off + ((mvy & 3) == 3) * stride;
    const uint8_t *s2 = src[ref1[idx]] + off + ((mvx & 3) == 3);
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            dst[y * dst_stride + x] = (idx & 5) ? (uint8_t)((s1[y * stride + x] + s2[y * stride + x] + 1) >> 1) : s1[y * stride + x];
}

/* common/mc.c:246-277 mc_chroma: bilinear 1/8-pel */
void pcamv_oracle_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int stride, int mvx, int mvy, int w, int h)
{
    const int dx = mvx & 7, dy = mvy & 7;
    const int cA = (8 - dx) * (8 - dy), cB = dx * (8 - dy), cC = (8 - dx) * dy, cD = dx * dy;
    const uint8_t *s = src + (mvy >> 3) * stride + (mvx >> 3);
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            dst[y * dst_stride + x] = (uint8_t)((cA * s[y * stride + x] + cB * s[y * stride + x + 1] +
                                                 cC * s[(y + 1) * stride + x] + cD * s[(y + 1) * stride + x + 1] + 32) >> 6);
}

/* common/mc.c:311-345 integral_init8h / integral_init8v as x264_frame_filter runs them (mc.c:477-511): a running
 * column-prefix of horizontal 8-pixel sums in uint16 (wrapping), turned into 8x8 box sums eight rows behind.
 * plane / sum8: top-left of the PADDED buffers (stride x rows).  On return sum8[y * stride + x] holds the sum of the
 * 8x8 pixels whose top-left corner is (x, y) for x < stride - 8, y <= rows - 9; the eight rows below keep running
 * prefixes, as in the reference's buffer. */
void pcamv_oracle_integral8(const uint8_t *plane, uint16_t *sum8, int stride, int rows)
{
    int x, y;
    memset(sum8, 0, (size_t)stride * sizeof(uint16_t));            /* the zero row above the first prefix row */
    for (y = 0; y < rows - 1; y++)
    {
        const uint8_t *pix = plane + (size_t)y * stride;
        uint16_t *sum = sum8 + (size_t)(y + 1) * stride;
        int v = pix[0] + pix[1] + pix[2] + pix[3] + pix[4] + pix[5] + pix[6] + pix[7];
        for (x = 0; x < stride - 8; x++)
        {
            sum[x] = (uint16_t)(v + sum[x - stride]);
            v += pix[x + 8] - pix[x];
        }
        if (y >= 7)
        {
            uint16_t *s = sum8 + (size_t)(y - 7) * stride;         /* integral_init8v on the row 8 above the newest prefix */
            for (x = 0; x < stride - 8; x++)
                s[x] = (uint16_t)(s[x + 8 * stride] - s[x]);
        }
    }
}

/* common/pixel.c:515-559 x264_pixel_ads4 / ads2 / ads1 (n_dc = 4 / 2 / 1): candidates whose DC lower bound + MV cost stays
 * below thresh, in increasing x */
int pcamv_oracle_ads(const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx, int16_t *mvs, int width,
                     int thresh, int n_dc)
{
    int nmv = 0, i;
    for (i = 0; i < width; i++, sums++)
    {
        int ads = abs(enc_dc[0] - sums[0]) + cost_mvx[i];
        if (n_dc == 2) ads += abs(enc_dc[1] - sums[delta]);
        if (n_dc == 4) ads += abs(enc_dc[1] - sums[8]) + abs(enc_dc[2] - sums[delta]) + abs(enc_dc[3] - sums[delta + 8]);
        if (ads < thresh)
            mvs[nmv++] = (int16_t)i;
    }
    return nmv;
}
