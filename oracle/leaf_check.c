/* TEST INFRASTRUCTURE ONLY (oracle/): pins oracle/leaf_oracle.c against the reference's own leaf functions.
 * Links oracle/_ref/libx264_wide.a (the reference's C sources compiled by oracle/build_ref.py) and compares, on
 * random and saturating buffers (tools/checkasm.c:246-256,1528-1534 pattern), the function tables filled by
 * x264_pixel_init(0, ...) (common/pixel.c:565) and x264_mc_init(0, ...) (common/mc.c:406) with the restatement.
 * Prints "checks=N mismatches=M"; exit code 1 on any mismatch. */
#include "common/common.h"
#include <stdio.h>

int pcamv_oracle_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
int pcamv_oracle_satd(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
void pcamv_oracle_mc_luma(uint8_t *dst, int dst_stride, uint8_t *const src[4], int stride, int mvx, int mvy, int w, int h);
void pcamv_oracle_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int stride, int mvx, int mvy, int w, int h);

void pcamv_oracle_integral8(const uint8_t *plane, uint16_t *sum8, int stride, int rows);
int pcamv_oracle_ads(const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx, int16_t *mvs, int width,
                     int thresh, int n_dc);

static uint64_t rng = 0x5043414D56ULL;
static uint32_t rnd(void) { rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27; return (uint32_t)((rng * 2685821657736338717ULL) >> 32); }

int main(void)
{
    static const int pw[7] = { 16, 16, 8, 8, 8, 4, 4 }, ph[7] = { 16, 8, 16, 8, 4, 8, 4 };
    enum { S = 128, N = S * 96 };
    static uint8_t a[N] __attribute__((aligned(16))), b[4][N], d1[32 * 32], d2[32 * 32];
    x264_pixel_function_t pixf; x264_mc_functions_t mc;
    long checks = 0, bad = 0;
    int it, i, k;
    x264_pixel_init(0, &pixf);
    x264_mc_init(0, &mc);
    for (it = 0; it < 400; it++)
    {
        const int mode = it % 4;          /* 0,1: random; 2: saturating (0 vs 255); 3: near-equal */
        for (i = 0; i < N; i++)
        {
            a[i] = mode == 2 ? (uint8_t)(rnd() & 1 ? 255 : 0) : (uint8_t)rnd();
            for (k = 0; k < 4; k++)
                b[k][i] = mode == 2 ? (uint8_t)(rnd() & 1 ? 255 : 0) : mode == 3 ? (uint8_t)(a[i] + (rnd() % 5) - 2) : (uint8_t)rnd();
        }
        for (i = 0; i < 7; i++)
        {
            const int off = 8 * S + 16 + (int)(rnd() % 40);        /* second operand unaligned, as in the search */
            checks += 2;
            bad += pixf.sad[i](a, 16, b[0] + off, S) != pcamv_oracle_sad(a, 16, b[0] + off, S, pw[i], ph[i]);
            bad += pixf.satd[i](a, 16, b[0] + off, S) != pcamv_oracle_satd(a, 16, b[0] + off, S, pw[i], ph[i]);
        }
        for (k = 0; k < 24; k++)
        {
            uint8_t *src[4] = { b[0] + 40 * S + 48, b[1] + 40 * S + 48, b[2] + 40 * S + 48, b[3] + 40 * S + 48 };
            const int mvx = (int)(rnd() % 65) - 32, mvy = (int)(rnd() % 65) - 32, i_pix = (int)(rnd() % 7);
            memset(d1, 0, sizeof(d1)); memset(d2, 0, sizeof(d2));
            mc.mc_luma(d1, 32, src, S, mvx, mvy, pw[i_pix], ph[i_pix]);
            pcamv_oracle_mc_luma(d2, 32, src, S, mvx, mvy, pw[i_pix], ph[i_pix]);
            checks++; bad += memcmp(d1, d2, sizeof(d1)) != 0;
            memset(d1, 0, sizeof(d1)); memset(d2, 0, sizeof(d2));
            mc.mc_chroma(d1, 32, src[0], S, mvx, mvy, pw[i_pix] / 2, ph[i_pix] / 2);
            pcamv_oracle_mc_chroma(d2, 32, src[0], S, mvx, mvy, pw[i_pix] / 2, ph[i_pix] / 2);
            checks++; bad += memcmp(d1, d2, sizeof(d1)) != 0;
        }
    }
    /* integral plane (mc.c:311-345 driven as in x264_frame_filter, mc.c:489-509) and the ads prefilters (pixel.c:515-559) */
    {
        enum { IS = 96, IR = 72 };
        static uint8_t pl[IS * IR + 64];
        static uint16_t s_ref[IS * (IR + 1) + 64], s_or[IS * (IR + 1) + 64], cost[IS];
        int y, x;
        for (it = 0; it < 40; it++)
        {
            for (i = 0; i < IS * IR; i++) pl[i] = it % 3 == 2 ? 255 : (uint8_t)rnd();
            memset(s_ref, 0, sizeof(s_ref)); memset(s_or, 0, sizeof(s_or));
            for (y = 0; y < IR - 1; y++)
            {
                mc.integral_init8h(s_ref + (y + 1) * IS, pl + y * IS, IS);
                if (y >= 7) mc.integral_init8v(s_ref + (y + 1) * IS - 8 * IS, IS);
            }
            pcamv_oracle_integral8(pl, s_or, IS, IR);
            for (y = 0; y <= IR - 9; y++)
                for (x = 0; x < IS - 8; x++)
                {
                    int box = 0, yy, xx;
                    for (yy = 0; yy < 8; yy++) for (xx = 0; xx < 8; xx++) box += pl[(y + yy) * IS + x + xx];
                    checks++; bad += s_ref[y * IS + x] != s_or[y * IS + x] || s_or[y * IS + x] != box;
                }
            for (k = 0; k < 30; k++)
            {
                int dc[4], n1, n2, th = (int)(rnd() % 6000), w = 4 * (1 + (int)(rnd() % 8));
                int16_t m1[64], m2[64];
                const int which = k % 3, pix = which == 0 ? PIXEL_16x16 : which == 1 ? PIXEL_16x8 : PIXEL_8x8;
                const int delta = which == 0 ? 8 * IS : 8;
                for (i = 0; i < 4; i++) dc[i] = (int)(rnd() % 16321);
                for (i = 0; i < IS; i++) cost[i] = (uint16_t)(rnd() % 200);
                n1 = pixf.ads[pix](dc, s_or + 10 * IS + 5, delta, cost, m1, w, th);
                n2 = pcamv_oracle_ads(dc, s_or + 10 * IS + 5, delta, cost, m2, w, th, which == 0 ? 4 : which == 1 ? 2 : 1);
                checks++; bad += n1 != n2 || memcmp(m1, m2, n1 * sizeof(int16_t)) != 0;
            }
        }
    }
    printf("checks=%ld mismatches=%ld\n", checks, bad);
    return bad ? 1 : 0;
}
