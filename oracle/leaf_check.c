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
    printf("checks=%ld mismatches=%ld\n", checks, bad);
    return bad ? 1 : 0;
}
