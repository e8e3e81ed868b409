/* pcamv_synth: deterministic synthetic I420 generator.
 *
 * No real video exists in the reference tree, so every config is driven by this generator:
 *   luma   = 3-octave value-noise texture (mean 128, sigma ~ 40, clipped) translating globally by
 *            (2,1) px/frame, wrapping over a canvas 64 px larger than the frame,
 *          + 16..24 rectangles of 24..96 px with their own velocities in [-6,6] px/frame at
 *            half-pel resolution (forces sub-partitions, non-zero mvd, occlusions),
 *          + i.i.d. noise (sigma = 2) re-drawn every frame;
 *   chroma = half-resolution texture of the same construction (not flat, so mc_chroma and the
 *            chroma-ME paths are exercised).
 * Integer arithmetic only, so the output is identical on every platform.
 * Seed = 0x5043414D56000000 | config << 8 | stream.
 *
 * usage: pcamv_synth WIDTH HEIGHT FRAMES CONFIG STREAM OUT.yuv [noise_sigma_x16=32]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state;
static inline uint64_t rng_next(void)
{
    uint64_t x = rng_state;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    rng_state = x;
    return x * 0x2545F4914F6CDD1DULL;
}
static inline uint32_t rng_u32(void) { return (uint32_t)(rng_next() >> 32); }
static inline int rng_range(int lo, int hi) { return lo + (int)(rng_u32() % (uint32_t)(hi - lo + 1)); }

/* stateless lattice hash -> [0,255] */
static inline int lattice(uint32_t seed, int x, int y)
{
    uint32_t h = seed ^ ((uint32_t)x * 0x9E3779B1u) ^ ((uint32_t)y * 0x85EBCA77u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (int)(h & 255);
}

/* bilinear value noise with lattice spacing 2^shift, period (pw,ph) in lattice cells; returns 0..255 in 8.8 */
static inline int vnoise(uint32_t seed, int x, int y, int shift, int pw, int ph)
{
    int cx = x >> shift, cy = y >> shift;
    int fx = x & ((1 << shift) - 1), fy = y & ((1 << shift) - 1);
    int one = 1 << shift;
    int x1 = (cx + 1) % pw, y1 = (cy + 1) % ph;
    int a = lattice(seed, cx % pw, cy % ph), b = lattice(seed, x1, cy % ph);
    int c = lattice(seed, cx % pw, y1),      d = lattice(seed, x1, y1);
    int top = a * (one - fx) + b * fx;
    int bot = c * (one - fx) + d * fx;
    return ((top * (one - fy) + bot * fy) << 8) >> (2 * shift);
}

/* texture sample at canvas position (x,y), canvas size (cw,ch) multiples of 64; returns 0..255 */
static inline int texture(uint32_t seed, int x, int y, int cw, int ch)
{
    /* octaves at lattice spacing 32, 8, 4 with weights 4,2,1 (x3/16): sigma ~ 40 about mean 128 */
    int v = 4 * (vnoise(seed, x, y, 5, cw >> 5, ch >> 5) - (128 << 8))
          + 2 * (vnoise(seed + 1, x, y, 3, cw >> 3, ch >> 3) - (128 << 8))
          + 1 * (vnoise(seed + 2, x, y, 2, cw >> 2, ch >> 2) - (128 << 8));
    v = 128 + ((v * 3) >> 12);
    return v < 0 ? 0 : v > 255 ? 255 : v;
}

typedef struct { int x2, y2, w, h, vx2, vy2; uint32_t seed; } rect_t;   /* positions/velocities in half-pel */

static inline int wrap(int v, int m) { v %= m; return v < 0 ? v + m : v; }

static void render_plane(uint8_t *dst, int w, int h, int frame, int sub, uint32_t seed,
                         const rect_t *rc, int nrect, int noise16)
{
    /* sub = 0 luma, 1 chroma (half resolution: positions and sizes halved) */
    int cw = ((w + 63) & ~63) + (64 >> sub), ch = ((h + 63) & ~63) + (64 >> sub);
    int gx = (2 * frame) >> sub, gy = frame >> sub;
    cw = (cw + 63) & ~63; ch = (ch + 63) & ~63;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            dst[y * w + x] = (uint8_t)texture(seed, wrap(x + gx, cw), wrap(y + gy, ch), cw, ch);
    for (int r = 0; r < nrect; r++)
    {
        int px2 = rc[r].x2 + rc[r].vx2 * frame, py2 = rc[r].y2 + rc[r].vy2 * frame;
        int rw = rc[r].w >> sub, rh = rc[r].h >> sub;
        int span_x = 2 * (w << sub) + 2 * rc[r].w, span_y = 2 * (h << sub) + 2 * rc[r].h;
        px2 = wrap(px2 + 2 * rc[r].w, span_x) - 2 * rc[r].w;
        py2 = wrap(py2 + 2 * rc[r].h, span_y) - 2 * rc[r].h;
        if (sub) { px2 >>= 1; py2 >>= 1; }      /* chroma moves at half the luma displacement */
        {
            int ix = px2 >> 1, iy = py2 >> 1, fx = px2 & 1, fy = py2 & 1;
            for (int y = 0; y < rh + fy; y++)
                for (int x = 0; x < rw + fx; x++)
                {
                    int X = ix + x, Y = iy + y;
                    if (X < 0 || Y < 0 || X >= w || Y >= h) continue;
                    /* half-pel placement = average of the rectangle texture at the two/four
                     * straddled integer positions (edge pixels blend with the rectangle itself) */
                    int x0 = x - fx < 0 ? 0 : x - fx, y0 = y - fy < 0 ? 0 : y - fy;
                    int x1 = x >= rw ? rw - 1 : x, y1 = y >= rh ? rh - 1 : y;
                    int a = texture(rc[r].seed, x0 + 7, y0 + 3, 256, 256), b = texture(rc[r].seed, x1 + 7, y0 + 3, 256, 256);
                    int c = texture(rc[r].seed, x0 + 7, y1 + 3, 256, 256), d = texture(rc[r].seed, x1 + 7, y1 + 3, 256, 256);
                    dst[Y * w + X] = (uint8_t)((a + b + c + d + 2) >> 2);
                }
        }
    }
    if (noise16 > 0)
        for (int i = 0; i < w * h; i++)
        {
            uint32_t r = rng_u32();
            int n = (int)((r & 255) + ((r >> 8) & 255) + ((r >> 16) & 255) + (r >> 24)) - 510;   /* sigma 147.8 */
            int v = dst[i] + (n * noise16 + (n >= 0 ? 1182 : -1182)) / 2365;                        /* sigma = noise16/16 */
            dst[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
}

int main(int argc, char **argv)
{
    if (argc < 7)
    {
        fprintf(stderr, "usage: %s WIDTH HEIGHT FRAMES CONFIG STREAM OUT.yuv [noise_sigma_x16=32]\n", argv[0]);
        return 2;
    }
    int w = atoi(argv[1]), h = atoi(argv[2]), frames = atoi(argv[3]);
    int config = atoi(argv[4]), stream = atoi(argv[5]);
    int noise16 = argc > 7 ? atoi(argv[7]) : 32;
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1) || frames <= 0)
    {
        fprintf(stderr, "pcamv_synth: bad geometry\n");
        return 2;
    }
    FILE *f = fopen(argv[6], "wb");
    if (!f) { perror(argv[6]); return 1; }
    rng_state = 0x5043414D56000000ULL | ((uint64_t)(config & 0xffff) << 8) | (uint64_t)(stream & 255);
    for (int i = 0; i < 8; i++) rng_next();
    uint32_t seed_y = rng_u32(), seed_u = rng_u32(), seed_v = rng_u32();
    int nrect = 16 + (w * h) / (352 * 288 * 4);
    if (nrect > 64) nrect = 64;
    rect_t rc[64];
    for (int r = 0; r < nrect; r++)
    {
        rc[r].w = rng_range(24, 96) & ~1; rc[r].h = rng_range(24, 96) & ~1;
        rc[r].x2 = 2 * rng_range(0, w - 1); rc[r].y2 = 2 * rng_range(0, h - 1);
        rc[r].vx2 = rng_range(-12, 12); rc[r].vy2 = rng_range(-12, 12);
        rc[r].seed = rng_u32();
    }
    uint8_t *Y = malloc((size_t)w * h), *U = malloc((size_t)w * h / 4), *V = malloc((size_t)w * h / 4);
    for (int t = 0; t < frames; t++)
    {
        rect_t ru[64], rv[64];
        for (int r = 0; r < nrect; r++) { ru[r] = rc[r]; ru[r].seed ^= 0x55u; rv[r] = rc[r]; rv[r].seed ^= 0xAAu; }
        render_plane(Y, w, h, t, 0, seed_y, rc, nrect, noise16);
        render_plane(U, w / 2, h / 2, t, 1, seed_u, ru, nrect, noise16);
        render_plane(V, w / 2, h / 2, t, 1, seed_v, rv, nrect, noise16);
        fwrite(Y, 1, (size_t)w * h, f);
        fwrite(U, 1, (size_t)w * h / 4, f);
        fwrite(V, 1, (size_t)w * h / 4, f);
    }
    fclose(f);
    free(Y); free(U); free(V);
    return 0;
}
