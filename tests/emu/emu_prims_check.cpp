// CPU check of the packed-pixel primitives (pcamv_prims.cuh / pcamv_me.cuh, compiled with -DPCAMV_EMU) against the plain-C
// leaf oracle (oracle/leaf_oracle.c, itself pinned to the reference's function tables): 4x4 SATD through the packed-pair
// Hadamard, packed SAD, rounding average and the two-pixels-per-multiply chroma interpolation, on random, near-equal and
// saturating (0 vs 255) inputs — the ranges where a packed 16-bit half could overflow.
// build: g++ -O2 tests/emu/emu_prims_check.cpp oracle/leaf_oracle.c ; prints "checks=N mismatches=M"
#define PCAMV_EMU 1
#include "../../video-steganography-pcamv_b200/csrc/pcamv_me.cuh"
#include <stdio.h>

extern "C" int pcamv_oracle_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
extern "C" int pcamv_oracle_satd(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
extern "C" void pcamv_oracle_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int stride, int mvx, int mvy, int w, int h);

using namespace pcamv;

static uint64_t rng = 0x5043414D56ULL;
static uint32_t rnd() { rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27; return (uint32_t)((rng * 2685821657736338717ULL) >> 32); }

int main()
{
    long checks = 0, bad = 0;
    for (int it = 0; it < 400000; it++)
    {
        const int mode = it % 5;      // 0,1 random; 2 saturating; 3 near-equal; 4 one side constant
        uint8_t f[16], a[16];
        for (int i = 0; i < 16; i++)
        {
            f[i] = mode == 2 ? (rnd() & 1 ? 255 : 0) : mode == 4 ? 255 : (uint8_t)rnd();
            a[i] = mode == 2 ? (rnd() & 1 ? 255 : 0) : mode == 3 ? (uint8_t)(f[i] + rnd() % 5 - 2) : mode == 4 ? (rnd() & 3 ? 0 : 255) : (uint8_t)rnd();
        }
        uint32_t fw[4], aw[4];
        for (int r = 0; r < 4; r++) { memcpy(&fw[r], f + 4 * r, 4); memcpy(&aw[r], a + 4 * r, 4); }
        checks++; bad += (int)(hadamard_4x4_sum(fw, aw) >> 1) != pcamv_oracle_satd(f, 4, a, 4, 4, 4);
        int s = 0; for (int r = 0; r < 4; r++) s += sad4(fw[r], aw[r]);
        checks++; bad += s != pcamv_oracle_sad(f, 4, a, 4, 4, 4);
        const uint32_t av = avg4(fw[0], aw[0]);
        for (int i = 0; i < 4; i++) { checks++; bad += px(av, i) != ((f[i] + a[i] + 1) >> 1); }
        // chroma: 4 pixels of one row from a 2-row, 8-byte-wide source at a random 1/8-pel vector
        if (it % 4 == 0)
        {
            uint8_t src[3 * 16], want[4];
            for (int i = 0; i < 48; i++) src[i] = mode == 2 ? (rnd() & 1 ? 255 : 0) : (uint8_t)rnd();
            const int mvx = (int)(rnd() % 16), mvy = (int)(rnd() % 8);      // integer part 0..1 / 0, fraction 0..7
            pcamv_oracle_mc_chroma(want, 4, src, 16, mvx, mvy, 4, 1);
            const uint32_t got = chroma4(src, 16, mvx, mvy, 0, 0);
            for (int i = 0; i < 4; i++) { checks++; bad += px(got, i) != want[i]; }
        }
    }
    printf("checks=%ld mismatches=%ld\n", checks, bad);
    return bad ? 1 : 0;
}
