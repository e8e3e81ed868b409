// emu_intra_check.cpp — csrc/pcamv_intra.cuh on the CPU against the reference's intra analysis.
// Input: a dump of oracle/_ref/x264_dump_rd.  Every 'INTR' record is one call of x264_mb_analyse_intra (encoder/analyse.c:628-879),
// I and P slices: the source macroblock, the reconstructed border of its neighbours, which neighbours exist, lambda, and what the
// analysis decided for the 16x16 luma modes (cost of every mode, best mode) and - where it had run - the chroma modes.
// Test infrastructure; prints key=value pairs.
#define PCAMV_EMU 1
#include "../../video-steganography-pcamv_b200/csrc/pcamv_device.h"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_frame.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_intra.cuh"
#include "dump_reader.h"

using namespace pcamv;

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s dump.bin\n", argv[0]); return 2; }
    Dump d;
    if (!d.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    long n = 0, bad16 = 0, n_c = 0, bad_c = 0, n_islice = 0, n_border = 0, n_sad = 0, n4 = 0, n4_done = 0, bad4 = 0;
    for (const DumpRec &r : d.recs)
    {
        if (strcmp(r.tag, "INTR")) continue;
        int32_t hd[20]; memcpy(hd, r.data, sizeof(hd));
        const uint8_t *pix = r.data + sizeof(hd), *border = pix + 256 + 128;
        const int has_left = hd[4], has_top = hd[5], has_topleft = hd[6], lambda = hd[7];
        if (d.cfg[6] < 2) { n_sad++; continue; }                 // subme < 2: the reference compares with SAD (mbcmp), not covered here
        IntraCosts o;
        intra_analyse_16x16(pix, has_left, has_top, has_topleft, border[0], border + 1, border + 17, lambda, o);
        n++; n_islice += hd[3] == 2; n_border += !has_topleft;
        bool ok = o.satd16 == hd[9] && o.pred16 == hd[10];
        // the four per-mode costs the reference keeps are indexed by mode; only evaluated modes are defined
        if (has_topleft) for (int m = 0; m < 4; m++) ok = ok && o.dir16[m] == hd[11 + m];
        if (!ok)
        {
            if (bad16 < 5) fprintf(stderr, "frame %d mb %d (slice %d, left %d top %d topleft %d): 16x16 cost %d mode %d [%d %d %d %d], reference %d mode %d [%d %d %d %d]\n",
                                   hd[0], hd[2], hd[3], has_left, has_top, has_topleft, o.satd16, o.pred16, o.dir16[0], o.dir16[1], o.dir16[2], o.dir16[3],
                                   hd[9], hd[10], hd[11], hd[12], hd[13], hd[14]);
            bad16++;
        }
        // the 4x4 modes: cost (or "gave up") and the modes of the blocks the analysis got to
        // (P slices take the flags from analyse.inter, cfg[12]; I slices from analyse.intra, whose default has the 4x4 modes on)
        if (hd[3] == 2 || (d.cfg[12] & 0x01))
        {
            const uint8_t *q = border + 33 + 34;
            int32_t p4[16]; memcpy(p4, q, 64); q += 64;
            int32_t x4[4]; memcpy(x4, q, 16); q += 16;
            Intra4x4In in;
            memcpy(in.nb4, q, 16); q += 16;
            memcpy(in.left_mode, q, 4); q += 4; memcpy(in.top_mode, q, 4); q += 4;
            const uint8_t *tr = q; q += 4;
            alignas(16) uint16_t mf[16], bias[16]; alignas(16) int32_t dq[96];
            memcpy(mf, q, 32); q += 32; memcpy(bias, q, 32); q += 32; memcpy(dq, q, 384);
            alignas(16) uint8_t fenc[256], buf[17 * 32];
            memcpy(fenc, pix, 256); memset(buf, 0, sizeof(buf));
            buf[3] = border[0];
            for (int i = 0; i < 16; i++) { buf[4 + i] = border[1 + i]; buf[(i + 1) * 32 + 3] = border[17 + i]; }
            for (int i = 0; i < 4; i++) buf[20 + i] = tr[i];
            in.fenc = fenc; in.buf = buf; in.lambda = lambda; in.qp = x4[0]; in.mbrd = x4[1]; in.fast_intra = x4[2]; in.satd8x8 = x4[3];
            in.satd_inter = hd[8]; in.satd16 = hd[9]; in.quant_mf = mf; in.quant_bias = bias; in.dequant_mf = dq;
            // the reference leaves the 4x4 analysis out entirely when fast intra decides so after the 16x16 modes (analyse.c:686)
            const bool skipped = in.fast_intra && in.satd16 > 2 * in.satd_inter;
            if (!skipped)
            {
                int pred[16]; for (int i = 0; i < 16; i++) pred[i] = -99;
                const int c4 = intra_analyse_4x4(in, pred);
                n4++; n4_done += c4 < (1 << 28);
                bool ok4 = c4 == hd[17] || (c4 >= (1 << 28) && hd[17] >= (1 << 28));
                for (int i = 0; i < 16 && ok4; i++) if (pred[i] != -99 && pred[i] != p4[i]) ok4 = false;
                if (!ok4)
                {
                    if (bad4 < 5)
                    {
                        fprintf(stderr, "frame %d mb %d (slice %d): 4x4 cost %d, reference %d; modes", hd[0], hd[2], hd[3], c4, hd[17]);
                        for (int i = 0; i < 16; i++) fprintf(stderr, " %d/%d", pred[i], p4[i]);
                        fprintf(stderr, "\n");
                    }
                    bad4++;
                }
            }
        }
        if (hd[15] < (1 << 28))                                  // the chroma analysis had run before this call
        {
            const uint8_t *bu = border + 33, *bv = border + 50;
            const int tl[2] = { bu[0], bv[0] };
            intra_analyse_chroma(pix + 256, pix + 320, has_left, has_top, has_topleft, tl, bu + 1, bu + 9, bv + 1, bv + 9, lambda, o);
            n_c++;
            if (o.satd_c != hd[15] || o.pred_c != hd[16])
            {
                if (bad_c < 5) fprintf(stderr, "frame %d mb %d: chroma cost %d mode %d, reference %d mode %d\n", hd[0], hd[2], o.satd_c, o.pred_c, hd[15], hd[16]);
                bad_c++;
            }
        }
    }
    printf("luma16x16=%ld bad16=%ld chroma=%ld bad_chroma=%ld luma4x4=%ld luma4x4_completed=%ld bad4x4=%ld in_i_slices=%ld at_picture_border=%ld sad_skipped=%ld\n",
           n, bad16, n_c, bad_c, n4, n4_done, bad4, n_islice, n_border, n_sad);
    return (bad16 || bad_c || bad4) ? 1 : 0;
}
