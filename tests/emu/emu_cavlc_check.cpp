// emu_cavlc_check.cpp — csrc/pcamv_cavlc.cuh on the CPU against the reference's RD mode decision.
//
// Input: a dump of oracle/_ref/x264_dump_rd run with --subme 6 --no-cabac.  Every 'RDMB' record is one call of
// x264_rd_cost_mb (encoder/rdo.c:139-172) that sized a macroblock with x264_macroblock_size_cavlc: the macroblock as that
// function saw it (type, partitioning, references, vector differences, cbp, quantised coefficients, the coefficient-count
// cache) and the bit count it returned.  For inter macroblocks cavlc_mb_inter_bits must return the same count; intra
// candidates (x264_intra_rd runs in P slices although intra is never chosen there) are counted and skipped.
// Code lengths come from the 'VLC0' record (the encoder's own tables).  Test infrastructure; prints key=value pairs.
#include "../../video-steganography-pcamv_b200/csrc/pcamv_cavlc.cuh"
#include "dump_reader.h"

using namespace pcamv;

static const int scan8[24] = { 4 + 1 * 8, 5 + 1 * 8, 4 + 2 * 8, 5 + 2 * 8, 6 + 1 * 8, 7 + 1 * 8, 6 + 2 * 8, 7 + 2 * 8,
                               4 + 3 * 8, 5 + 3 * 8, 4 + 4 * 8, 5 + 4 * 8, 6 + 3 * 8, 7 + 3 * 8, 6 + 4 * 8, 7 + 4 * 8,
                               1 + 1 * 8, 2 + 1 * 8, 1 + 2 * 8, 2 + 2 * 8, 1 + 4 * 8, 2 + 4 * 8, 1 + 5 * 8, 2 + 5 * 8 };   // common/common.h:217-231

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s dump.bin\n", argv[0]); return 2; }
    Dump d;
    if (!d.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const DumpRec *v = d.find("VLC0");
    if (!v || v->size != sizeof(CavlcSizes)) { fprintf(stderr, "dump lacks VLC0 (not made by x264_dump_rd --subme 6 --no-cabac?)\n"); return 2; }
    CavlcSizes z; memcpy(&z, v->data, sizeof(z));
    long n_inter = 0, n_intra = 0, bad = 0, bits_sum = 0, n_p8x8 = 0, n_sub = 0, n_multi_ref = 0, n_coded = 0;
    for (const DumpRec &r : d.recs)
    {
        if (strcmp(r.tag, "RDMB")) continue;
        int32_t hd[20]; memcpy(hd, r.data, sizeof(hd));
        const uint8_t *p = r.data + sizeof(hd);
        CavlcMb m; memset(&m, 0, sizeof(m));
        m.type = hd[3]; m.partition = hd[4];
        for (int i = 0; i < 4; i++) m.sub[i] = hd[5 + i];
        m.n_ref = hd[9]; m.psub8x8 = hd[10]; m.cbp_luma = hd[11]; m.cbp_chroma = hd[12]; m.qp_delta = hd[13];
        const int want = hd[15];
        m.n_mvd = hd[17];
        memcpy(m.ref, p, 4); p += 4;
        memcpy(m.mvd, p, sizeof(m.mvd)); p += sizeof(m.mvd);
        const uint8_t *nnz = p; p += 48;
        m.coef = (const int16_t (*)[16])p; p += 24 * 16 * 2;
        m.chroma_dc = (const int16_t (*)[4])p;
        if (m.type != 4 && m.type != 5) { n_intra++; continue; }
        // left column / top row of the coefficient-count cache (h->mb.cache.non_zero_count, x264_scan8 layout)
        for (int y = 0; y < 4; y++) m.nnz_left[y] = nnz[scan8[0] - 1 + 8 * y];
        for (int x = 0; x < 4; x++) m.nnz_top[x] = nnz[scan8[0] - 8 + x];
        for (int pl = 0; pl < 2; pl++)
            for (int k = 0; k < 2; k++)
            {
                m.nnz_left_c[pl][k] = nnz[scan8[16 + 4 * pl] - 1 + 8 * k];
                m.nnz_top_c[pl][k] = nnz[scan8[16 + 4 * pl] - 8 + k];
            }
        for (int i = 0; i < 24; i++) m.coded[i] = nnz[scan8[i]];
        m.coded[24] = nnz[5 + 5 * 8]; m.coded[25] = nnz[6 + 5 * 8];                      // x264_scan8[25], [26]: chroma DC
        const int got = cavlc_mb_inter_bits(z, m);
        n_inter++; bits_sum += want;
        n_p8x8 += m.type == 5;
        n_sub += m.type == 5 && (m.sub[0] != 3 || m.sub[1] != 3 || m.sub[2] != 3 || m.sub[3] != 3);
        n_multi_ref += m.n_ref > 1;
        n_coded += (m.cbp_luma | m.cbp_chroma) != 0;
        if (got != want)
        {
            if (bad < 5)
                fprintf(stderr, "frame %d mb %d type %d partition %d cbp %x/%d: %d bits, reference %d\n", hd[0], hd[2], m.type, m.partition, m.cbp_luma, m.cbp_chroma, got, want);
            bad++;
        }
    }
    printf("inter=%ld bad=%ld intra_skipped=%ld bits=%ld p8x8=%ld sub8x8=%ld multi_ref=%ld coded=%ld\n", n_inter, bad, n_intra, bits_sum, n_p8x8, n_sub, n_multi_ref, n_coded);
    return bad ? 1 : 0;
}
