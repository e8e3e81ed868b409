// emu_rd_check.cpp — both halves of x264_rd_cost_mb (encoder/rdo.c:139-172) on the CPU with the device code.
//
// Input: a dump of oracle/_ref/x264_dump_rd run with --subme 6 --no-cabac and planes.  Every 'RDMB' record of an inter candidate
// (P_L0 / P_8x8) is one call of x264_rd_cost_mb: the candidate's type, partitioning, references and vectors, what the reference
// returned for its distortion (ssd_mb incl. the psy term) and its CAVLC size.  This checker
//   (1) reconstructs the candidate with the product's own device code (recon_mb, csrc/pcamv_recon.cuh: motion compensation +
//       DCT / quantisation / decimation / dequantisation / IDCT — what the reference frame is built with) from the dumped source
//       and reference planes, and computes the distortion of THAT reconstruction with csrc/pcamv_rd.cuh;
//   (2) checks that the luma blocks the device kept coefficients for are the blocks the reference counts coefficients for;
//   (3) sizes the candidate with csrc/pcamv_cavlc.cuh from the reference's coefficient arrays (as tests/emu/emu_cavlc_check.cpp).
// distortion and bits must both equal the reference's for every candidate.  Test infrastructure; prints key=value pairs.
#define PCAMV_EMU 1
#include "../../video-steganography-pcamv_b200/csrc/pcamv_device.h"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_recon.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_cavlc.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_rd.cuh"
#include "dump_reader.h"
#include <vector>

using namespace pcamv;

static const int x264_scan8_tab[24] = { 4 + 1 * 8, 5 + 1 * 8, 4 + 2 * 8, 5 + 2 * 8, 6 + 1 * 8, 7 + 1 * 8, 6 + 2 * 8, 7 + 2 * 8,
                               4 + 3 * 8, 5 + 3 * 8, 4 + 4 * 8, 5 + 4 * 8, 6 + 3 * 8, 7 + 3 * 8, 6 + 4 * 8, 7 + 4 * 8,
                               1 + 1 * 8, 2 + 1 * 8, 1 + 2 * 8, 2 + 2 * 8, 1 + 4 * 8, 2 + 4 * 8, 1 + 5 * 8, 2 + 5 * 8 };   // common/common.h:217-231

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s dump.bin\n", argv[0]); return 2; }
    Dump d;
    if (!d.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const DumpRec *cmv = d.find("CMV0"), *qnt = d.find("QNT0"), *vlc = d.find("VLC0");
    if (!cmv || !qnt || !vlc || vlc->size != sizeof(CavlcSizes)) { fprintf(stderr, "dump lacks CMV0 / QNT0 / VLC0\n"); return 2; }
    CavlcSizes z; memcpy(&z, vlc->data, sizeof(z));
    int cmv_qp, cmv_lambda; memcpy(&cmv_qp, cmv->data, 4); memcpy(&cmv_lambda, cmv->data + 4, 4);
    int32_t q3[3]; memcpy(q3, qnt->data, 12);
    const int mb_w = d.cfg[2], mb_h = d.cfg[3], n_mb = mb_w * mb_h;
    DevFrameCtx fc; memset(&fc, 0, sizeof(fc));
    fc.width = 16 * mb_w; fc.height = 16 * mb_h; fc.mb_w = mb_w; fc.mb_h = mb_h;
    fc.me_method = d.cfg[4]; fc.me_range = d.cfg[5]; fc.subme = d.cfg[6]; fc.max_refs = d.cfg[7]; fc.chroma_me = d.cfg[8];
    fc.mv_range = d.cfg[9]; fc.b_cabac = d.cfg[11]; fc.analyse_inter = d.cfg[12]; fc.b_fast_pskip = d.cfg[14]; fc.b_dct_decimate = d.cfg[15];
    fc.tab.cost_mv = (const int16_t *)(cmv->data + 8) + 16384;
    fc.tab.cost_ref = (const uint16_t *)(cmv->data + 8 + 32769 * 2);
    fc.tab.qp = q3[0]; fc.tab.chroma_qp = q3[1]; fc.tab.lambda2_chroma = q3[2]; fc.tab.lambda = cmv_lambda;
    const uint8_t *qp_ = qnt->data + 12;
    fc.tab.quant4_mf[0] = (const uint16_t *)qp_; fc.tab.quant4_bias[0] = (const uint16_t *)(qp_ + 32);
    fc.tab.quant4_mf[1] = (const uint16_t *)(qp_ + 64); fc.tab.quant4_bias[1] = (const uint16_t *)(qp_ + 96);
    fc.tab.dequant4_mf[0] = (const int32_t *)(qp_ + 128); fc.tab.dequant4_mf[1] = (const int32_t *)(qp_ + 128 + 384);

    MbWork work;
    std::vector<uint8_t> rec_y, rec_u, rec_v; std::vector<uint16_t> rec_nnz(n_mb);
    long n = 0, bad_ssd = 0, bad_bits = 0, bad_nnz = 0, n_intra = 0, n_psy = 0, n_q1 = 0, bad_device = 0;
    for (size_t ri = 0; ri < d.recs.size(); ri++)
    {
        if (strcmp(d.recs[ri].tag, "SLCB")) continue;
        SlicePlanes sp; sp.parse(d.recs[ri]);
        if (!sp.hd.with_planes || sp.hd.type != 0 || sp.hd.qp != cmv_qp) continue;
        const DumpRec &sx = d.recs[ri + 1];
        if (strcmp(sx.tag, "SLCX")) { fprintf(stderr, "SLCX missing\n"); return 2; }
        int32_t hx[36]; memcpy(hx, sx.data, sizeof(hx));
        fc.stride_y = sp.hd.stride_y; fc.stride_c = sp.hd.stride_c;
        fc.fenc_y = sp.fenc[0]; fc.fenc_u = sp.fenc[1]; fc.fenc_v = sp.fenc[2];
        FrameParams fp; memset(&fp, 0, sizeof(fp));
        fp.pass = sp.hd.pass; fp.n_ref = hx[1]; fp.cur_poc = hx[0];
        for (int i = 0; i < fp.n_ref; i++)
        {
            fp.ref_slot[i] = i; fp.ref_poc[i] = hx[2 + i];
            DevRef &r = fc.ref[i];
            for (int k = 0; k < 4; k++) r.y[k] = (uint8_t *)sp.refs[i].y[k];
            r.u = (uint8_t *)sp.refs[i].u; r.v = (uint8_t *)sp.refs[i].v; r.valid = 1;
        }
        rec_y.assign((size_t)fc.stride_y * (sp.hd.lines_y + 64), 0); rec_u.assign((size_t)fc.stride_c * (sp.hd.lines_y / 2 + 32), 0); rec_v = rec_u;
        ReconPlanes rp;
        rp.y = rec_y.data() + (size_t)fc.stride_y * 32 + 32; rp.u = rec_u.data() + (size_t)fc.stride_c * 16 + 16; rp.v = rec_v.data() + (size_t)fc.stride_c * 16 + 16;
        rp.stride_y = fc.stride_y; rp.stride_c = fc.stride_c; rp.nnz = rec_nnz.data();
        for (size_t rj = ri + 1; rj < d.recs.size(); rj++)
        {
            const DumpRec &r = d.recs[rj];
            if (!strcmp(r.tag, "SLCE") || !strcmp(r.tag, "SLCB")) break;
            if (strcmp(r.tag, "RDMB")) continue;
            int32_t hd[20]; memcpy(hd, r.data, sizeof(hd));
            if (hd[3] != 4 && hd[3] != 5) { n_intra++; continue; }
            const uint8_t *p = r.data + sizeof(hd);
            CavlcMb m; memset(&m, 0, sizeof(m));
            m.type = hd[3]; m.partition = hd[4];
            for (int i = 0; i < 4; i++) m.sub[i] = hd[5 + i];
            m.n_ref = hd[9]; m.psub8x8 = hd[10]; m.cbp_luma = hd[11]; m.cbp_chroma = hd[12]; m.qp_delta = hd[13]; m.n_mvd = hd[17];
            const int want_ssd = hd[14], want_bits = hd[15], psy_rd = hd[18], lambda = hd[19];
            memcpy(m.ref, p, 4); p += 4;
            memcpy(m.mvd, p, sizeof(m.mvd)); p += sizeof(m.mvd);
            const uint8_t *nnz = p; p += 48;
            m.coef = (const int16_t (*)[16])p; p += 24 * 16 * 2;
            m.chroma_dc = (const int16_t (*)[4])p; p += 2 * 4 * 2;
            int16_t mv[16][2]; memcpy(mv, p, sizeof(mv)); p += sizeof(mv);
            int32_t skip_mc; memcpy(&skip_mc, p, 4);
            // quirk q1 (SURVEY / DESIGN.md §4): b_skip_mc left set, the reference does not motion-compensate this "candidate" but codes
            // it against whatever its last analysis left in fdec; in the product those macroblocks are the host's (early_skip == 2)
            if (skip_mc) { n_q1++; continue; }
            for (int y = 0; y < 4; y++) m.nnz_left[y] = nnz[x264_scan8_tab[0] - 1 + 8 * y];
            for (int x = 0; x < 4; x++) m.nnz_top[x] = nnz[x264_scan8_tab[0] - 8 + x];
            for (int pl = 0; pl < 2; pl++)
                for (int k = 0; k < 2; k++) { m.nnz_left_c[pl][k] = nnz[x264_scan8_tab[16 + 4 * pl] - 1 + 8 * k]; m.nnz_top_c[pl][k] = nnz[x264_scan8_tab[16 + 4 * pl] - 8 + k]; }
            for (int i = 0; i < 24; i++) m.coded[i] = nnz[x264_scan8_tab[i]];
            m.coded[24] = nnz[5 + 5 * 8]; m.coded[25] = nnz[6 + 5 * 8];

            // (1) the candidate through the device's reconstruction
            const int mb = hd[2];
            MbResult res; memset(&res, 0, sizeof(res));
            res.type = (int8_t)m.type; res.partition = (int8_t)m.partition;
            memcpy(res.ref, m.ref, 4);
            for (int i = 0; i < 16; i++) res.mv[i] = pack_mv(mv[i][0], mv[i][1]);
            MbCtx c(fc, fp, work);
            c.mb_x = mb % mb_w; c.mb_y = mb / mb_w; c.mb_xy = mb;
            for (int y = 0; y < 16; y++) memcpy(work.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * c.mb_y + y) * fc.stride_y + 16 * c.mb_x, 16);
            for (int y = 0; y < 8; y++)
            {
                memcpy(work.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                memcpy(work.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
            }
            recon_mb(c, res, rp);
            const int got_ssd = rd_distortion_mb(work.fenc_y, work.pred_y, work.fenc_u, work.pred_u, work.fenc_v, work.pred_v, psy_rd, lambda);
            n++; n_psy += psy_rd != 0;
            if (got_ssd != want_ssd)
            {
                if (bad_ssd < 5) fprintf(stderr, "frame %d mb %d type %d partition %d: distortion %d, reference %d\n", hd[0], mb, m.type, m.partition, got_ssd, want_ssd);
                bad_ssd++;
            }
            // (2) which luma blocks keep coefficients (raster bit x + 4 y in rec_nnz)
            int want_mask = 0;
            for (int i = 0; i < 16; i++)
                if (m.coded[i]) want_mask |= 1 << (((i & 1) | ((i >> 1) & 2)) + 4 * (((i >> 1) & 1) | ((i >> 2) & 2)));
            if (want_mask != rec_nnz[mb]) { if (bad_nnz < 5) fprintf(stderr, "frame %d mb %d: kept-coefficient mask %04x, reference %04x\n", hd[0], mb, rec_nnz[mb], want_mask); bad_nnz++; }
            // (3) size, from the reference's coefficient arrays
            const int got_bits = cavlc_mb_inter_bits(z, m);
            if (got_bits != want_bits) { if (bad_bits < 5) fprintf(stderr, "frame %d mb %d: %d bits, reference %d\n", hd[0], mb, got_bits, want_bits); bad_bits++; }
            // (4) the same candidate once more, device-only: motion compensation, the device's OWN levels / kept flags / cbp, then the
            //     residual path the product runs, distortion and size from nothing the reference computed but the vector differences
            //     and the neighbours' coefficient counts
            {
                rd_mc_inter(c, res);
                static RdLevels lv;
                rd_levels_mb(c, lv);
                encode_mb_residual(c);
                const int ssd2 = rd_distortion_mb(work.fenc_y, work.pred_y, work.fenc_u, work.pred_u, work.fenc_v, work.pred_v, psy_rd, lambda);
                CavlcMb m2 = m;
                m2.coef = lv.coef; m2.chroma_dc = lv.chroma_dc; m2.cbp_luma = lv.cbp_luma; m2.cbp_chroma = lv.cbp_chroma;
                memcpy(m2.coded, lv.coded, 26);
                const int bits2 = cavlc_mb_inter_bits(z, m2);
                const int cost = ssd2 + ((bits2 * hd[16] + 128) >> 8), want_cost = want_ssd + ((want_bits * hd[16] + 128) >> 8);
                if (cost != want_cost || lv.cbp_luma != hd[11] || lv.cbp_chroma != hd[12])
                {
                    if (bad_device < 5)
                        fprintf(stderr, "frame %d mb %d type %d: device-only RD cost %d (ssd %d, %d bits, cbp %x/%d), reference %d (ssd %d, %d bits, cbp %x/%d)\n",
                                hd[0], mb, m.type, cost, ssd2, bits2, lv.cbp_luma, lv.cbp_chroma, want_cost, want_ssd, want_bits, hd[11], hd[12]);
                    if (bad_device < 3 && getenv("PCAMV_EMU_RD_DEBUG"))
                    {
                        for (int pl = 0; pl < 2; pl++)
                            fprintf(stderr, "   chroma dc %d: device %d %d %d %d (coded %d), reference %d %d %d %d (coded %d)\n", pl, lv.chroma_dc[pl][0], lv.chroma_dc[pl][1], lv.chroma_dc[pl][2], lv.chroma_dc[pl][3], lv.coded[24 + pl],
                                    m.chroma_dc[pl][0], m.chroma_dc[pl][1], m.chroma_dc[pl][2], m.chroma_dc[pl][3], m.coded[24 + pl]);
                        for (int i = 0; i < 24; i++)
                            if (lv.coded[i] || m.coded[i])
                            {
                                fprintf(stderr, "   block %d coded %d/%d device:", i, lv.coded[i], m.coded[i]);
                                for (int k = 0; k < 16; k++) fprintf(stderr, " %d", lv.coef[i][k]);
                                fprintf(stderr, "  reference:");
                                for (int k = 0; k < 16; k++) fprintf(stderr, " %d", m.coef[i][k]);
                                fprintf(stderr, "\n");
                            }
                    }
                    bad_device++;
                }
            }
        }
    }
    printf("candidates=%ld bad_distortion=%ld bad_bits=%ld bad_kept_mask=%ld bad_device_only_cost=%ld intra_skipped=%ld with_psy=%ld q1_skipped=%ld\n", n, bad_ssd, bad_bits, bad_nnz, bad_device, n_intra, n_psy, n_q1);
    return (bad_ssd || bad_bits || bad_nnz || bad_device) ? 1 : 0;
}
