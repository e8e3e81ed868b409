// CPU-side logic check of the frame-level device code (-DPCAMV_EMU: lane team of one).
// Runs the P-slice analysis of every dumped frame/pass macroblock by macroblock in raster order and compares
//   (1) the per-MB log of search / refine results with the reference's recorded calls, in call order,
//   (2) the per-MB decision (type, partition, final cache MVs / refs) with the reference's 'MBAN' records,
//   (3) pass 1: the candidate-MV cost table (delta, cost) with info.cache[] from the 'EMBD' record,
// for dumps produced by oracle/_ref/x264_dump.  usage: emu_frame_check DUMP.bin [max_frames]
#define PCAMV_EMU 1
#include "../../video-steganography-pcamv_b200/csrc/pcamv_device.h"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_cost.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_recon.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_glue.h"
#include "dump_reader.h"
#include <map>

using namespace pcamv;

struct MbanRec { int32_t frame, pass, mb_xy, type, partition, sub[4], b_skip_mc, qp; int16_t mv[16][2]; int8_t ref[4]; int16_t pskip[2]; };
#pragma pack(push, 1)
struct EmbdMb { int32_t type, qp, partition; uint8_t used, sub[4], pad[3]; int16_t mv_stego[16][2]; int32_t cost[16]; int8_t ref[16]; int16_t mv[16][2]; int16_t pskip[2]; };
#pragma pack(pop)
static_assert(sizeof(EmbdMb) == 232, "EMBD layout");
static_assert(sizeof(MbanRec) == 116, "MBAN layout");

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s dump.bin [max_frames]\n", argv[0]); return 2; }
    const int max_frames = argc > 2 ? atoi(argv[2]) : 1000;
    Dump d;
    if (!d.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const DumpRec *cmv = d.find("CMV0"), *qnt = d.find("QNT0");
    if (!cmv || !qnt) { fprintf(stderr, "dump lacks CMV0/QNT0\n"); return 2; }
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

    std::vector<int8_t> type(n_mb), ref8(4 * n_mb); std::vector<uint32_t> mv4(16 * n_mb), mvr((size_t)PCAMV_MAX_REFS * n_mb);
    std::vector<LogEntry> log((size_t)n_mb * PCAMV_LOG_MAX); std::vector<MbResult> results(n_mb);
    std::vector<ForcedOut> forced(n_mb);
    std::vector<PartInfo> subparts((size_t)16 * n_mb);
    std::map<int, std::vector<uint32_t>> last_mv_pass1;      // frame -> final cache MVs of the last MB of pass 1
    MbWork work;
    // checker-side integral planes (8x8 box sums of the padded integer plane, what k_box_sum8 builds on the GPU) and the
    // --me tesa candidate list of the single emulated team
    std::vector<std::vector<uint16_t>> integral(PCAMV_MAX_REFS), integral4(PCAMV_MAX_REFS);
    std::vector<unsigned long long> mvsads((size_t)(2 * fc.me_range + 4) * (2 * fc.me_range + 1));

    // PCAMV_EMU_ASYNC=1: run the resumable (split wavefront) form of the analysis; not for configurations with sub-8x8 partitions
    // PCAMV_EMU_CONFORMANT=1: the dump comes from oracle/_ref/x264_dump_conformant (tools/reftree.py::conformance_switch); the device
    // logic runs with pcamv_set_conformant's switch on
    fc.conformant = getenv("PCAMV_EMU_CONFORMANT") && atoi(getenv("PCAMV_EMU_CONFORMANT"));
    const bool async_mode = getenv("PCAMV_EMU_ASYNC") && atoi(getenv("PCAMV_EMU_ASYNC")) && !(fc.analyse_inter & 0x20);
    long n_yield = 0;
    // (4) reconstruction + deblocking of the frame's final pass (pcamv_recon.cuh) against the reference planes the NEXT frame was
    //     dumped with (its reference 0 = this frame after x264_fdec_filter_row)
    std::vector<uint8_t> rec_y, rec_u, rec_v; std::vector<uint16_t> rec_nnz(n_mb);
    int rec_frame = -1, rec_q1 = 0; long n_recon = 0, bad_recon = 0, n_recon_q1 = 0;
    std::vector<int> rec_type(n_mb); std::vector<uint32_t> rec_mv0(n_mb); std::vector<int8_t> rec_ref(4 * n_mb);
    long n_call = 0, bad_call = 0, n_mbs = 0, bad_mb = 0, n_ih = 0, bad_ih = 0, n_passes = 0;
    int frames_done = 0;
    for (size_t ri = 0; ri < d.recs.size(); ri++)
    {
        if (strcmp(d.recs[ri].tag, "SLCB")) continue;
        SlicePlanes sp; sp.parse(d.recs[ri]);
        if (!sp.hd.with_planes || sp.hd.type != 0) continue;
        if (sp.hd.qp != cmv_qp) continue;
        const DumpRec &sx = d.recs[ri + 1];
        if (strcmp(sx.tag, "SLCX")) { fprintf(stderr, "SLCX missing\n"); return 2; }
        int32_t hx[36]; memcpy(hx, sx.data, sizeof(hx));
        if (frames_done >= max_frames && sp.hd.pass != 2) break;

        fc.stride_y = sp.hd.stride_y; fc.stride_c = sp.hd.stride_c;
        fc.fenc_y = sp.fenc[0]; fc.fenc_u = sp.fenc[1]; fc.fenc_v = sp.fenc[2];
        if (rec_frame >= 0 && sp.hd.frame == rec_frame + 1 && sp.hd.pass != 2 && rec_q1)
        {
            // macroblocks the host reconstructed from its intra analysis' leftovers (quirk q1): the device gets those as patches
            // from the host (pcamv_reconstruct_ref), which this dump-driven check does not have — the GPU host tests cover it
            n_recon_q1++;
            rec_frame = -1;
        }
        if (rec_frame >= 0 && sp.hd.frame == rec_frame + 1 && sp.hd.pass != 2)
        {
            const int W = 16 * mb_w, H = 16 * mb_h;
            const uint8_t *py = rec_y.data() + (size_t)fc.stride_y * 32 + 32, *pu = rec_u.data() + (size_t)fc.stride_c * 16 + 16, *pv = rec_v.data() + (size_t)fc.stride_c * 16 + 16;
            long bad = 0; int fx = -1, fy = -1, fpl = -1;
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    if (py[(size_t)y * fc.stride_y + x] != sp.refs[0].y[0][(size_t)y * fc.stride_y + x]) { if (!bad) { fx = x; fy = y; fpl = 0; } bad++; }
            for (int y = 0; y < H / 2; y++)
                for (int x = 0; x < W / 2; x++)
                {
                    if (pu[(size_t)y * fc.stride_c + x] != sp.refs[0].u[(size_t)y * fc.stride_c + x]) { if (!bad) { fx = x; fy = y; fpl = 1; } bad++; }
                    if (pv[(size_t)y * fc.stride_c + x] != sp.refs[0].v[(size_t)y * fc.stride_c + x]) { if (!bad) { fx = x; fy = y; fpl = 2; } bad++; }
                }
            n_recon++;
            if (bad && getenv("PCAMV_EMU_RECON_DEBUG"))
            {
                int shown = 0;
                for (int mb = 0; mb < n_mb && shown < 6; mb++)
                {
                    const int mx = mb % mb_w, my = mb / mb_w;
                    int cnt = 0; char map[17][17]; memset(map, 0, sizeof(map));
                    for (int y = 0; y < 16; y++)
                        for (int x = 0; x < 16; x++)
                        {
                            const size_t o = (size_t)(16 * my + y) * fc.stride_y + 16 * mx + x;
                            const bool d = py[o] != sp.refs[0].y[0][o];
                            map[y][x] = d ? 'X' : '.'; cnt += d;
                        }
                    if (!cnt) continue;
                    shown++;
                    fprintf(stderr, "  MB %d,%d type %d nnz %04x (left %04x top %04x) %d luma pixels differ; mv0 (%d,%d) ref %d %d %d %d\n", mx, my, rec_type[mb], rec_nnz[mb],
                            mx ? rec_nnz[mb - 1] : 0, my ? rec_nnz[mb - mb_w] : 0, cnt, mv_x(rec_mv0[mb]), mv_y(rec_mv0[mb]), rec_ref[4 * mb], rec_ref[4 * mb + 1], rec_ref[4 * mb + 2], rec_ref[4 * mb + 3]);
                    for (int y = 0; y < 16; y++) fprintf(stderr, "    %s\n", map[y]);
                }
            }
            if (bad)
            {
                bad_recon++;
                fprintf(stderr, "frame %d: reconstruction + deblocking differs from the reference's plane in %ld pixels (first: plane %d x %d y %d = MB %d,%d)\n",
                        rec_frame, bad, fpl, fx, fy, fpl ? fx / 8 : fx / 16, fpl ? fy / 8 : fy / 16);
            }
            rec_frame = -1;
        }
        FrameParams fp; memset(&fp, 0, sizeof(fp));
        fp.pass = sp.hd.pass; fp.n_ref = hx[1]; fp.cur_poc = hx[0];
        for (int i = 0; i < fp.n_ref; i++)
        {
            fp.ref_slot[i] = i; fp.ref_poc[i] = hx[2 + i];
            DevRef &r = fc.ref[i];
            for (int k = 0; k < 4; k++) r.y[k] = (uint8_t *)sp.refs[i].y[k];
            r.u = (uint8_t *)sp.refs[i].u; r.v = (uint8_t *)sp.refs[i].v; r.valid = 1;
            r.integral = nullptr; r.integral4 = nullptr;
            if (fc.me_method >= ME_ESA && (fc.analyse_inter & 0x20))
            {
                const int st = fc.stride_y, rows = sp.hd.lines_y + 64;
                const uint8_t *base = sp.refs[i].y[0] - (size_t)st * 32 - 32;
                std::vector<uint16_t> &sum = integral4[i];
                sum.assign((size_t)st * rows, 0);
                for (int y = 0; y + 4 <= rows; y++)
                    for (int x = 0; x + 4 <= st; x++)
                    {
                        int a = 0;
                        for (int yy = 0; yy < 4; yy++)
                            for (int xx = 0; xx < 4; xx++) a += base[(size_t)(y + yy) * st + x + xx];
                        sum[(size_t)y * st + x] = (uint16_t)a;
                    }
                r.integral4 = sum.data() + (size_t)st * 32 + 32;
            }
            if (fc.me_method >= ME_ESA)
            {
                const int st = fc.stride_y, rows = sp.hd.lines_y + 64;
                const uint8_t *base = sp.refs[i].y[0] - (size_t)st * 32 - 32;
                std::vector<uint16_t> &sum = integral[i];
                sum.assign((size_t)st * rows, 0);
                for (int y = 0; y + 8 <= rows; y++)
                    for (int x = 0; x + 8 <= st; x++)
                    {
                        int a = 0;
                        for (int yy = 0; yy < 8; yy++)
                            for (int xx = 0; xx < 8; xx++) a += base[(size_t)(y + yy) * st + x + xx];
                        sum[(size_t)y * st + x] = (uint16_t)a;
                    }
                r.integral = sum.data() + (size_t)st * 32 + 32;
            }
        }
        fp.col_n_ref = hx[18];
        for (int i = 0; i < 16; i++) fp.col_inv_ref_poc[i] = hx[19 + i];
        fp.col_ref8 = (const int8_t *)(sx.data + sizeof(hx));
        fp.col_mv4 = (const uint32_t *)(sx.data + sizeof(hx) + 4 * n_mb);
        fp.cur.type = type.data(); fp.cur.ref8 = ref8.data(); fp.cur.mv4 = mv4.data(); fp.cur.mvr = mvr.data();
        fp.log = log.data(); fp.log_stride = PCAMV_LOG_MAX; fp.results = results.data();
        fp.subparts = (fc.analyse_inter & 0x20) ? subparts.data() : nullptr;
        fp.mvsads = fc.me_method == ME_TESA ? mvsads.data() : nullptr; fp.mvsads_cap = 0;       // one team: every row shares the list

        // reference records of this frame/pass
        std::vector<std::vector<CallRec>> calls(n_mb);
        std::vector<MbanRec> mban(n_mb); std::vector<char> have_mban(n_mb, 0);
        const uint8_t *embd = nullptr; int embd_len = 0;
        for (size_t rj = ri + 1; rj < d.recs.size(); rj++)
        {
            const DumpRec &r = d.recs[rj];
            if (!strcmp(r.tag, "SLCE")) break;
            if (!strcmp(r.tag, "MESR") || !strcmp(r.tag, "MERQ"))
            {
                CallRec c; memcpy(&c, r.data, sizeof(c));
                c.me_method = !strcmp(r.tag, "MERQ");       // reuse the field as "is refine"
                calls[c.mb_xy].push_back(c);
            }
            else if (!strcmp(r.tag, "MBAN")) { MbanRec m; memcpy(&m, r.data, sizeof(m)); mban[m.mb_xy] = m; have_mban[m.mb_xy] = 1; }
            else if (!strcmp(r.tag, "EMBD")) { embd = r.data; memcpy(&embd_len, r.data + 8, 4); }
        }
        if (fp.pass == 2)
        {
            // forced decisions from the EMBD record of pass 1 of the same frame (search backwards)
            const uint8_t *e1 = nullptr;
            for (size_t rj = ri; rj-- > 0;)
                if (!strcmp(d.recs[rj].tag, "EMBD")) { int32_t f; memcpy(&f, d.recs[rj].data, 4); if (f == sp.hd.frame) { e1 = d.recs[rj].data; } break; }
            if (!e1) { fprintf(stderr, "pass 2 without EMBD\n"); return 2; }
            int32_t eh[5]; memcpy(eh, e1, 20);
            const EmbdMb *mbs = (const EmbdMb *)(e1 + 20);
            const uint8_t *p = e1 + 20 + 232 * (size_t)eh[1];
            const int len = eh[2], an = eh[3];
            const int8_t *filp = (const int8_t *)(p + len + 4 * (size_t)len + (an > 0 ? an : 0) + len);
            std::vector<Pass1Mb> in(n_mb);
            for (int i = 0; i < n_mb; i++)
            {
                in[i].type = mbs[i].type; in[i].partition = mbs[i].partition; in[i].used = mbs[i].used;
                memcpy(in[i].sub, mbs[i].sub, 4); memcpy(in[i].ref, mbs[i].ref, 16);
                memcpy(in[i].mv, mbs[i].mv, 64); memcpy(in[i].mv_stego, mbs[i].mv_stego, 64);
            }
            const int used = build_forced(n_mb, in.data(), filp, forced.data());
            if (used != len) { fprintf(stderr, "glue consumed %d flips, reference cover length %d\n", used, len); return 1; }
            fp.forced = (const ForcedMb *)forced.data();
            if (last_mv_pass1.count(sp.hd.frame))
                memcpy(fp.stale_mv, last_mv_pass1[sp.hd.frame].data(), 64);
        }

        for (int mb = 0; mb < n_mb; mb++)
        {
            MbCtx c(fc, fp, work);
            c.mb_x = mb % mb_w; c.mb_y = mb / mb_w; c.mb_xy = mb;
            for (int y = 0; y < 16; y++) memcpy(work.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * c.mb_y + y) * fc.stride_y + 16 * c.mb_x, 16);
            for (int y = 0; y < 8; y++)
            {
                memcpy(work.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                memcpy(work.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
            }
            if (!async_mode)
                analyse_p_mb<3>(c, mb ? results[mb - 1].mv : fp.stale_mv);
            else
            {
                // the split wavefront's protocol on one lane: every search leaves as a request; the macroblock's persistent
                // bytes (everything before MbWork::fenc_y, and the scalars of MbCtx) are parked, the rest of both structures is
                // trashed, a "search team" serves the request from a scratch MbWork of its own, and the analysis is resumed
                static MbWork park, team;
                static unsigned char park_ctx[sizeof(MbCtx)];
                work.pt.stage = 0;
                int guard = 0;
                while (analyse_p_mb_rs<1, 1>(c, mb ? results[mb - 1].mv : fp.stale_mv) == PT_YIELD)
                {
                    n_yield++;
                    if (++guard > 200) { fprintf(stderr, "mb %d never finishes\n", mb); return 1; }
                    const size_t keep = offsetof(MbWork, fenc_y), ctx0 = offsetof(MbCtx, mb_x);
                    memcpy(&park, &work, keep);
                    memcpy(park_ctx, (unsigned char *)&c + ctx0, sizeof(MbCtx) - ctx0);
                    const SearchReq rq = work.rq;
                    memset(&work, 0xA5, sizeof(work));
                    memset((unsigned char *)&c + ctx0, 0x5A, sizeof(MbCtx) - ctx0);
                    memset(&team, 0x3C, sizeof(team));
                    for (int y = 0; y < 16; y++) memcpy(team.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * rq.mb_y + y) * fc.stride_y + 16 * rq.mb_x, 16);
                    for (int y = 0; y < 8; y++)
                    {
                        memcpy(team.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * rq.mb_y + y) * fc.stride_c + 8 * rq.mb_x, 8);
                        memcpy(team.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * rq.mb_y + y) * fc.stride_c + 8 * rq.mb_x, 8);
                    }
                    SearchRes rs;
                    serve_request<1>(fc, fp, team, rq, rs);
                    memcpy(&work, &park, keep);
                    memcpy((unsigned char *)&c + ctx0, park_ctx, sizeof(MbCtx) - ctx0);
                    work.rs = rs;
                    // the control side restages the source pixels when it resumes (the P_SKIP probe reads them)
                    for (int y = 0; y < 16; y++) memcpy(work.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * c.mb_y + y) * fc.stride_y + 16 * c.mb_x, 16);
                    for (int y = 0; y < 8; y++)
                    {
                        memcpy(work.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                        memcpy(work.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                    }
                }
            }
            MbResult &res = results[mb];
            // (1) call log
            const std::vector<CallRec> &rc = calls[mb];
            const LogEntry *le = &log[(size_t)mb * PCAMV_LOG_MAX];
            bool ok = (int)rc.size() == res.n_log;
            for (size_t i = 0; ok && i < rc.size(); i++)
            {
                const bool thr_exit = !rc[i].me_method && rc[i].has_thresh;   // cost_mv may be stale after the multi-ref early-out
                ok = le[i].kind == (rc[i].me_method ? LOG_REFINE : LOG_SEARCH) && le[i].i_pixel == rc[i].i_pixel &&
                     le[i].mv[0] == rc[i].mv[0] && le[i].mv[1] == rc[i].mv[1] && le[i].cost == rc[i].cost &&
                     (thr_exit || le[i].cost_mv == rc[i].cost_mv);
            }
            n_call += rc.size();
            if (!ok)
            {
                if (bad_call < 8)
                {
                    fprintf(stderr, "frame %d pass %d mb %d: log differs (ours %d entries, reference %zu)\n", sp.hd.frame, fp.pass, mb, res.n_log, rc.size());
                    for (size_t i = 0; i < rc.size() || (int)i < res.n_log; i++)
                    {
                        if ((int)i < res.n_log) fprintf(stderr, "   ours k%d pix%d ref%d mv(%d,%d) cost %d cmv %d", le[i].kind, le[i].i_pixel, le[i].i_ref, le[i].mv[0], le[i].mv[1], le[i].cost, le[i].cost_mv);
                        if (i < rc.size()) fprintf(stderr, "   | ref k%d pix%d ref%d mvp(%d,%d) mv(%d,%d) cost %d cmv %d", rc[i].me_method, rc[i].i_pixel, rc[i].i_ref, rc[i].mvp[0], rc[i].mvp[1], rc[i].mv[0], rc[i].mv[1], rc[i].cost, rc[i].cost_mv);
                        fprintf(stderr, "\n");
                    }
                }
                bad_call++;
            }
            // (2) decision
            if (have_mban[mb])
            {
                const MbanRec &m = mban[mb];
                bool okm = m.type == res.type && (res.type == MB_P_SKIP || m.partition == res.partition);
                for (int i = 0; okm && i < 16; i++) okm = pack_mv(m.mv[i][0], m.mv[i][1]) == res.mv[i];
                for (int i = 0; okm && i < 4; i++) okm = m.ref[i] == res.ref[i];
                okm = okm && m.pskip[0] == res.pskip_mv[0] && m.pskip[1] == res.pskip_mv[1];
                n_mbs++;
                if (!okm)
                {
                    if (bad_mb < 8)
                        fprintf(stderr, "frame %d pass %d mb %d: decision differs: ours type %d part %d mv0 (%d,%d) pskip (%d,%d) | ref type %d part %d mv0 (%d,%d) pskip (%d,%d)\n",
                                sp.hd.frame, fp.pass, mb, res.type, res.partition, mv_x(res.mv[0]), mv_y(res.mv[0]), res.pskip_mv[0], res.pskip_mv[1],
                                m.type, m.partition, m.mv[0][0], m.mv[0][1], m.pskip[0], m.pskip[1]);
                    bad_mb++;
                }
            }
        }
        if (fp.pass != 1 && !(getenv("PCAMV_EMU_NO_RECON")))
        {
            rec_y.assign((size_t)fc.stride_y * (sp.hd.lines_y + 64), 0); rec_u.assign((size_t)fc.stride_c * (sp.hd.lines_y / 2 + 32), 0); rec_v = rec_u;
            ReconPlanes rp;
            rp.y = rec_y.data() + (size_t)fc.stride_y * 32 + 32; rp.u = rec_u.data() + (size_t)fc.stride_c * 16 + 16; rp.v = rec_v.data() + (size_t)fc.stride_c * 16 + 16;
            rp.stride_y = fc.stride_y; rp.stride_c = fc.stride_c; rp.nnz = rec_nnz.data();
            for (int mb = 0; mb < n_mb; mb++)
            {
                MbCtx c(fc, fp, work);
                c.mb_x = mb % mb_w; c.mb_y = mb / mb_w; c.mb_xy = mb;
                for (int y = 0; y < 16; y++) memcpy(work.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * c.mb_y + y) * fc.stride_y + 16 * c.mb_x, 16);
                for (int y = 0; y < 8; y++)
                {
                    memcpy(work.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                    memcpy(work.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                }
                recon_mb(c, results[mb], rp);
            }
            DbFrame df; df.type = type.data(); df.ref8 = ref8.data(); df.mv4 = mv4.data(); df.nnz = rec_nnz.data(); df.mb_w = mb_w;
            DeblockParams dp; memset(&dp, 0, sizeof(dp));
            dp.qp = fc.tab.qp; dp.qp_chroma = fc.tab.chroma_qp; dp.no_sub8x8_all = !(fc.analyse_inter & 0x20);
            alignas(16) uint8_t db_stage[(DB_STAGE_BYTES + 15) & ~15];
            std::vector<uint32_t> db_bs(8 * (size_t)n_mb);             // strengths of the whole frame first, as k_deblock_bs does
            for (int t = 0; t < 32 * n_mb; t++)
                ((uint8_t *)db_bs.data())[t] = (uint8_t)db_strength_piece(df, dp, (t >> 5) % mb_w, (t >> 5) / mb_w, t & 31);
            for (int mb = 0; mb < n_mb; mb++)
                deblock_mb(rp, dp, mb % mb_w, mb / mb_w, (const uint8_t *)(db_bs.data() + 8 * (size_t)mb), db_stage);
            rec_frame = sp.hd.frame; rec_q1 = 0;
            for (int mb = 0; mb < n_mb; mb++) rec_q1 += results[mb].early_skip == 2;
            for (int mb = 0; mb < n_mb; mb++) { rec_type[mb] = results[mb].type; rec_mv0[mb] = results[mb].mv[0]; memcpy(&rec_ref[4 * mb], results[mb].ref, 4); }
        }
        // (3) cost table (pass 1)
        if (fp.pass == 1 && embd)
        {
            int32_t eh[5]; memcpy(eh, embd, 20);
            const EmbdMb *mbs = (const EmbdMb *)(embd + 20);
            for (int mb = 0; mb < n_mb; mb++)
            {
                MbResult &res = results[mb];
                if (res.type == MB_P_SKIP) continue;
                MbCtx c(fc, fp, work);
                c.mb_x = mb % mb_w; c.mb_y = mb / mb_w; c.mb_xy = mb;
                for (int y = 0; y < 16; y++) memcpy(work.fenc_y + 16 * y, fc.fenc_y + (size_t)(16 * c.mb_y + y) * fc.stride_y + 16 * c.mb_x, 16);
                for (int y = 0; y < 8; y++)
                {
                    memcpy(work.fenc_u + 8 * y, fc.fenc_u + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                    memcpy(work.fenc_v + 8 * y, fc.fenc_v + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x, 8);
                }
                const int n0 = res.n_log;
                cost_table_mb(c, res);
                const LogEntry *le = &log[(size_t)mb * PCAMV_LOG_MAX + n0];
                for (int k = 0; k < res.n_part; k++)
                {
                    const PartInfo &pk = (res.type == MB_P_8x8 && fp.subparts) ? fp.subparts[(size_t)16 * mb + k] : res.part[k];
                    int slot = res.partition == PART_16x16 ? 0 : res.partition == PART_16x8 ? 8 * k : res.partition == PART_8x16 ? 4 * k : 4 * k;
                    if (res.type == MB_P_8x8)
                    {
                        // info.cache slot of a (sub-)block of P_8x8 (analyse.c:3559-3601): 4 * i8 + { 8x8: 0, 8x4: 2j, 4x8: j, 4x4: j }
                        const int i8 = (pk.xoff >> 3) + 2 * (pk.yoff >> 3), sx = (pk.xoff & 7) >> 2, sy = (pk.yoff & 7) >> 2;
                        slot = 4 * i8 + (pk.i_pixel == PIX_8x8 ? 0 : pk.i_pixel == PIX_8x4 ? 2 * sy : pk.i_pixel == PIX_4x8 ? sx : sx + 2 * sy);
                    }
                    const int want_x = mbs[mb].mv_stego[slot][0] - pk.mv[0], want_y = mbs[mb].mv_stego[slot][1] - pk.mv[1];
                    n_ih++;
                    if (le[k].mv[0] != want_x || le[k].mv[1] != want_y || le[k].cost != mbs[mb].cost[slot])
                    {
                        if (bad_ih < 8)
                            fprintf(stderr, "frame %d mb %d part %d: cost table differs: ours d(%d,%d) cost %d | ref d(%d,%d) cost %d\n",
                                    sp.hd.frame, mb, k, le[k].mv[0], le[k].mv[1], le[k].cost, want_x, want_y, mbs[mb].cost[slot]);
                        bad_ih++;
                    }
                }
            }
        }
        if (fp.pass == 1)
            last_mv_pass1[sp.hd.frame] = std::vector<uint32_t>(results[n_mb - 1].mv, results[n_mb - 1].mv + 16);
        n_passes++;
        if (fp.pass != 1) frames_done++;
    }
    if (async_mode) fprintf(stderr, "async: %ld searches handed out\n", n_yield);
    printf("passes=%ld calls=%ld bad_mb_logs=%ld mbs=%ld bad_decisions=%ld ih=%ld bad_ih=%ld recon=%ld bad_recon=%ld recon_q1=%ld\n", n_passes, n_call, bad_call, n_mbs, bad_mb, n_ih, bad_ih,
           n_recon, bad_recon, n_recon_q1);
    return (bad_call || bad_mb || bad_ih || !n_passes) ? 1 : 0;
}
