// CPU-side logic check of the device search code (compiled with -DPCAMV_EMU: a lane team of one).
// Replays every x264_me_search_ref / x264_me_refine_qpel call recorded by the instrumented reference
// (oracle/_ref/x264_dump) and compares mv / cost / cost_mv / halfpel threshold.
// usage: emu_search_check DUMP.bin   -> prints "calls=N mismatches=M", exit code 1 on any mismatch
#define PCAMV_EMU 1
#include "../../video-steganography-pcamv_b200/csrc/pcamv_me.cuh"
#include "dump_reader.h"
#include <map>
#include <vector>

using namespace pcamv;

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s dump.bin\n", argv[0]); return 2; }
    Dump d;
    if (!d.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const DumpRec *cmv = d.find("CMV0");
    if (!cmv) { fprintf(stderr, "no CMV0 record\n"); return 2; }
    const int16_t *cost_mv = (const int16_t *)(cmv->data + 8) + 16384;
    int cmv_qp; memcpy(&cmv_qp, cmv->data, 4);

    SlicePlanes sp; bool have = false;
    long calls = 0, bad = 0, skipped = 0;
    for (const DumpRec &r : d.recs)
    {
        if (!strcmp(r.tag, "SLCB")) { sp.parse(r); have = sp.hd.with_planes; continue; }
        const bool is_search = !strcmp(r.tag, "MESR"), is_refine = !strcmp(r.tag, "MERQ");
        if (!is_search && !is_refine) continue;
        CallRec c; memcpy(&c, r.data, sizeof(c));
        if (!have || c.frame != sp.hd.frame || c.qp != cmv_qp) { skipped++; continue; }
        MeEnv env; memset(&env, 0, sizeof(env));
        env.cost_mv = cost_mv;
        env.me_method = c.me_method; env.me_range = c.me_range; env.subme = c.subme; env.chroma_me = c.b_chroma_me;
        env.mbcmp_satd = c.subme > 1;
        for (int k = 0; k < 2; k++)
        {
            env.mv_min_fpel[k] = c.mv_min_fpel[k]; env.mv_max_fpel[k] = c.mv_max_fpel[k];
            env.mv_min_spel[k] = c.mv_min_spel[k]; env.mv_max_spel[k] = c.mv_max_spel[k];
        }
        // stage the macroblock's source pixels: Y 16x16 stride 16, U/V 8x8 stride 8
        alignas(16) uint8_t fy[256], fu[64], fv[64];
        for (int y = 0; y < 16; y++)
            memcpy(fy + 16 * y, sp.fenc[0] + (size_t)(16 * c.mb_y + y) * sp.hd.stride_y + 16 * c.mb_x, 16);
        for (int y = 0; y < 8; y++)
        {
            memcpy(fu + 8 * y, sp.fenc[1] + (size_t)(8 * c.mb_y + y) * sp.hd.stride_c + 8 * c.mb_x, 8);
            memcpy(fv + 8 * y, sp.fenc[2] + (size_t)(8 * c.mb_y + y) * sp.hd.stride_c + 8 * c.mb_x, 8);
        }
        const SlicePlanes::Ref &rf = sp.refs[c.i_ref];
        MeBlock b; memset(&b, 0, sizeof(b));
        b.i_pixel = c.i_pixel; b.bw = pix_w(c.i_pixel); b.bh = pix_h(c.i_pixel);
        b.fenc = fy + c.yoff * 16 + c.xoff;
        b.fenc_u = fu + (c.yoff >> 1) * 8 + (c.xoff >> 1);
        b.fenc_v = fv + (c.yoff >> 1) * 8 + (c.xoff >> 1);
        b.stride = sp.hd.stride_y; b.stride_c = sp.hd.stride_c;
        const size_t off = (size_t)(16 * c.mb_y + c.yoff) * b.stride + 16 * c.mb_x + c.xoff;
        for (int k = 0; k < 4; k++) b.ref[k] = rf.y[k] + off;
        const size_t offc = (size_t)(8 * c.mb_y + (c.yoff >> 1)) * b.stride_c + 8 * c.mb_x + (c.xoff >> 1);
        b.ref_u = rf.u + offc; b.ref_v = rf.v + offc;
        if (c.me_method >= ME_ESA)
        {
            // checker-side integral planes of this reference (8x8 and 4x4 box sums of the padded plane), built on first use
            static std::map<const uint8_t *, std::pair<std::vector<uint16_t>, std::vector<uint16_t>>> sums;
            auto it = sums.find(rf.y[0]);
            if (it == sums.end())
            {
                const int st = b.stride, rows = sp.hd.lines_y + 64;
                const uint8_t *base = rf.y[0] - (size_t)st * 32 - 32;
                auto &pr = sums[rf.y[0]];
                for (int n = 8; n >= 4; n -= 4)
                {
                    std::vector<uint16_t> &sum = n == 8 ? pr.first : pr.second;
                    sum.assign((size_t)st * rows, 0);
                    for (int y = 0; y + n <= rows; y++)
                        for (int x = 0; x + n <= st; x++)
                        {
                            int a = 0;
                            for (int yy = 0; yy < n; yy++)
                                for (int xx = 0; xx < n; xx++) a += base[(size_t)(y + yy) * st + x + xx];
                            sum[(size_t)y * st + x] = (uint16_t)a;
                        }
                }
                it = sums.find(rf.y[0]);
            }
            b.integral = it->second.first.data() + (size_t)b.stride * 32 + 32 + off;
            b.integral4 = it->second.second.data() + (size_t)b.stride * 32 + 32 + off;
            static std::vector<unsigned long long> mvsads;
            mvsads.resize((size_t)(2 * c.me_range + 4) * (2 * c.me_range + 1));
            env.mvsads = mvsads.data();
        }
        block_set_mvp(b, env, c.mvp[0], c.mvp[1]);

        MeResult m; int thresh = c.thresh_in;
        if (is_search)
        {
            int mvc[10][2];
            for (int i = 0; i < c.i_mvc && i < 10; i++) { mvc[i][0] = c.mvc[i][0]; mvc[i][1] = c.mvc[i][1]; }
            m.mv[0] = m.mv[1] = 0; m.cost = 0; m.cost_mv = 0;
            me_search_ref<1>(env, b, mvc, c.i_mvc, c.has_thresh ? &thresh : nullptr, m);
        }
        else
        {
            m.mv[0] = c.mv_in[0]; m.mv[1] = c.mv_in[1]; m.cost = c.cost_in; m.cost_mv = c.cost_mv_in;
            me_refine_qpel(env, b, m, c.i_ref_cost);
        }
        calls++;
        // cost_mv is left stale by the multi-ref early-out of refine_subpel; compare it only otherwise
        bool ok = m.mv[0] == c.mv[0] && m.mv[1] == c.mv[1] && m.cost == c.cost;
        if (is_search && c.has_thresh) ok = ok && thresh == c.thresh_out;
        else ok = ok && m.cost_mv == c.cost_mv;
        if (!ok)
        {
            if (bad < 10)
                fprintf(stderr, "%s frame %d pass %d mb %d pix %d ref %d: got mv (%d,%d) cost %d cost_mv %d thr %d, want (%d,%d) %d %d thr %d\n",
                        r.tag, c.frame, c.pass, c.mb_xy, c.i_pixel, c.i_ref, m.mv[0], m.mv[1], m.cost, m.cost_mv, thresh,
                        c.mv[0], c.mv[1], c.cost, c.cost_mv, c.thresh_out);
            bad++;
        }
    }
    if (getenv("PCAMV_EMU_STATS"))
    {
        const char *kn[4] = { "sad_qpel", "satd", "satd_chroma", "sad_fpel" };
        for (int k = 0; k < 4; k++)
            for (int n = 1; n <= 4; n++)
            {
                long tot = 0;
                for (int p = 0; p < 7; p++) tot += g_eval_calls[k][n][p];
                if (tot)
                    fprintf(stderr, "%-12s n=%d calls=%8ld  by pixel: %ld %ld %ld %ld %ld %ld %ld\n", kn[k], n, tot, g_eval_calls[k][n][0], g_eval_calls[k][n][1],
                            g_eval_calls[k][n][2], g_eval_calls[k][n][3], g_eval_calls[k][n][4], g_eval_calls[k][n][5], g_eval_calls[k][n][6]);
            }
    }
    printf("calls=%ld mismatches=%ld skipped=%ld\n", calls, bad, skipped);
    return bad ? 1 : 0;
}
