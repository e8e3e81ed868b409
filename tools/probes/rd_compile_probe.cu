// rd_compile_probe.cu — compile-only evidence that the RD pieces (csrc/pcamv_cavlc.cuh, csrc/pcamv_rd.cuh) and the intra mode analysis
// (csrc/pcamv_intra.cuh, second kernel) are device code:
// one macroblock-per-team kernel that runs a candidate through motion compensation, the kept levels, the product's residual path,
// the distortion and the CAVLC size, built for sm_100a by tests/test_emu_cavlc.py::test_rd_pieces_compile_for_sm_100a
// (nvcc -c; nothing launches it — the pieces are not on the product path).
#include "../../video-steganography-pcamv_b200/csrc/pcamv_device.h"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_recon.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_cavlc.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_rd.cuh"
#include "../../video-steganography-pcamv_b200/csrc/pcamv_intra.cuh"

using namespace pcamv;

struct RdCandidate { MbResult res; CavlcMb mb; int mb_xy, psy_rd, lambda, lambda2; };

__global__ void k_rd_cost_probe(DevFrameCtx fc, FrameParams fp, const CavlcSizes *sizes, const RdCandidate *cand, int n, RdLevels *levels, int *cost)
{
    __shared__ MbWork work;
    for (int i = blockIdx.x; i < n; i += gridDim.x)
    {
        MbCtx c(fc, fp, work);
        c.mb_xy = cand[i].mb_xy; c.mb_x = c.mb_xy % fc.mb_w; c.mb_y = c.mb_xy / fc.mb_w;
        PCAMV_FOR_ITEMS(it, 64)
            st4a(work.fenc_y + 4 * it, ld4a(fc.fenc_y + (size_t)(16 * c.mb_y + (it >> 2)) * fc.stride_y + 16 * c.mb_x + 4 * (it & 3)));
        PCAMV_FOR_ITEMS(it, 32)
        {
            const int pl = it >> 4, y = (it >> 1) & 7, x = (it & 1) << 2;
            st4a((pl ? work.fenc_v : work.fenc_u) + 8 * y + x, ld4a((pl ? fc.fenc_v : fc.fenc_u) + (size_t)(8 * c.mb_y + y) * fc.stride_c + 8 * c.mb_x + x));
        }
        team_sync();
        rd_mc_inter(c, cand[i].res);
        team_sync();
        if (team_lane() == 0) rd_levels_mb(c, levels[i]);
        team_sync();
        encode_mb_residual(c);
        team_sync();
        if (team_lane() == 0)
        {
            CavlcMb m = cand[i].mb;
            m.coef = levels[i].coef; m.chroma_dc = levels[i].chroma_dc; m.cbp_luma = levels[i].cbp_luma; m.cbp_chroma = levels[i].cbp_chroma;
            for (int k = 0; k < 26; k++) m.coded[k] = levels[i].coded[k];
            const int ssd = rd_distortion_mb(work.fenc_y, work.pred_y, work.fenc_u, work.pred_u, work.fenc_v, work.pred_v, cand[i].psy_rd, cand[i].lambda);
            cost[i] = ssd + ((cavlc_mb_inter_bits(*sizes, m) * cand[i].lambda2 + 128) >> 8);
        }
        team_sync();
    }
}

// the intra mode analysis of one macroblock per team (lane 0 does the work: the pieces are one-lane code so far)
struct IntraJob { Intra4x4In in4; uint8_t fenc_u[64], fenc_v[64], border[33 + 34]; int has_left, has_top, has_topleft; };

__global__ void k_intra_probe(const IntraJob *jobs, int n, IntraCosts *costs, int *cost4, int *modes4)
{
    for (int i = blockIdx.x; i < n; i += gridDim.x)
        if (threadIdx.x == 0)
        {
            const IntraJob &j = jobs[i];
            IntraCosts o;
            intra_analyse_16x16(j.in4.fenc, j.has_left, j.has_top, j.has_topleft, j.border[0], j.border + 1, j.border + 17, j.in4.lambda, o);
            const int tl[2] = { j.border[33], j.border[50] };
            intra_analyse_chroma(j.fenc_u, j.fenc_v, j.has_left, j.has_top, j.has_topleft, tl, j.border + 34, j.border + 42, j.border + 51, j.border + 59, j.in4.lambda, o);
            costs[i] = o;
            int pred[16];
            cost4[i] = intra_analyse_4x4(j.in4, pred);
            for (int k = 0; k < 16; k++) modes4[16 * i + k] = pred[k];
        }
}
