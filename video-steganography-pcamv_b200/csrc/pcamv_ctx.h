// pcamv_ctx.h — the context object behind the opaque pcamv_ctx handle (private to csrc/).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "pcamv_device.h"
#include "pcamv_frame_types.h"

struct pcamv_ctx
{
    pcamv_cfg cfg;
    pcamv::DevFrameCtx fc;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;          // per-kernel timing of pcamv_frame_run
    std::string err;
    bool failed = false;
    long long launches = 0;

    // HBM
    uint8_t *d_fenc = nullptr;                 // Y | U | V, strides = stride_y / stride_c
    uint8_t *d_ref[PCAMV_SLOTS] = {};          // per slot: 4 luma planes | U | V (+slack)
    uint16_t *d_integral[PCAMV_SLOTS] = {}, *d_integral4[PCAMV_SLOTS] = {};
    size_t luma_bytes = 0, chroma_bytes = 0, ref_bytes = 0;
    int16_t *d_cost_mv = nullptr;              // 32769
    uint8_t *d_tables = nullptr;               // cost_ref | quant mf/bias | dequant
    // batch staging (stateless search seam)
    pcamv_me_call *d_calls = nullptr; pcamv_me_result *d_results = nullptr; int batch_cap = 0, batch_n = 0;
    pcamv_me_call *h_calls = nullptr; pcamv_me_result *h_results = nullptr;     // pinned
    uint8_t *h_stage = nullptr; size_t h_stage_bytes = 0;                       // pinned frame staging

    // frame-level analysis (pcamv_frame_api.cu): per-frame motion state and per-MB outputs, all in HBM
    pcamv::FrameArrays fa = {};                // type / ref8 / mv4 / mvr of the frame being analysed
    int8_t *d_col_ref8 = nullptr; uint32_t *d_col_mv4 = nullptr;      // co-located frame's ref / mv (temporal candidates)
    pcamv::ForcedMb *d_forced = nullptr;       // pass-2 forced decisions
    pcamv::LogEntry *d_log = nullptr;          // [n_mb][log_stride]
    int log_stride = PCAMV_LOG_MAX;            // entries per macroblock: what the configured search can produce at most
    pcamv::MbResult *d_mb_results = nullptr;   // [n_mb]
    pcamv::PartInfo *d_subparts = nullptr;     // [n_mb][16], only with sub-8x8 partitions enabled
    pcamv::BatchItem *d_batch = nullptr, *h_batch = nullptr; int batch_items_cap = 0;   // leader of a multi-context launch
    int *d_batch_claim = nullptr;
    unsigned char *d_split = nullptr; size_t split_bytes = 0; int split_warps = 0;     // split wavefront (rows_per_cta = -2): rings, requests, parked rows
    int batch_max_ctas = 0, batch_claim_cap = 0;   // persistent grid size of multi-context launches (0 = not yet computed)
    unsigned long long *d_mvsads = nullptr; int mvsads_cap = 0;    // --me tesa: per-row candidate lists
    unsigned long long *d_seam_mvsads = nullptr;                   // --me tesa, stateless search seam: PCAMV_SEAM_CHUNK lists
    uint8_t *d_stc_io = nullptr; size_t stc_io_bytes = 0;        // pcamv_stc_embed's staging of host cover / rho / stego
    uint8_t *d_stc = nullptr; size_t stc_bytes = 0;               // embed stage (pcamv_stc_embed): cover | stego | rho | elems | path | total
    // embed stage on the device (pcamv_embed.cu): one allocation carved into offsets | info.cache records | cover | stego | filp | rho | total
    uint8_t *d_emb = nullptr; int *emb_offsets = nullptr; void *emb_pass1 = nullptr; uint8_t *emb_cover = nullptr, *emb_stego = nullptr;
    int8_t *emb_filp = nullptr; float *emb_rho = nullptr; double *emb_total = nullptr;
    int emb_length = 0, emb_state = 0;         // 0 = nothing, 1 = cover / rho built for the frame in HBM, 2 = flips + forced decisions built
    uint16_t *d_recon_nnz = nullptr; uint8_t *d_recon_patches = nullptr; int recon_patch_cap = 0;   // device-side reconstruction (pcamv_recon.cu)
    unsigned long long *d_trace = nullptr;     // [n_mb][2] per-MB start/end timestamps (pcamv_frame_trace)
    bool trace_on = false;
    int *d_progress = nullptr;                 // [mb_h] row progress + [1] row claim counter + [mb_h] row owners (row pool)
    uint8_t *h_frame = nullptr; size_t h_frame_bytes = 0, h_frame_in_bytes = 0;
    bool dl_mbs_direct = false, dl_log_direct = false;             // pinned staging for frame inputs / outputs
    pcamv::FrameParams fp[3] = {};             // parameters of the last uploaded frame, per pass (0 / 1 / 2)
    bool frame_ready[3] = { false, false, false }, frame_cost_table = false;
    int frame_last = -1;
    // results while the wavefront is still running (pcamv_analyse_p_begin / pcamv_analyse_p_rows): a second stream polls the
    // row counters and copies finished macroblock rows out behind the kernel
    cudaStream_t side = nullptr; int *h_progress = nullptr;
    cudaStream_t sr_kstream = nullptr;         // the stream the kernel was launched on (the leader's, for a multi-context launch)
    pcamv_mb_out *sr_mbs = nullptr; pcamv_log_entry *sr_log = nullptr;
    int sr_rows = 0; bool sr_active = false;   // rows already in the caller's buffers; an analysis is being streamed
};

namespace pcamv {
int ctx_fail(pcamv_ctx *c, const char *what, cudaError_t e);
int filter_slot(pcamv_ctx *ctx, int slot);     // borders + half-pel planes (+ integral planes) of a slot whose integer interiors are in HBM
}
