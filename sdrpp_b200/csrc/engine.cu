// The front-end engine behind the C ABI (include/sdrpp_cuda.h): device IQ ring, spectrum framing,
// VFO registry/grouping, per-block launch sequencing, pinned result buffers.
//
// Mirrors the wiring of IQFrontEnd (signal_path/iq_frontend.cpp:15-228) and the per-VFO objects
// built by RxVFO::init (dsp/channel/rx_vfo.h:19-33), but as one stream-ordered launch sequence per
// IQ block instead of a thread per block.
#include "common.cuh"
#include "design.h"
#include "kernels.h"
#include "comm.h"
#include "../../include/sdrpp_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <unordered_map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

namespace sdrpp {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
static int fail(int code, const std::string& msg) { set_last_error(msg); return code; }

static uint64_t turns_to_u64(double turns) {
    // fractional turns in [-0.5, 0.5] -> 64-bit phase step (two's complement wrap)
    double f = turns - floor(turns);
    long double v = (long double)f * 18446744073709551616.0L;
    if (v >= 18446744073709551615.0L) return 0;
    return (uint64_t)v;
}

template <class T>
static cudaError_t dev_alloc(T** p, size_t n, bool zero = true) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    // cudaMemset on device memory returns before the fill has run, and the legacy default stream it runs on is not
    // ordered against this library's non-blocking streams: wait for it, or a kernel / copy enqueued right after the
    // allocation can be overtaken by the fill (seen as a VFO table zeroed after its upload, and as fp16 planes losing
    // their low half on the first tensor-core block). Allocation happens on configuration paths only.
    if (e == cudaSuccess && zero) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess && zero) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}

// Configuration-path fills and uploads that are complete on return: the synchronous runtime calls may return while the
// fill / the DMA out of the runtime's staging buffer is still queued on the legacy default stream, which this library's
// non-blocking streams are not ordered against.
static cudaError_t memset_sync(void* p, int v, size_t bytes) {
    cudaError_t e = cudaMemset(p, v, bytes);
    return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e;
}
static cudaError_t upload_sync(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e;
}

// ---------------------------------------------------------------------------------------------
// VFO plan: the stage list RxVFO::init would build for (inSR, outSR, bw)
// ---------------------------------------------------------------------------------------------
constexpr int kChanHistPad = 2048; // history pad of a channel FIR stage (taps - 1 <= 2048)

struct TailPlanStage {
    int type = TAIL_FIR, T = 1, D = 1, interp = 1;
    std::vector<float> taps; // FIR taps or polyphase bank
    float* d_taps = nullptr;
    uint32_t in_off = 0;     // slab offset of the input data area
    int cap_in = 0;
};

struct VfoPlan {
    double inSR = 0, outSR = 0, bw = 0;
    ResamplerPlan rp;
    bool filter_needed = false;
    int chan_taps = 0;
    // stage 1
    bool s1_fir = false;
    int s1_D = 1, s1_T = 1, s1_A = 1, s1_tap_off = -1;
    std::vector<float> s1_taps;
    // tensor-core stage 1 (channelizer_tc.cu): device copy of the first FIR, block exponent of the B image
    bool tc_ok = false;
    float* d_s1_taps = nullptr;
    int tc_escale = 0;
    std::vector<TailPlanStage> tail;
    uint32_t s1_off[2] = { 0, 0 }; // the two stage-1 output regions (data areas)
    uint32_t final_off = 0;
    int cap_final = 0;
    size_t slab_elems = 0;
    int max_block = 0;

    ~VfoPlan() { for (auto& s : tail) if (s.d_taps) cudaFree(s.d_taps); if (d_s1_taps) cudaFree(d_s1_taps); }
};

static int build_plan(VfoPlan& p, double inSR, double outSR, double bw, int max_block, std::string* err) {
    p.inSR = inSR; p.outSR = outSR; p.bw = bw; p.max_block = max_block;
    if (!(inSR > 0) || !(outSR > 0) || !(bw > 0)) { *err = "sample rates and bandwidth must be positive"; return SDRPP_ERR_ARG; }
    p.rp = design_resampler(inSR, outSR);
    p.filter_needed = (bw != outSR); // rx_vfo.h:24
    std::vector<DecimStage> dst;
    if (p.rp.mode == 0 || p.rp.mode == 1) {
        dst = decim_plan(p.rp.predec);
        if (dst.empty()) { *err = "no PowerDecimator plan for ratio " + std::to_string(p.rp.predec); return SDRPP_ERR_ARG; }
    }
    size_t first_tail = 0;
    if (!dst.empty()) {
        const int D = dst[0].decimation, T = dst[0].ntaps, A = stage1_A(T, D);
        const int tap_off = stage1_supported(A, D) ? stage1_tap_offset(p.rp.predec) : -1;
        if (tap_off >= 0) {
            p.s1_fir = true; p.s1_D = D; p.s1_T = T; p.s1_A = A; p.s1_tap_off = tap_off;
            p.s1_taps.assign(dst[0].taps, dst[0].taps + T);
            first_tail = 1;
        }
    }
    for (size_t i = first_tail; i < dst.size(); i++) {
        TailPlanStage s;
        s.type = TAIL_DECFIR; s.T = dst[i].ntaps; s.D = dst[i].decimation;
        s.taps.assign(dst[i].taps, dst[i].taps + dst[i].ntaps);
        p.tail.push_back(std::move(s));
    }
    if (p.rp.mode == 0 || p.rp.mode == 2) {
        TailPlanStage s;
        s.type = TAIL_POLY; s.D = p.rp.decim; s.interp = p.rp.interp;
        int tpp = 0;
        s.taps = build_polyphase_bank(p.rp.taps, p.rp.interp, &tpp);
        s.T = tpp;
        p.tail.push_back(std::move(s));
    }
    if (p.filter_needed) {
        TailPlanStage s;
        s.type = TAIL_FIR;
        const double fw = bw / 2.0; // RxVFO::generateTaps, rx_vfo.h:117-121
        s.taps = design_lowpass(fw, fw * 0.1, outSR);
        s.T = (int)s.taps.size();
        if (s.T < 1) { *err = "channel filter has no taps"; return SDRPP_ERR_ARG; }
        p.chan_taps = s.T;
        p.tail.push_back(std::move(s));
    }
    if ((int)p.tail.size() > kTailMaxStages) { *err = "too many stages"; return SDRPP_ERR_ARG; }
    // capacities and slab layout. The stage-1 output (= input of the first tail stage, or the final output when
    // there is no tail stage) is double-buffered so that the tail of block i can run concurrently with stage 1
    // of block i+1; each region is [history pad | data].
    long long cap = p.s1_fir ? (max_block / p.s1_D + 2) : max_block;
    uint32_t off = 0;
    {
        const int hist0 = p.tail.empty() ? 1 : p.tail[0].T - 1;
        if (hist0 > kChanHistPad) { *err = "filter longer than 2049 taps is not supported"; return SDRPP_ERR_ARG; }
        // the channel filter's pad does not depend on its length, so setBandwidth keeps the slab layout
        const uint32_t hc = (!p.tail.empty() && p.tail[0].type == TAIL_FIR) ? (uint32_t)kChanHistPad : (uint32_t)((hist0 + 1) & ~1);
        for (int r = 0; r < 2; r++) {
            p.s1_off[r] = off + hc;
            off = p.s1_off[r] + (uint32_t)((cap + 1) & ~1LL);
        }
    }
    for (size_t i = 0; i < p.tail.size(); i++) {
        TailPlanStage& s = p.tail[i];
        if (s.T - 1 > 2048) { *err = "filter longer than 2049 taps is not supported"; return SDRPP_ERR_ARG; }
        if (i > 0) {
            const uint32_t hc = (s.type == TAIL_FIR) ? (uint32_t)kChanHistPad : (uint32_t)((s.T - 1 + 1) & ~1);
            s.in_off = off + hc;
            off = s.in_off + (uint32_t)((cap + 1) & ~1LL);
        } else {
            s.in_off = p.s1_off[0];
        }
        s.cap_in = (int)cap;
        if (s.type == TAIL_DECFIR) cap = cap / s.D + 2;
        else if (s.type == TAIL_POLY) cap = cap * s.interp / s.D + 2;
    }
    p.cap_final = (int)cap;
    if (p.tail.empty()) {
        p.final_off = p.s1_off[0];
        p.slab_elems = off;
    } else {
        p.final_off = off + 2;
        p.slab_elems = (size_t)p.final_off + (size_t)((cap + 1) & ~1LL);
    }
    if (p.s1_fir && s1t_supported(p.s1_T, p.s1_D)) {
        if (dev_alloc(&p.d_s1_taps, p.s1_taps.size(), false) != cudaSuccess ||
            upload_sync(p.d_s1_taps, p.s1_taps.data(), p.s1_taps.size() * sizeof(float)) != cudaSuccess) {
            *err = std::string("tap upload: ") + cudaGetErrorString(cudaGetLastError());
            return SDRPP_ERR_CUDA;
        }
        p.tc_escale = s1t_b_exponent(p.s1_taps.data(), p.s1_T);
        p.tc_ok = true;
    }
    for (auto& s : p.tail) {
        if (dev_alloc(&s.d_taps, s.taps.size(), false) != cudaSuccess ||
            upload_sync(s.d_taps, s.taps.data(), s.taps.size() * sizeof(float)) != cudaSuccess) {
            *err = std::string("tap upload: ") + cudaGetErrorString(cudaGetLastError());
            return SDRPP_ERR_CUDA;
        }
    }
    return SDRPP_OK;
}

// Integer state of one group (identical for all members: same plan, same epoch).
struct GroupState {
    int s1_offset = 0;
    int st_offset[kTailMaxStages] = { 0 };
    int st_phase[kTailMaxStages] = { 0 };
    int64_t abs_valid = 0; // epoch: samples before this absolute index read as zero
    int64_t abs_out = 0;   // output samples produced since the epoch
};

// RDS side output of BroadcastFM (broadcast_fm.h:50-51): xlator.init(NULL, -57000.0, samplerate) and
// rdsResamp.init(NULL, samplerate, 5000.0). Taps are shared by every VFO of the same IF sample rate (at 250 kS/s the
// reference's plan rounds 7812.5 to 7813 and ends up with a 5000-phase bank of 593,750 taps: 2.4 MB).
struct RdsPlan {
    ResamplerPlan rp;
    std::vector<DecimStage> stages;
    float* d_taps[kRdsMaxStages] = { nullptr };
    float* d_bank = nullptr;
    int tpp = 0;
    uint64_t dphi = 0;
    ~RdsPlan() { for (float* t : d_taps) if (t) cudaFree(t); if (d_bank) cudaFree(d_bank); }
};

struct Vfo {
    bool alive = false;
    double outSR = 0, bw = 0, offset = 0;
    int demod = 0;
    std::shared_ptr<VfoPlan> plan;
    int group = -1;
    // NCO
    uint64_t phi_ref = 0; int64_t n_ref = 0; uint64_t dphi = 0; uint64_t dphi2 = 0;
    double turns = 0; // phase step in turns (for the folded taps)
    float2* slab = nullptr;
    uint32_t out_off = 0;
    int dev_index = -1;
    // post-detector stages
    sdrpp_cuda_post_cfg post{};
    float* post_state = nullptr;   // device: scalars | FIR history | work area
    float* post_taps = nullptr;    // device
    int post_ntaps = 0, post_hist_pad = 0;
    // BroadcastFM (POST_WFM): pilot band-pass (complex) in post_taps, audio low-pass in post_taps2
    float* post_taps2 = nullptr;
    int post_ntaps2 = 0, post_delay = 0, post_cap = 0;
    float pll_alpha = 0, pll_beta = 0, pll_min = 0, pll_max = 0;
    // RDS side output (post.wfm_rds): shared plan, own history buffers, the integer state of the reference's blocks
    std::shared_ptr<RdsPlan> rds;
    float2* rds_state = nullptr;
    uint32_t rds_buf_off[kRdsMaxStages] = { 0 }, rds_pbuf_off = 0, rds_out_off = 0;
    int rds_off[kRdsMaxStages] = { 0 }, rds_pphase = 0, rds_poff = 0, rds_out_cap = 0, rds_slot = -1;
    uint64_t rds_n = 0;            // discriminator samples translated so far (NCO phase = rds_n * dphi)
    // radio IF chain: device record of IF_FLOATS floats, null until first configured
    float* if_state = nullptr;
    // level / SNR read-out on every spectrum row: slot in the signal-info table, -1 = off
    bool sig_on = false;
    int sig_slot = -1;
    float sig_snr = 0.0f;              // smoothed SNR carried from row to row (waterfall.cpp:928-932)
    std::vector<float> sig_levels;     // last 10 levels (selectedVFO_LevelHistory)
};

struct Group {
    std::shared_ptr<VfoPlan> plan;
    int demod = 0;
    GroupState st;
    std::vector<int> members;
    int first_dev = 0;
    float4* d_G = nullptr;
    std::vector<float4> h_G;
    bool g_dirty = true;
    int last_n_final = 0;
    // tensor-core stage 1: B image (shifted taps x member phasors), rebuilt when a member retunes
    uint8_t* d_B = nullptr; size_t b_cap = 0;
    int tc_shift = -1, tc_A = 0;
    bool tc_dirty = true;
};

struct ResultSet {
    sdrpp_cf32* iq = nullptr;
    float* demod = nullptr;
    float* audio = nullptr;
    float* audio_r = nullptr;     // right channel of stereo demodulators
    float2* rds = nullptr;        // RDS side outputs (pinned), rows at rds_offs
    std::vector<int> rds_counts;  // per VFO id: RDS samples of this block (-1: output off)
    std::vector<uint32_t> rds_offs;
    std::vector<char> stereo;     // per VFO id: stereo demodulator at submit
    float* rows = nullptr;
    float* zoom = nullptr;
    float* hold = nullptr;        // peak-hold row after this block (zoom_out floats)
    float2* sig = nullptr;        // (strength, snr) per row and signal-info slot
    int nsig = 0;
    std::vector<int> sig_slot;    // per VFO id: slot at submit, -1 = off
    std::set<int> sig_done;       // VFO ids whose row-to-row recurrences (SNR smoothing, level history) were applied
    int nrows = 0;
    std::vector<int> counts;      // per VFO id
    std::vector<uint32_t> offs;   // per VFO id: arena offset of its rows WHEN THE BLOCK WAS SUBMITTED (the layout may change later)
    std::vector<int> demods;      // per VFO id: demod kind at submit (0: no demod row)
    std::vector<char> has_audio;
    cudaEvent_t done = nullptr;
    bool pending = false;
};

} // namespace sdrpp

using namespace sdrpp;

constexpr int kSets = 5; // result sets / raw staging buffers / result arenas: blocks the caller may have in flight
constexpr size_t kDescBytes = 256 * 1024; // per-slot descriptor arena (per-block kernel arguments)
constexpr size_t kMaxGraphs = 96;         // instantiated graphs kept per kind

struct sdrpp_cuda_frontend {
    int device = 0;
    sdrpp_cuda_frontend_cfg cfg{};
    double eff_sr = 0;
    cudaStream_t st = nullptr, st_copy = nullptr, st_fft = nullptr, st_tail = nullptr, st_s1b = nullptr, st_d2h = nullptr;
    cudaEvent_t ev_ingest = nullptr, ev_s1 = nullptr, ev_fft[2] = { nullptr, nullptr }, ev_tail[2] = { nullptr, nullptr };
    bool ev_fft_valid[2] = { false, false };
    cudaEvent_t ev_rows[kSets] = { nullptr };   // per result set: the block's spectrum rows (zoomed rows, level read-outs) are on the host
    // One submitting thread and one waiting thread may use a front end concurrently (submit / wait_input on one,
    // wait + result getters on the other); everything else is serialised by the caller. wait() drops the lock while
    // it blocks on the block's completion event.
    std::mutex api_mtx;
    int last_in_slot = -1;

    // multi-GPU feed (comm.cu): every submit is one ncclBroadcast of the raw block from comm_root on st_bcast
    sdrpp_cuda_comm* comm = nullptr;
    int comm_root = 0;
    cudaStream_t st_bcast = nullptr;
    cudaEvent_t ev_bcast[kSets] = { nullptr };
    cudaEvent_t ev_join = nullptr;
    cudaEvent_t ev_s1_fork = nullptr, ev_s1_join = nullptr;
    bool ev_tail_valid[2] = { false, false };
    long long blk = 0; // blocks processed (parity selects the stage-1 output region)
    long long launches = 0;
    std::string sticky;

    // input staging
    // kSets blocks may be in flight: the caller can submit block i+2 before waiting for block i, so that the
    // host-to-device copy of a block never waits for the results of an earlier one to reach the host
    void* h_stage[kSets] = { nullptr };
    void* d_raw[kSets] = { nullptr };
    cudaEvent_t ev_h2d[kSets] = { nullptr }, ev_consumed[kSets] = { nullptr };
    bool consumed_valid[kSets] = { false };
    size_t raw_cap = 0;
    long long seq = 0;

    // pre-processing
    std::vector<DecimStage> fe_stages;
    std::vector<float*> fe_taps;
    std::vector<float2*> fe_buf;   // per stage: [T-1 hist | max input]
    std::vector<int> fe_offset;
    float2* dc_in = nullptr; float2* dc_state = nullptr; float2* dc_scratch = nullptr;

    // ring
    float2* ring = nullptr; uint32_t ring_mask = 0; int ring_log2 = 0;
    int64_t abs_pos = 0; // absolute index of the next sample to be written
    int last_count = 0;  // post-preprocessing samples of the last block

    // spectrum
    int nz = 0, skip = 0;
    float* d_window = nullptr;
    float2* d_inter = nullptr; int inter_frames = 0;
    float* d_rows = nullptr; int rows_cap = 0;
    int64_t fft_next = 0;
    // waterfall zoom (fft_scaler): bin boundaries on the device, zoomed rows beside the raw rows
    int zoom_out = 0; bool zoom_ranged = false, zoom_keep_raw = true;
    double zoom_view[3] = { 0, 0, 0 };
    int* d_zoom_idx = nullptr; float* d_zoom = nullptr;
    // display state on the zoomed row (WaterFall::pushFFT): smoothing buffer, peak hold, latest row
    bool disp_smoothing = false, disp_hold = false;
    float disp_alpha = 0.5f, disp_hold_speed = 0.3f;
    float* d_disp = nullptr;      // [smooth | hold | latest] x zoom_out
    // level / SNR of selected VFOs on every raw row (WaterFall::calculateVFOSignalInfo)
    int4* d_sig_bins = nullptr; float2* d_sig = nullptr; int sig_cap = 0, nsig = 0; bool sig_dirty = true;
    bool snr_smoothing = false; float snr_alpha = 0.5f;

    // VFOs
    std::vector<Vfo> vfos;
    std::vector<Group> groups;
    std::map<std::tuple<double, double, double>, std::weak_ptr<VfoPlan>> plan_cache;
    bool layout_dirty = true;
    VfoDev* d_vfos = nullptr; int d_vfos_cap = 0;
    PostDev* d_post = nullptr; int post_active = 0; // per-VFO post-detector records (same indexing as d_vfos)
    float2* d_arena_iq = nullptr; float* d_arena_demod = nullptr; float* d_arena_audio = nullptr; float* d_arena_audio_r = nullptr;
    size_t arena_cap = 0, arena_used = 0;
    int post_stereo = 0;   // VFOs with a stereo demodulator (POST_WFM): the right-channel arena is copied back too
    // RDS side outputs: table of the VFOs that have one (device order), arena of their rows per result set
    std::map<long long, std::shared_ptr<RdsPlan>> rds_plans;   // by IF sample rate in milli-hertz
    std::vector<int> rds_ids;                  // VFO ids in table order
    RdsDev* d_rds_tab = nullptr; int rds_tab_cap = 0;
    float2* d_arena_rds = nullptr; size_t rds_arena_cap = 0, rds_arena_used = 0, rs_rds_cap = 0;

    // results
    ResultSet rs[kSets];
    size_t rs_arena_cap = 0; int rs_rows_cap = 0; int rs_rows_n = 0;
    int cur = -1; // result set of the last waited block
    long long waited = 0;
    bool readback = true;

    // tensor-core stage 1: fp16 hi/lo planes of the ring per first-stage decimation (index D / 64: 32 -> 0, 64 -> 1)
    int s1_mode = 0;        // 0: tensor cores where the plan allows, 1: FP32 FMA kernel only
    int tail_mode = 0;      // 0: low-latency tail kernel when the VFO set is smaller than the machine (default), 1: general
                            // tail kernel only (SDRPP_TAIL_MODE=general), 2: low-latency kernel wherever a block fits (=fast)
    int num_sms = 148;
    S1TPlanes tc_planes[2] = {};
    bool planes_skipped[2] = { false, false }; // a block went by without refreshing the planes (FP32-only mode)
    long long s1t_launches = 0;

    // command list of the block being planned, per-slot descriptors, instantiated graphs (launcher.h)
    Launcher L;
    unsigned char* h_desc[kSets] = { nullptr };   // pinned
    unsigned char* d_desc[kSets] = { nullptr };
    cudaStream_t st_desc = nullptr;               // descriptor uploads
    cudaEvent_t ev_desc = nullptr;
    bool run_replayable[GRAPH_KINDS] = { true, true, true };
    bool graphs_on = true;                        // SDRPP_GRAPHS=0: every block command by command
    struct GraphEntry { cudaGraphExec_t exec = nullptr; std::vector<unsigned char> sig; };
    std::unordered_map<uint64_t, GraphEntry> gcache[GRAPH_KINDS];
    std::unordered_map<uint64_t, int> gseen[GRAPH_KINDS];
    long long graph_replays = 0, graph_captures = 0, direct_runs = 0, capture_us = 0;
    cudaEvent_t ev_fftk = nullptr;                // join of the spectrum branch inside the stage-1 graph (fft_order 0)
    // Where the spectrum kernels of a block run (SDRPP_FFT_ORDER): 0 = a branch of the stage-1 graph, joined at its end;
    // 1 = a graph of their own on the spectrum stream behind stage 1 (beside the tail); 2 = a graph of their own right
    // behind ingest (ingest launched directly, beside the fp16 split and stage 1: the stream topology of round 1).
    int fft_order = -1;                           // -1: by where the block comes from (2 device-resident, 0 host), see process_block
    bool blk_device_src = false;                  // the block being planned was submitted from device memory
    bool fuse_ingest = true;                      // SDRPP_FUSE_INGEST=0: always the separate ingest kernel

    // timeline probe (SDRPP_TIMELINE=1, tools/timeline_probe.py): timing events on the streams the kernels really run on,
    // for the last kTlBlocks blocks; blocks run command by command while it is on
    static constexpr int kTlBlocks = 32, kTlMarks = 10;
    bool timeline = false;
    cudaEvent_t tl_ev[kTlBlocks][kTlMarks] = {};
    long long tl_blk[kTlBlocks] = {};
    // profiling
    bool profiling = false;
    cudaEvent_t pev[5] = { nullptr };
    float kernel_ms[4] = { 0, 0, 0, 0 };
    bool pev_valid = false;
};

namespace sdrpp {

#define FE_TRY(fe, expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            (fe)->sticky = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
            set_last_error((fe)->sticky);                                                        \
            return SDRPP_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

// SDRPP_DEBUG_ERR=1: report a CUDA error left behind by an earlier, unchecked runtime call (it would otherwise surface at the
// next kernel launch's cudaGetLastError and be blamed on that launch)
static void debug_stale(const char* where) {
    static const bool on = getenv("SDRPP_DEBUG_ERR") != nullptr;
    if (!on) return;
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) fprintf(stderr, "[sdrpp_cuda] stale CUDA error at %s: %s\n", where, cudaGetErrorString(e));
}

static int fe_check(sdrpp_cuda_frontend* fe) {
    if (!fe) return fail(SDRPP_ERR_ARG, "null front end");
    if (!fe->sticky.empty()) { set_last_error(fe->sticky); return SDRPP_ERR_CUDA; }
    cudaError_t e = cudaSetDevice(fe->device);
    if (e != cudaSuccess) return fail(SDRPP_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return SDRPP_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// ---- pre-processing configuration ----------------------------------------------------------
static int configure_preproc(sdrpp_cuda_frontend* fe) {
    for (float* t : fe->fe_taps) cudaFree(t);
    for (float2* b : fe->fe_buf) cudaFree(b);
    fe->fe_taps.clear(); fe->fe_buf.clear(); fe->fe_offset.clear(); fe->fe_stages.clear();
    const int ratio = fe->cfg.decim_ratio;
    if (ratio > 1) {
        fe->fe_stages = decim_plan(ratio);
        if (fe->fe_stages.empty()) return fail(SDRPP_ERR_ARG, "invalid decimation ratio (PowerDecimator::checkRatio)");
        int cap = fe->cfg.max_block;
        for (const DecimStage& s : fe->fe_stages) {
            float* t = nullptr; float2* b = nullptr;
            FE_TRY(fe, dev_alloc(&t, (size_t)s.ntaps, false));
            FE_TRY(fe, upload_sync(t, s.taps, sizeof(float) * s.ntaps));
            FE_TRY(fe, dev_alloc(&b, (size_t)(s.ntaps - 1) + (size_t)cap + 8));
            fe->fe_taps.push_back(t); fe->fe_buf.push_back(b); fe->fe_offset.push_back(0);
            cap = cap / s.decimation + 2;
        }
    }
    fe->eff_sr = fe->cfg.sample_rate / (double)std::max(1, ratio);
    if (!fe->dc_in) {
        FE_TRY(fe, dev_alloc(&fe->dc_in, (size_t)fe->cfg.max_block + 8));
        FE_TRY(fe, dev_alloc(&fe->dc_state, 1));
        FE_TRY(fe, dev_alloc(&fe->dc_scratch, (size_t)dc_block_chunks(fe->cfg.max_block) + 8));
    }
    return SDRPP_OK;
}

static int configure_zoom(sdrpp_cuda_frontend* fe) {
    if (fe->d_zoom_idx) { cudaFree(fe->d_zoom_idx); fe->d_zoom_idx = nullptr; }
    if (fe->d_zoom) { cudaFree(fe->d_zoom); fe->d_zoom = nullptr; }
    if (fe->d_disp) { cudaFree(fe->d_disp); fe->d_disp = nullptr; }
    for (int i = 0; i < kSets; i++) if (fe->rs[i].zoom) { cudaFreeHost(fe->rs[i].zoom); fe->rs[i].zoom = nullptr; }
    for (int i = 0; i < kSets; i++) if (fe->rs[i].hold) { cudaFreeHost(fe->rs[i].hold); fe->rs[i].hold = nullptr; }
    if (fe->zoom_out <= 0 || fe->cfg.fft_size <= 0 || fe->rows_cap <= 0) return SDRPP_OK;
    std::vector<int> idx;
    fe->zoom_ranged = zoom_indices(fe->zoom_view[0], fe->zoom_view[1], fe->zoom_view[2], fe->cfg.fft_size, fe->zoom_out, &idx);
    FE_TRY(fe, dev_alloc(&fe->d_zoom_idx, idx.size(), false));
    FE_TRY(fe, upload_sync(fe->d_zoom_idx, idx.data(), idx.size() * sizeof(int)));
    FE_TRY(fe, dev_alloc(&fe->d_zoom, (size_t)fe->rows_cap * fe->zoom_out, false));
    for (int i = 0; i < kSets; i++) FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].zoom, (size_t)fe->rows_cap * fe->zoom_out * sizeof(float)));
    // smoothing buffer, peak hold and latest row start at -1000 dB ("hide everything", waterfall.cpp:774-788)
    {
        std::vector<float> init((size_t)3 * fe->zoom_out, -1000.0f);
        FE_TRY(fe, dev_alloc(&fe->d_disp, init.size(), false));
        FE_TRY(fe, upload_sync(fe->d_disp, init.data(), init.size() * sizeof(float)));
        for (int i = 0; i < kSets; i++) FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].hold, (size_t)fe->zoom_out * sizeof(float)));
    }
    return SDRPP_OK;
}

// ---- spectrum configuration (IQFrontEnd::updateFFTPath / updateFFTSize) ------------------------
static int configure_fft(sdrpp_cuda_frontend* fe) {
    if (fe->d_window) { cudaFree(fe->d_window); fe->d_window = nullptr; }
    if (fe->d_inter) { cudaFree(fe->d_inter); fe->d_inter = nullptr; }
    if (fe->d_rows) { cudaFree(fe->d_rows); fe->d_rows = nullptr; }
    for (int i = 0; i < kSets; i++) if (fe->rs[i].rows) { cudaFreeHost(fe->rs[i].rows); fe->rs[i].rows = nullptr; }
    fe->rows_cap = 0; fe->rs_rows_cap = 0;
    const int N = fe->cfg.fft_size;
    if (N == 0) return SDRPP_OK;
    int N1, N2;
    if (spectrum_split(N, &N1, &N2) < 0) return fail(SDRPP_ERR_ARG, "fft_size must be a power of two in 64..4194304");
    if (!(fe->cfg.fft_rate > 0)) return fail(SDRPP_ERR_ARG, "fft_rate must be positive");
    if (fe->cfg.fft_window < 0 || fe->cfg.fft_window >= SDRPP_WIN_COUNT) return fail(SDRPP_ERR_ARG, "unknown window type");
    reshape_params(fe->eff_sr, N, fe->cfg.fft_rate, &fe->skip, &fe->nz);
    if (fe->nz < 1) return fail(SDRPP_ERR_ARG, "fft_rate too high for this sample rate");
    const long long interval = (long long)fe->nz + fe->skip;
    if ((long long)fe->nz + 3LL * fe->cfg.max_block + 4096 > (1LL << fe->ring_log2))
        return fail(SDRPP_ERR_ARG, "ring too small for this fft size / block size (needs nz + 3 * max_block + 4096 samples)");
    std::vector<float> w((size_t)fe->nz + 2);
    design_window(fe->cfg.fft_window, w.data(), fe->nz, true);
    FE_TRY(fe, dev_alloc(&fe->d_window, (size_t)fe->nz, false));
    FE_TRY(fe, upload_sync(fe->d_window, w.data(), sizeof(float) * fe->nz));
    int rows = fe->cfg.max_fft_rows > 0 ? fe->cfg.max_fft_rows : (int)(fe->cfg.max_block / interval + 2);
    fe->rows_cap = rows;
    // one set of rows per result set: the copy of block i's rows to the host may still be crossing the link while the spectrum
    // kernels of the following blocks run (the main stream waits for spectrum KERNELS only, never for a row copy)
    FE_TRY(fe, dev_alloc(&fe->d_rows, (size_t)kSets * rows * N, false));
    if (N1 > 1) {
        // frames per launch group: keep the four-step intermediate within ~32 MB so it stays in L2
        fe->inter_frames = std::max(1, std::min(rows, (int)((32u << 20) / ((size_t)N * 8))));
        FE_TRY(fe, dev_alloc(&fe->d_inter, (size_t)fe->inter_frames * N, false));
    } else {
        fe->inter_frames = rows;
    }
    for (int i = 0; i < kSets; i++) FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].rows, (size_t)rows * N * sizeof(float)));
    fe->rs_rows_cap = rows;
    fe->fft_next = fe->abs_pos;
    return configure_zoom(fe);
}

// ---- VFO layout ------------------------------------------------------------------------------
static void fill_g_for_vfo(Group& g, int lane_index, const Vfo& v) {
    // F[p] = exp(+j*2*pi*turns*p): the NCO advance over p samples inside one row of D samples
    const VfoPlan& p = *g.plan;
    for (int k = 0; k < p.s1_D; k++) {
        double ang = v.turns * (double)k;
        ang -= floor(ang);
        const float re = (float)cos(2.0 * kPi * ang), im = (float)sin(2.0 * kPi * ang);
        size_t idx; int half;
        stage1_g_index(p.s1_A, p.s1_D, lane_index, k, &idx, &half);
        float4& e = g.h_G[idx];
        if (half == 0) { e.x = re; e.y = im; } else { e.z = re; e.w = im; }
    }
}

static int rebuild_layout(sdrpp_cuda_frontend* fe) {
    // A new layout makes new command sequences (grids, table and arena addresses): the instantiated graphs of the old one can
    // never match again, and left in the cache they would fill it (kMaxGraphs) after a few dozen control calls and leave every
    // later block to the command-by-command path. The streams are idle here (control calls quiesce before they mark the layout).
    for (int k = 0; k < GRAPH_KINDS; k++) {
        for (auto& e : fe->gcache[k]) if (e.second.exec) cudaGraphExecDestroy(e.second.exec);
        fe->gcache[k].clear();
        fe->gseen[k].clear();
    }
    // device VFO table ordered by group; arena offsets by VFO
    int total = 0;
    size_t arena = 0;
    for (Group& g : fe->groups) { g.first_dev = total; total += (int)g.members.size(); }
    std::vector<VfoDev> h((size_t)std::max(total, 1));
    std::vector<PostDev> hp((size_t)std::max(total, 1));
    fe->post_active = 0;
    fe->post_stereo = 0;
    for (Group& g : fe->groups) {
        for (size_t i = 0; i < g.members.size(); i++) {
            Vfo& v = fe->vfos[(size_t)g.members[i]];
            v.dev_index = g.first_dev + (int)i;
            v.out_off = (uint32_t)arena;
            arena += (size_t)((g.plan->cap_final + 3) & ~3);
            VfoDev& d = h[(size_t)v.dev_index];
            d.slab = v.slab; d.phi_ref = v.phi_ref; d.n_ref = v.n_ref; d.dphi = v.dphi; d.dphi2 = v.dphi2;
            d.out_off = v.out_off; d.pad = 0; d.ifs = v.if_state;
            PostDev& pd = hp[(size_t)v.dev_index];
            pd = PostDev{};
            if (v.post.enabled && v.post_state) {
                pd.kind = v.demod == SDRPP_DEMOD_QUADRATURE ? (v.post.wfm ? POST_WFM : POST_FM) : v.demod == SDRPP_DEMOD_AM ? POST_AM : POST_SSB;
                pd.mode = pd.kind == POST_FM ? (v.post.fm_lowpass != 0) : pd.kind == POST_AM ? v.post.am_agc_mode : (v.post.ssb_agc != 0);
                if (pd.kind == POST_WFM) {
                    pd.mode = (v.post.wfm_stereo ? 1 : 0) | (v.post.fm_lowpass ? 2 : 0);
                    pd.taps2 = v.post_taps2; pd.ntaps2 = v.post_ntaps2; pd.delay = v.post_delay; pd.cap = v.post_cap;
                    pd.pll_alpha = v.pll_alpha; pd.pll_beta = v.pll_beta; pd.pll_min_freq = v.pll_min; pd.pll_max_freq = v.pll_max;
                }
                pd.ntaps = v.post_ntaps; pd.hist_pad = v.post_hist_pad; pd.taps = v.post_taps; pd.state = v.post_state;
                pd.out_off = v.out_off;
                // AGC::init(NULL, 1.0, attack, decay, 10e6, 10.0, INFINITY) (am.h:32-33, ssb.h:27); coefficients as floats (agc.h:22-33)
                pd.attack = (float)v.post.agc_attack; pd.inv_attack = 1.0f - pd.attack;
                pd.decay = (float)v.post.agc_decay; pd.inv_decay = 1.0f - pd.decay;
                pd.dc_rate = (float)v.post.dc_block_rate; pd.set_point = 1.0f; pd.max_gain = (float)10e6; pd.max_out = 10.0f;
                fe->post_active++;
                if (pd.kind == POST_WFM) fe->post_stereo++;
            }
        }
        if (g.g_dirty) g.tc_dirty = true;
        if (g.plan->s1_fir && g.g_dirty) {
            const size_t n = stage1_g_elems(g.plan->s1_A, g.plan->s1_D, (int)g.members.size());
            g.h_G.assign(n, make_float4(0.f, 0.f, 0.f, 0.f));
            for (size_t i = 0; i < g.members.size(); i++) fill_g_for_vfo(g, (int)i, fe->vfos[(size_t)g.members[i]]);
            if (g.d_G) { FE_TRY(fe, cudaStreamSynchronize(fe->st)); cudaFree(g.d_G); g.d_G = nullptr; }
            FE_TRY(fe, dev_alloc(&g.d_G, n, false));
            FE_TRY(fe, cudaMemcpyAsync(g.d_G, g.h_G.data(), n * sizeof(float4), cudaMemcpyHostToDevice, fe->st));
            FE_TRY(fe, cudaStreamSynchronize(fe->st));
            g.g_dirty = false;
        }
    }
    if (total > fe->d_vfos_cap) {
        if (fe->d_vfos) { FE_TRY(fe, cudaStreamSynchronize(fe->st)); cudaFree(fe->d_vfos); cudaFree(fe->d_post); }
        fe->d_vfos_cap = std::max(total, 64);
        FE_TRY(fe, dev_alloc(&fe->d_vfos, (size_t)fe->d_vfos_cap));
        FE_TRY(fe, dev_alloc(&fe->d_post, (size_t)fe->d_vfos_cap));
    }
    if (total > 0) {
        FE_TRY(fe, cudaMemcpyAsync(fe->d_vfos, h.data(), sizeof(VfoDev) * (size_t)total, cudaMemcpyHostToDevice, fe->st));
        FE_TRY(fe, cudaMemcpyAsync(fe->d_post, hp.data(), sizeof(PostDev) * (size_t)total, cudaMemcpyHostToDevice, fe->st));
        FE_TRY(fe, cudaStreamSynchronize(fe->st));
    }
    // RDS side outputs: table in device order, rows packed in an arena of their own
    {
        std::vector<RdsDev> tab;
        fe->rds_ids.clear();
        size_t rarena = 0;
        for (Group& g : fe->groups)
            for (int id : g.members) {
                Vfo& v = fe->vfos[(size_t)id];
                v.rds_slot = -1;
                if (!(v.post.enabled && v.post_state && v.post.wfm && v.post.wfm_rds && v.rds && v.rds_state && v.demod == SDRPP_DEMOD_QUADRATURE)) continue;
                RdsDev d{};
                d.in_off = v.out_off; d.out_off = (uint32_t)rarena; v.rds_out_off = d.out_off;
                rarena += (size_t)v.rds_out_cap;
                d.nstages = (int)v.rds->stages.size();
                for (int s = 0; s < d.nstages; s++) {
                    d.T[s] = v.rds->stages[(size_t)s].ntaps; d.D[s] = v.rds->stages[(size_t)s].decimation;
                    d.taps[s] = v.rds->d_taps[s]; d.buf_off[s] = v.rds_buf_off[s];
                }
                d.interp = v.rds->rp.interp; d.decim = v.rds->rp.decim; d.tpp = v.rds->tpp; d.bank = v.rds->d_bank; d.pbuf_off = v.rds_pbuf_off;
                d.state = v.rds_state; d.dphi = v.rds->dphi;
                v.rds_slot = (int)tab.size();
                tab.push_back(d);
                fe->rds_ids.push_back(id);
            }
        fe->rds_arena_used = rarena;
        if ((int)tab.size() > fe->rds_tab_cap) {
            FE_TRY(fe, cudaStreamSynchronize(fe->st));
            cudaFree(fe->d_rds_tab); fe->d_rds_tab = nullptr;
            fe->rds_tab_cap = std::max((int)tab.size(), 16);
            FE_TRY(fe, dev_alloc(&fe->d_rds_tab, (size_t)fe->rds_tab_cap));
        }
        if (!tab.empty()) {
            FE_TRY(fe, cudaMemcpyAsync(fe->d_rds_tab, tab.data(), sizeof(RdsDev) * tab.size(), cudaMemcpyHostToDevice, fe->st));
            FE_TRY(fe, cudaStreamSynchronize(fe->st));
        }
        if (rarena > fe->rds_arena_cap) {
            FE_TRY(fe, cudaStreamSynchronize(fe->st));
            cudaFree(fe->d_arena_rds); fe->d_arena_rds = nullptr;
            fe->rds_arena_cap = rarena + rarena / 2 + 256;
            FE_TRY(fe, dev_alloc(&fe->d_arena_rds, kSets * fe->rds_arena_cap));
        }
        if (rarena > fe->rs_rds_cap) {
            FE_TRY(fe, cudaStreamSynchronize(fe->st));
            for (int i = 0; i < kSets; i++) {
                fe->rs[i].rds_counts.clear();
                if (fe->rs[i].rds) cudaFreeHost(fe->rs[i].rds);
                fe->rs[i].rds = nullptr;
                FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].rds, fe->rds_arena_cap * sizeof(float2)));
            }
            fe->rs_rds_cap = fe->rds_arena_cap;
        }
    }
    fe->arena_used = arena;
    if (arena > fe->arena_cap) {
        FE_TRY(fe, cudaStreamSynchronize(fe->st));
        if (fe->d_arena_iq) cudaFree(fe->d_arena_iq);
        if (fe->d_arena_demod) cudaFree(fe->d_arena_demod);
        if (fe->d_arena_audio) cudaFree(fe->d_arena_audio);
        if (fe->d_arena_audio_r) cudaFree(fe->d_arena_audio_r);
        fe->arena_cap = arena + arena / 2 + 1024;
        // one result arena per result set: the device-to-host copy of block i runs beside the tail of block i+1
        FE_TRY(fe, dev_alloc(&fe->d_arena_iq, kSets * fe->arena_cap));
        FE_TRY(fe, dev_alloc(&fe->d_arena_demod, kSets * fe->arena_cap));
        FE_TRY(fe, dev_alloc(&fe->d_arena_audio, kSets * fe->arena_cap));
        FE_TRY(fe, dev_alloc(&fe->d_arena_audio_r, kSets * fe->arena_cap));
    }
    if (arena > fe->rs_arena_cap) {
        FE_TRY(fe, cudaStreamSynchronize(fe->st));
        for (int i = 0; i < kSets; i++) {
            fe->rs[i].counts.clear(); // results of earlier blocks do not survive a growth of the result arena
            if (fe->rs[i].iq) cudaFreeHost(fe->rs[i].iq);
            if (fe->rs[i].demod) cudaFreeHost(fe->rs[i].demod);
            if (fe->rs[i].audio) cudaFreeHost(fe->rs[i].audio);
            if (fe->rs[i].audio_r) cudaFreeHost(fe->rs[i].audio_r);
            FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].audio_r, fe->arena_cap * sizeof(float)));
            FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].iq, fe->arena_cap * sizeof(sdrpp_cf32)));
            FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].demod, fe->arena_cap * sizeof(float)));
            FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].audio, fe->arena_cap * sizeof(float)));
        }
        fe->rs_arena_cap = fe->arena_cap;
    }
    fe->layout_dirty = false;
    return SDRPP_OK;
}

static void remove_from_group(sdrpp_cuda_frontend* fe, int id) {
    Vfo& v = fe->vfos[(size_t)id];
    if (v.group < 0) return;
    Group& g = fe->groups[(size_t)v.group];
    g.members.erase(std::remove(g.members.begin(), g.members.end(), id), g.members.end());
    g.g_dirty = true;
    v.group = -1;
    // drop empty groups (and fix up indices)
    for (size_t gi = 0; gi < fe->groups.size();) {
        if (fe->groups[gi].members.empty()) {
            if (fe->groups[gi].d_G) { cudaStreamSynchronize(fe->st); cudaFree(fe->groups[gi].d_G); }
            if (fe->groups[gi].d_B) { cudaStreamSynchronize(fe->st); cudaFree(fe->groups[gi].d_B); }
            fe->groups.erase(fe->groups.begin() + (long)gi);
            for (Vfo& o : fe->vfos) if (o.alive && o.group > (int)gi) o.group--;
        } else gi++;
    }
    fe->layout_dirty = true;
}

// Put a VFO into the group of its (plan, demod, epoch); a fresh epoch = zero history from now on.
static void join_group(sdrpp_cuda_frontend* fe, int id, int64_t epoch) {
    Vfo& v = fe->vfos[(size_t)id];
    for (size_t gi = 0; gi < fe->groups.size(); gi++) {
        Group& g = fe->groups[gi];
        if (g.plan == v.plan && g.demod == v.demod && g.st.abs_valid == epoch && g.st.abs_out == 0 && epoch == fe->abs_pos) {
            g.members.push_back(id); g.g_dirty = true; v.group = (int)gi; fe->layout_dirty = true;
            return;
        }
    }
    Group g;
    g.plan = v.plan; g.demod = v.demod; g.st.abs_valid = epoch;
    g.members.push_back(id);
    fe->groups.push_back(std::move(g));
    v.group = (int)fe->groups.size() - 1;
    fe->layout_dirty = true;
}

// Put a VFO whose filter state continues (setBandwidth) into a group with exactly this integer state.
static void join_group_with_state(sdrpp_cuda_frontend* fe, int id, const GroupState& st) {
    Vfo& v = fe->vfos[(size_t)id];
    for (size_t gi = 0; gi < fe->groups.size(); gi++) {
        Group& g = fe->groups[gi];
        if (g.plan == v.plan && g.demod == v.demod && memcmp(&g.st, &st, sizeof(GroupState)) == 0) {
            g.members.push_back(id); g.g_dirty = true; v.group = (int)gi; fe->layout_dirty = true;
            return;
        }
    }
    Group g;
    g.plan = v.plan; g.demod = v.demod; g.st = st;
    g.members.push_back(id);
    fe->groups.push_back(std::move(g));
    v.group = (int)fe->groups.size() - 1;
    fe->layout_dirty = true;
}

static void set_nco(sdrpp_cuda_frontend* fe, Vfo& v, double offset, bool keep_phase) {
    // phase continuity on retune (frequency_xlator.h:25-29): phase at the next input sample is kept
    const int64_t now = fe->abs_pos;
    uint64_t phi_now = 0;
    if (keep_phase) phi_now = v.phi_ref + (uint64_t)(now - v.n_ref) * v.dphi;
    double turns;
    xlator_increment(-offset, fe->eff_sr, nullptr, nullptr, &turns); // RxVFO: xlator.init(NULL, -_offset, inSR), rx_vfo.h:27
    v.turns = turns;
    v.dphi = turns_to_u64(turns);
    v.phi_ref = phi_now; v.n_ref = now;
    v.offset = offset;
    fe->sig_dirty = true;
    // SSB second translation (demod/ssb.h:119-126) at the output rate
    double tr = 0.0;
    if (v.demod == SDRPP_DEMOD_USB) tr = v.bw / 2.0;
    else if (v.demod == SDRPP_DEMOD_LSB) tr = -v.bw / 2.0;
    double t2;
    xlator_increment(tr, v.outSR, nullptr, nullptr, &t2);
    v.dphi2 = turns_to_u64(t2);
}

static int get_plan(sdrpp_cuda_frontend* fe, double outSR, double bw, std::shared_ptr<VfoPlan>* out) {
    auto key = std::make_tuple(fe->eff_sr, outSR, bw);
    auto it = fe->plan_cache.find(key);
    if (it != fe->plan_cache.end()) {
        if (auto sp = it->second.lock()) { *out = sp; return SDRPP_OK; }
    }
    auto sp = std::make_shared<VfoPlan>();
    std::string err;
    int rc = build_plan(*sp, fe->eff_sr, outSR, bw, fe->cfg.max_block, &err);
    if (rc != SDRPP_OK) return fail(rc, err);
    fe->plan_cache[key] = sp;
    *out = sp;
    return SDRPP_OK;
}

// BroadcastFM::init with rdsOut (broadcast_fm.h:50-51): the plan (shared per IF sample rate) and this VFO's fresh state.
static int apply_rds(sdrpp_cuda_frontend* fe, Vfo& v) {
    const long long key = (long long)llround(v.outSR * 1000.0);
    std::shared_ptr<RdsPlan> pl;
    auto it = fe->rds_plans.find(key);
    if (it != fe->rds_plans.end()) pl = it->second;
    else {
        pl = std::make_shared<RdsPlan>();
        pl->rp = design_resampler(v.outSR, 5000.0);
        if (pl->rp.mode == 0 || pl->rp.mode == 1) pl->stages = decim_plan(pl->rp.predec);
        if ((int)pl->stages.size() > kRdsMaxStages) return fail(SDRPP_ERR_ARG, "BroadcastFM RDS output: the IF sample rate needs more PowerDecimator stages than supported");
        for (size_t s = 0; s < pl->stages.size(); s++) {
            FE_TRY(fe, dev_alloc(&pl->d_taps[s], (size_t)pl->stages[s].ntaps, false));
            FE_TRY(fe, upload_sync(pl->d_taps[s], pl->stages[s].taps, sizeof(float) * (size_t)pl->stages[s].ntaps));
        }
        if (pl->rp.mode == 0 || pl->rp.mode == 2) {
            const std::vector<float> bank = build_polyphase_bank(pl->rp.taps, pl->rp.interp, &pl->tpp);
            FE_TRY(fe, dev_alloc(&pl->d_bank, bank.size(), false));
            FE_TRY(fe, upload_sync(pl->d_bank, bank.data(), bank.size() * sizeof(float)));
        }
        double turns = 0;
        xlator_increment(-57000.0, v.outSR, nullptr, nullptr, &turns);
        pl->dphi = turns_to_u64(turns);
        fe->rds_plans[key] = pl;
    }
    v.rds = pl;
    // history buffers [T-1 | block] per stage, sized for the VFO's largest block
    size_t elems = 0;
    long long cap = v.plan->cap_final + 8;
    for (size_t s = 0; s < pl->stages.size(); s++) {
        v.rds_buf_off[s] = (uint32_t)elems;
        elems += (size_t)(pl->stages[s].ntaps - 1) + (size_t)cap + 8;
        cap = cap / pl->stages[s].decimation + 2;
    }
    if (pl->tpp > 0) {
        v.rds_pbuf_off = (uint32_t)elems;
        elems += (size_t)(pl->tpp - 1) + (size_t)cap + 8;
        cap = (cap * pl->rp.interp) / pl->rp.decim + 4;
    }
    v.rds_out_cap = (int)((cap + 3) & ~3LL);
    FE_TRY(fe, dev_alloc(&v.rds_state, elems + 8));   // zeroed: the blocks' cleared buffers
    for (int& o : v.rds_off) o = 0;
    v.rds_pphase = 0; v.rds_poff = 0; v.rds_n = 0;
    return SDRPP_OK;
}

// (Re)build the post-detector objects of a VFO for its current (demod, outSR, bw): the demodulators' init()
// (fm.h:25-44, am.h:27-44, ssb.h:21-36). Filter and AGC state start from reset.
static int apply_post(sdrpp_cuda_frontend* fe, Vfo& v) {
    cudaFree(v.post_state); cudaFree(v.post_taps); cudaFree(v.post_taps2);
    v.post_state = nullptr; v.post_taps = nullptr; v.post_taps2 = nullptr; v.post_ntaps = 0; v.post_hist_pad = 0; v.post_ntaps2 = 0;
    cudaFree(v.rds_state); v.rds_state = nullptr; v.rds.reset(); v.rds_slot = -1;
    fe->layout_dirty = true;
    if (!v.post.enabled) return SDRPP_OK;
    if (v.demod == SDRPP_DEMOD_NONE) return fail(SDRPP_ERR_STATE, "post-detector stages need a demodulator front end");
    if (v.demod == SDRPP_DEMOD_QUADRATURE && v.post.wfm) {
        // BroadcastFM::init (broadcast_fm.h:35-65)
        const std::vector<float> pilot = design_bandpass_complex(18750.0, 19250.0, 3000.0, v.outSR, true);
        const std::vector<float> audio = design_lowpass(15000.0, 4000.0, v.outSR);
        const int Tp = (int)pilot.size() / 2, Ta = (int)audio.size();
        if (Tp < 3 || Tp > 2048 || Ta < 1 || Ta > 2048) return fail(SDRPP_ERR_ARG, "BroadcastFM: the IF sample rate gives an unusable pilot / audio filter (needs roughly 100 kS/s .. 1.6 MS/s)");
        v.post_ntaps = Tp; v.post_ntaps2 = Ta; v.post_hist_pad = 0;
        v.post_delay = ((Tp - 1) / 2) + 1;
        v.post_cap = v.plan->cap_final + 8;
        pll_critically_damped((float)(25000.0 / v.outSR), &v.pll_alpha, &v.pll_beta);
        v.pll_min = (float)(2.0 * kPi * (18750.0 / v.outSR)); v.pll_max = (float)(2.0 * kPi * (19250.0 / v.outSR));
        const size_t n = 16 + 4 * (size_t)v.post_cap + (size_t)(Tp - 1) + (size_t)v.post_delay + 2 * (size_t)(Ta - 1) + 4 * (size_t)v.post_cap + 16;
        FE_TRY(fe, dev_alloc(&v.post_state, n));
        FE_TRY(fe, dev_alloc(&v.post_taps, pilot.size(), false));
        FE_TRY(fe, upload_sync(v.post_taps, pilot.data(), pilot.size() * sizeof(float)));
        FE_TRY(fe, dev_alloc(&v.post_taps2, audio.size(), false));
        FE_TRY(fe, upload_sync(v.post_taps2, audio.data(), audio.size() * sizeof(float)));
        // PLL::reset: phase = initPhase (0), freq = initFreq = hzToRads(19000, samplerate) as a float
        float init[2] = { 0.0f, (float)(2.0 * kPi * (19000.0 / v.outSR)) };
        FE_TRY(fe, upload_sync(v.post_state, init, sizeof(init)));
        if (v.post.wfm_rds) {
            const int rc = apply_rds(fe, v);
            if (rc != SDRPP_OK) return rc;
        }
        return SDRPP_OK;
    }
    std::vector<float> taps;
    const bool fm = v.demod == SDRPP_DEMOD_QUADRATURE, am = v.demod == SDRPP_DEMOD_AM;
    if ((fm && v.post.fm_lowpass) || am) {
        const double fw = v.bw / 2.0; // lowPass(bandwidth / 2, (bandwidth / 2) * 0.1, samplerate): fm.h:121-123, am.h:35
        taps = design_lowpass(fw, fw * 0.1, v.outSR);
        if (taps.empty() || taps.size() > 2048) return fail(SDRPP_ERR_ARG, "post-detector low-pass must have 1..2048 taps");
    }
    if (am && (v.post.am_agc_mode < 0 || v.post.am_agc_mode > 2)) return fail(SDRPP_ERR_ARG, "am_agc_mode must be 0 (off), 1 (carrier) or 2 (audio)");
    v.post_ntaps = (int)taps.size();
    v.post_hist_pad = (std::max(v.post_ntaps - 1, 0) + 3) & ~3;
    const size_t n = 16 + (size_t)v.post_hist_pad + (size_t)v.plan->cap_final + 8;
    FE_TRY(fe, dev_alloc(&v.post_state, n));
    if (!taps.empty()) {
        FE_TRY(fe, dev_alloc(&v.post_taps, taps.size(), false));
        FE_TRY(fe, upload_sync(v.post_taps, taps.data(), taps.size() * sizeof(float)));
    }
    // AGC::init(.., maxGain = 10e6, .., initGain = INFINITY): amp = setPoint / initGain = 0, gain = min(initGain, maxGain)
    float init[5] = { 0.0f, (float)10e6, 0.0f, (float)10e6, 0.0f };
    if (v.post.agc_gain > 0.0f) init[1] = v.post.agc_gain; // setAGCGain (am.h:69-73, ssb.h)
    FE_TRY(fe, upload_sync(v.post_state, init, sizeof(init)));
    return SDRPP_OK;
}

// Signal-info table: one (minSide, min, max, maxSide) bin record per VFO with the read-out enabled, rebuilt when a VFO is
// retuned, resized, enabled or removed. centerOffset / bandwidth are the VFO's own (what its WaterfallVFO carries),
// wholeBandwidth the effective sample rate (gui::waterfall.setBandwidth, core.cpp:52-55).
static int refresh_signal_info(sdrpp_cuda_frontend* fe) {
    if (!fe->sig_dirty) return SDRPP_OK;
    std::vector<int4> bins;
    for (Vfo& v : fe->vfos) {
        v.sig_slot = -1;
        if (!v.alive || !v.sig_on) continue;
        int b[4];
        signal_info_bins(v.offset, v.bw, fe->eff_sr, fe->cfg.fft_size, b);
        v.sig_slot = (int)bins.size();
        bins.push_back(make_int4(b[0], b[1], b[2], b[3]));
    }
    fe->nsig = (int)bins.size();
    const int need = fe->nsig * std::max(fe->rows_cap, 1);
    if (fe->nsig > 0 && (need > fe->sig_cap || !fe->d_sig_bins)) {
        FE_TRY(fe, cudaStreamSynchronize(fe->st_fft));
        cudaFree(fe->d_sig_bins); cudaFree(fe->d_sig);
        fe->d_sig_bins = nullptr; fe->d_sig = nullptr;
        FE_TRY(fe, dev_alloc(&fe->d_sig_bins, (size_t)std::max(fe->nsig, 16), false));
        FE_TRY(fe, dev_alloc(&fe->d_sig, (size_t)std::max(need, 16), false));
        for (int i = 0; i < kSets; i++) {
            if (fe->rs[i].sig) cudaFreeHost(fe->rs[i].sig);
            fe->rs[i].sig = nullptr;
            FE_TRY(fe, cudaMallocHost((void**)&fe->rs[i].sig, (size_t)std::max(need, 16) * sizeof(float2)));
        }
        fe->sig_cap = std::max(need, 16);
    }
    if (fe->nsig > 0) {
        FE_TRY(fe, cudaStreamSynchronize(fe->st_fft));
        FE_TRY(fe, upload_sync(fe->d_sig_bins, bins.data(), bins.size() * sizeof(int4)));
    }
    fe->sig_dirty = false;
    return SDRPP_OK;
}

// ---- one block -------------------------------------------------------------------------------
static void advance_decim(int& offset, int D, int count, int* nout) {
    // for (; offset < count; offset += D) out++; offset -= count;   (decimating_fir.h:51-62)
    int n = 0;
    if (offset < count) n = (count - offset + D - 1) / D;
    offset = offset + n * D - count;
    *nout = n;
}


// ---- execution of a planned block (launcher.h) ---------------------------------------------------------------------
static uint64_t fnv1a(const unsigned char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Commands [i, j) carry the same graph tag. Replay the instantiated graph of exactly this sequence if there is one;
// instantiate one (stream capture of the very commands) when the sequence shows up for the second time; run the commands
// one by one otherwise. The launch stream is the one the run's commands are ordered on: main for stage 1, tail for the tail.
static int run_tagged(sdrpp_cuda_frontend* fe, int kind, size_t i, size_t j) {
    Launcher& L = fe->L;
    bool replayable = fe->graphs_on && fe->run_replayable[kind];
    for (size_t k = i; k < j && replayable; k++) if (L.cmds[k].type == Cmd::CALL) replayable = false;
    if (replayable) {
        const unsigned char* sig = reinterpret_cast<const unsigned char*>(&L.cmds[i]);
        const size_t nbytes = (j - i) * sizeof(Cmd);
        const uint64_t h = fnv1a(sig, nbytes);
        cudaStream_t origin = L.streams[kind == GRAPH_S1 ? SID_MAIN : kind == GRAPH_FFT ? SID_FFT : SID_TAIL];
        auto it = fe->gcache[kind].find(h);
        if (it != fe->gcache[kind].end() && it->second.sig.size() == nbytes && memcmp(it->second.sig.data(), sig, nbytes) == 0) {
            FE_TRY(fe, cudaGraphLaunch(it->second.exec, origin));
            fe->graph_replays++;
            return SDRPP_OK;
        }
        if (it == fe->gcache[kind].end() && ++fe->gseen[kind][h] >= 2 && fe->gcache[kind].size() < kMaxGraphs) {
            cudaGraph_t graph = nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            FE_TRY(fe, cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = cudaSuccess;
            for (size_t k = i; k < j && e == cudaSuccess; k++) e = L.exec(L.cmds[k]);
            const cudaError_t e2 = cudaStreamEndCapture(origin, &graph);
            if (e == cudaSuccess) e = e2;
            sdrpp_cuda_frontend::GraphEntry ge;
            if (e == cudaSuccess) e = cudaGraphInstantiate(&ge.exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                // not capturable after all: remember not to try again and run the commands directly
                cudaGetLastError();
                fe->gseen[kind][h] = -(1 << 30);
            } else {
                ge.sig.assign(sig, sig + nbytes);
                fe->capture_us += (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
                FE_TRY(fe, cudaGraphLaunch(ge.exec, origin));
                fe->gcache[kind].emplace(h, std::move(ge));
                fe->gseen[kind].erase(h);
                fe->graph_captures++;
                return SDRPP_OK;
            }
        }
        if (fe->gseen[kind].size() > 4096) fe->gseen[kind].clear();
    }
    for (size_t k = i; k < j; k++) FE_TRY(fe, L.exec(L.cmds[k]));
    fe->direct_runs++;
    return SDRPP_OK;
}

static int execute_block(sdrpp_cuda_frontend* fe) {
    Launcher& L = fe->L;
    if (L.desc_used > L.desc_cap) return fail(SDRPP_ERR_STATE, "block descriptor overflow");
    // The per-block descriptor travels on a stream of its own, off the block's critical path: the slot's previous block is
    // done (submit waited for it), so the copy needs no ordering against kernels and is over long before the main stream gets
    // to this block. As the first node of the stage-1 graph it cost ~6 us per step (copy latency in front of ingest).
    if (L.desc_used > 0) {
        FE_TRY(fe, cudaMemcpyAsync(L.d_desc, L.h_desc, (L.desc_used + 255) & ~(size_t)255, cudaMemcpyHostToDevice, fe->st_desc));
        FE_TRY(fe, cudaEventRecord(fe->ev_desc, fe->st_desc));
        FE_TRY(fe, cudaStreamWaitEvent(L.streams[SID_MAIN], fe->ev_desc, 0));
    }
    const size_t n = L.cmds.size();
    size_t i = 0;
    while (i < n) {
        const int kind = L.cmds[i].graph;
        if (kind == GRAPH_NONE) { FE_TRY(fe, L.exec(L.cmds[i])); i++; continue; }
        size_t j = i;
        while (j < n && L.cmds[j].graph == kind) j++;
        const int rc = run_tagged(fe, kind, i, j);
        if (rc != SDRPP_OK) return rc;
        i = j;
    }
    L.deferred = false;
    return SDRPP_OK;
}

static int process_block(sdrpp_cuda_frontend* fe, int fmt, const void* d_in, int count, ResultSet& rs, float scale = 1.0f) {
    const bool prof = fe->profiling;
    debug_stale("process_block entry");
    // PLAN: every command of the block goes into the launcher's list (launcher.h); execute_block() at the end runs it,
    // as instantiated graphs where the same sequence has been seen before. Profiling serialises everything on the main stream.
    Launcher& L = fe->L;
    const int slot_d = (int)(&rs - fe->rs);
    L.begin_block(fe->h_desc[slot_d], fe->d_desc[slot_d], kDescBytes);
    L.deferred = true;
    fe->run_replayable[GRAPH_S1] = fe->run_replayable[GRAPH_TAIL] = fe->run_replayable[GRAPH_FFT] = true;
    // Default: a block that arrives over the host link hides the join of order 0 behind its copy and gains from the smaller
    // number of submissions (end to end 5.0 -> 5.3 GS/s); a device-resident block runs the spectrum beside the split and the
    // start of stage 1 (order 2: 6.2 -> 6.5 GS/s). profiles/r2l_graph_ab.txt.
    const int fft_order = prof ? 0 : fe->fft_order >= 0 ? fe->fft_order : (fe->blk_device_src ? 2 : 0);
    const long long kernels0 = L.kernels;
    const int st = SID_MAIN;
    if (prof) FE_TRY(fe, L.record(st, fe->pev[0]));
    // The spectrum is the one reader of the ring that is not ordered on `st`: before block i overwrites ring samples,
    // the spectrum work of block i-2 must be done. Frames of block i-1 may still be in flight; they reach back at most
    // nz + max_block samples from the end of block i-1, so a ring of nz + 3*max_block samples (what create() sizes and
    // configure_fft checks) can never be overwritten under a frame that is still being read.
    if (!prof && fe->cfg.fft_size > 0 && fe->ev_fft_valid[fe->blk & 1]) FE_TRY(fe, L.wait(st, fe->ev_fft[fe->blk & 1]));

    const int par = (int)(fe->blk & 1);
    if (!prof && fe->ev_tail_valid[par]) FE_TRY(fe, L.wait(st, fe->ev_tail[par])); // stage-1 region `par` was last read by the tail of block i-2

    auto tl_mark = [&](int sid, int k) -> int {
        if (!fe->timeline || prof) return SDRPP_OK;
        const int slot = (int)(fe->blk % sdrpp_cuda_frontend::kTlBlocks);
        fe->tl_blk[slot] = fe->blk;
        FE_TRY(fe, L.record(sid, fe->tl_ev[slot][k]));
        return SDRPP_OK;
    };
    if (int rc = tl_mark(st, 0); rc != SDRPP_OK) return rc;
    // ---- GRAPH_S1: ingest, [spectrum kernels (forked),] fp16 split, stage 1 (the descriptor is uploaded by execute_block) --
    L.cur_graph = (prof || fft_order == 2) ? GRAPH_NONE : GRAPH_S1;

    // ---- pre-processing chain: [decim] -> [dc block] -> [conjugate] -> ring (iq_frontend.cpp:30-37)
    const RingRef ring{ fe->ring, fe->ring_mask };
    const uint32_t wpos = (uint32_t)((uint64_t)fe->abs_pos & fe->ring_mask);
    const bool conj = fe->cfg.invert_iq != 0;
    const bool dc = fe->cfg.dc_blocking != 0;
    int n = count;
    // A cf32 block that needs no conversion is not copied to the ring by a kernel of its own: the fp16 split of the
    // tensor-core stage 1 reads it in place and writes the ring on the way (launch_s1t_split with `raw`). Decided below, when
    // the split is planned; until then the plain ingest is pending.
    bool plain_pending = false;
    if (fe->fe_stages.empty() && !dc) {
        plain_pending = true;
    } else {
        const RingRef lin_dc{ fe->dc_in, 0xFFFFFFFFu };
        if (fe->fe_stages.empty()) {
            FE_TRY(fe, launch_ingest(L, st, fmt, d_in, count, lin_dc, 0, false, scale));
        } else {
            const size_t ns = fe->fe_stages.size();
            RingRef first{ fe->fe_buf[0], 0xFFFFFFFFu };
            FE_TRY(fe, launch_ingest(L, st, fmt, d_in, count, first, (uint32_t)(fe->fe_stages[0].ntaps - 1), false, scale));
            for (size_t s = 0; s < ns; s++) {
                const DecimStage& ds = fe->fe_stages[s];
                int off = fe->fe_offset[s], nout = 0;
                const int off0 = off;
                advance_decim(off, ds.decimation, n, &nout);
                const bool last = (s + 1 == ns);
                RingRef dst; uint32_t pos; bool cj = false;
                if (!last) { dst = RingRef{ fe->fe_buf[s + 1], 0xFFFFFFFFu }; pos = (uint32_t)(fe->fe_stages[s + 1].ntaps - 1); }
                else if (dc) { dst = lin_dc; pos = 0; }
                else { dst = ring; pos = wpos; cj = conj; }
                // the front-end decimator and the DC blocker keep their by-value launches: a block that runs them is executed
                // command by command (Launcher::call), never replayed as a graph
                {
                    const float2* buf = fe->fe_buf[s]; const float* tp = fe->fe_taps[s];
                    const int T = ds.ntaps, D = ds.decimation, nin = n;
                    FE_TRY(fe, L.call(st, [=](cudaStream_t cs) {
                        cudaError_t e = launch_decim_stage(buf, tp, T, D, off0, nout, dst, pos, cj, cs);
                        return e != cudaSuccess ? e : launch_shift_history(const_cast<float2*>(buf), T - 1, nin, cs);
                    }));
                }
                fe->launches += (nout > 0 ? 1 : 0) + ((ds.ntaps > 1 && n > 0) ? 1 : 0);
                fe->fe_offset[s] = off;
                n = nout;
            }
        }
        if (dc) {
            const float rate = (float)(50.0 / fe->eff_sr); // IQFrontEnd::genDCBlockRate, iq_frontend.h:52-54
            {
                float2* din = fe->dc_in; float2* dst8 = fe->dc_state; float2* dsc = fe->dc_scratch; const int nn = n;
                FE_TRY(fe, L.call(st, [=](cudaStream_t cs) { long long dummy = 0; return launch_dc_block(din, nn, rate, dst8, dsc, ring, wpos, conj, cs, &dummy); }));
                fe->launches += 3;
            }
        }
    }
    const int64_t abs_block = fe->abs_pos;
    fe->abs_pos += n;
    fe->last_count = n;
    // ---- channelizer ----------------------------------------------------------------------------
    debug_stale("before rebuild_layout");
    if (fe->layout_dirty) { int rc = rebuild_layout(fe); if (rc != SDRPP_OK) return rc; }
    debug_stale("after rebuild_layout");
    int total_vfos_all = 0;
    for (const Group& g : fe->groups) total_vfos_all += (int)g.members.size();
    std::vector<TailArgs> tails, tails_fast;
    // The low-latency tail kernel (all 1024 threads of an SM on one VFO) is chosen when the VFO set is smaller than the machine:
    // there one CTA's chain of round trips is what a step waits for. With several CTAs per SM the tail is bound by its
    // instruction count and the general kernel is ahead. SDRPP_TAIL_MODE pins one: general / fast / narrow -- narrow is the
    // low-latency kernel with 256 threads and <= 41 KB, built to run BESIDE the persistent stage-1 CTA of the next block
    // (with SDRPP_S1T_SMEM_CAP=190000 and -DSDRPP_S1T_MAXNREG=80); measured: both kernels then run at half speed, they
    // compete for shared-memory bandwidth (profiles/r2m_coresidency.txt), so it is not a default.
    const int fast_threads = fe->tail_mode == 3 ? 256 : 1024;
    std::vector<int> tail_totals, tail_fast_totals;
    // Tensor-core stage 1: refresh the fp16 hi/lo planes of the ring for every first-stage decimation in use
    // (one conversion serves all VFOs and plans of that decimation), then collect the eligible groups per plane set.
    S1TArgs tc_args[2] = {};
    if (fe->s1_mode != 0 && n > 0) fe->planes_skipped[0] = fe->planes_skipped[1] = true;
    if (fe->s1_mode == 0 && n > 0) {
        for (int pi = 0; pi < 2; pi++) {
            const int D = pi ? 64 : 32;
            const uint32_t ngroups = (uint32_t)(((uint64_t)fe->ring_mask + 1) / (uint64_t)(8 * D));
            if (ngroups < 64) continue;
            // Row origin: every group's windows start on its own lattice n0 + m*D; rows aligned so that the starts
            // fall early in a row keep the shifted tap matrix at ceil(T/D) rows instead of one more.
            S1TPlanes& pl = fe->tc_planes[pi];
            long long best_cost = -1, cur_cost = -1; int best = 0; bool any = false;
            for (int o = 0; o < D; o += 4) {
                long long cost = 0;
                for (const Group& g : fe->groups) {
                    const VfoPlan& p = *g.plan;
                    if (!p.tc_ok || p.s1_D != D) continue;
                    any = true;
                    const int64_t first = abs_block - (p.s1_T - 1) + g.st.s1_offset;
                    const int sh = (int)(((first - o) % D + D) % D);
                    cost += (long long)g.members.size() * s1t_A(p.s1_T, D, sh);
                }
                if (!any) break;
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = o; }
                if (pl.hi && o == pl.origin) cur_cost = cost;
            }
            if (!any) continue;
            int64_t split_from = abs_block;
            if (!pl.hi) {
                const size_t bytes = ((size_t)fe->ring_mask + 1) * 4;
                FE_TRY(fe, dev_alloc(&pl.hi, bytes));
                FE_TRY(fe, dev_alloc(&pl.lo, bytes));
                FE_TRY(fe, dev_alloc(&pl.sinv, (size_t)ngroups));
                pl.D = D; pl.group_mask = ngroups - 1; pl.origin = best;
                pl.valid_from = abs_block;       // nothing before this block has been converted
            } else if (best_cost < cur_cost) {
                // new origin: the history the next windows reach back into has to be converted again
                pl.origin = best;
                split_from = std::max<int64_t>(0, abs_block - 4096);
                pl.valid_from = std::max(pl.valid_from, split_from);
            } else if (pl.valid_from > abs_block) {
                pl.valid_from = abs_block;
            }
            if (fe->planes_skipped[pi]) {
                // blocks went by without a conversion (FP32-only mode): the rows before this block are stale
                pl.valid_from = std::max(pl.valid_from, abs_block);
                fe->planes_skipped[pi] = false;
            }
            tc_args[pi].pl = pl;
            if (plain_pending && fe->fuse_ingest && fmt == SDRPP_FMT_CF32 && scale == 1.0f && (((uintptr_t)d_in) & 15u) == 0 &&
                s1t_split_covers(pl, split_from, abs_block, abs_block + n)) {
                FE_TRY(fe, launch_s1t_split(L, st, ring, pl, split_from, abs_block + n, d_in, abs_block, conj));
                plain_pending = false;
                continue;
            }
            if (plain_pending) { FE_TRY(fe, launch_ingest(L, st, fmt, d_in, count, ring, wpos, conj, scale)); plain_pending = false; }
            FE_TRY(fe, launch_s1t_split(L, st, ring, pl, split_from, abs_block + n));
        }
    }
    if (plain_pending) { FE_TRY(fe, launch_ingest(L, st, fmt, d_in, count, ring, wpos, conj, scale)); plain_pending = false; }
    if (int rc = tl_mark(st, 1); rc != SDRPP_OK) return rc;
    if (prof) FE_TRY(fe, L.record(st, fe->pev[1]));
    // Three consumers of the ring run concurrently (the reference's Splitter fan-out): the spectrum on st_fft,
    // stage 1 on st, and the tail of the PREVIOUS block on st_tail. Profiling serialises everything on st so
    // that per-family CUDA-event times are those of the kernels alone.
    const int sf = prof ? st : SID_FFT;
    const int stl = prof ? st : SID_TAIL;
    const int aset = (int)(&rs - fe->rs); // result set (and result arena) of this block
    if (!prof && fft_order != 1) {
        FE_TRY(fe, L.record(st, fe->ev_ingest));
        FE_TRY(fe, L.wait(sf, fe->ev_ingest));
    }
    // fft_order 1 / 2: the spectrum commands go to a list of their own (a GRAPH_FFT run on the spectrum stream), spliced in
    // behind stage 1 (1) or right here (2)
    std::vector<Cmd> main_cmds;
    const int graph_before_fft = L.cur_graph;
    if (fft_order != 0) { main_cmds.swap(L.cmds); L.cur_graph = GRAPH_FFT; }

    if (int rc = tl_mark(sf, 3); rc != SDRPP_OK) return rc;
    // ---- spectrum frames completed by this block (reshaper.h:102-129 keep/skip + handler) --------
    // Their kernels are a branch of the stage-1 graph; the device-to-host copies of the rows follow the graph (a copy
    // inside it would hold the next block's stage 1 back until 4 MB per row have crossed the link).
    struct PendingCopy { void* dst; const void* src; size_t bytes; };
    std::vector<PendingCopy> fft_copies;
    auto fft_copy = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (prof) return L.memcpy_async(st, dst, src, bytes, cudaMemcpyDeviceToHost);
        if (fft_order != 0) { const int g = L.cur_graph; L.cur_graph = GRAPH_NONE; const cudaError_t e = L.memcpy_async(sf, dst, src, bytes, cudaMemcpyDeviceToHost); L.cur_graph = g; return e; }
        fft_copies.push_back(PendingCopy{ dst, src, bytes });
        return cudaSuccess;
    };
    rs.nrows = 0;
    float* const rows_dev = fe->d_rows ? fe->d_rows + (size_t)aset * fe->rows_cap * fe->cfg.fft_size : nullptr;
    if (fe->cfg.fft_size > 0) {
        const int N = fe->cfg.fft_size;
        const int64_t interval = (int64_t)fe->nz + fe->skip;
        int frames = 0;
        const int64_t first = fe->fft_next;
        while (fe->fft_next + fe->nz <= fe->abs_pos && frames < fe->rows_cap) { frames++; fe->fft_next += interval; }
        // frames beyond the row buffer are dropped (like a waterfall that cannot keep up)
        while (fe->fft_next + fe->nz <= fe->abs_pos) fe->fft_next += interval;
        for (int f0 = 0; f0 < frames; f0 += fe->inter_frames) {
            SpectrumArgs a{};
            a.in = fe->ring; a.ring_mask = fe->ring_mask;
            a.start = (uint32_t)((uint64_t)(first + (int64_t)f0 * interval) & fe->ring_mask);
            a.frame_stride = (uint32_t)interval;
            a.nz = fe->nz; a.window = fe->d_window; a.inter = fe->d_inter;
            a.rows = rows_dev + (size_t)f0 * N; a.X = nullptr;
            a.frames = std::min(fe->inter_frames, frames - f0);
            FE_TRY(fe, launch_spectrum(L, sf, N, a, nullptr));
        }
        rs.nrows = frames;
    }
    // the ring is free again (for the main stream, two blocks on) once the spectrum KERNELS have read it
    if (!prof && fft_order != 0) {
        const int g = L.cur_graph; L.cur_graph = GRAPH_NONE;
        FE_TRY(fe, L.record(sf, fe->ev_fft[par]));
        L.cur_graph = g;
        fe->ev_fft_valid[par] = true;
    }
    // zoom / display state / level read-out write buffers that the copies behind the graph read (not double-buffered): such a
    // block runs command by command, where the spectrum stream orders kernels and copies
    if (fft_order == 0 && rs.nrows > 0 && (fe->zoom_out > 0 || fe->nsig > 0 || fe->sig_dirty)) fe->run_replayable[GRAPH_S1] = false;
    if (rs.nrows > 0 && fe->zoom_out > 0) {
        FE_TRY(fe, launch_fft_zoom(L, sf, rows_dev, fe->cfg.fft_size, rs.nrows, fe->d_zoom_idx, fe->zoom_out, fe->zoom_ranged, fe->d_zoom));
        if (fe->disp_smoothing || fe->disp_hold) {
            const int W = fe->zoom_out;
            FE_TRY(fe, launch_fft_display(L, sf, fe->d_zoom, W, rs.nrows, fe->disp_smoothing, fe->disp_alpha, fe->d_disp, fe->disp_hold, fe->disp_hold_speed,
                                          fe->d_disp + W, fe->d_disp + 2 * W));
            if (fe->readback && fe->disp_hold) FE_TRY(fe, fft_copy(rs.hold, fe->d_disp + W, (size_t)W * sizeof(float)));
        }
        if (fe->readback)
            FE_TRY(fe, fft_copy(rs.zoom, fe->d_zoom, (size_t)rs.nrows * fe->zoom_out * sizeof(float)));
    }
    if (int rc = tl_mark(sf, 4); rc != SDRPP_OK) return rc;
    rs.nsig = 0;
    rs.sig_done.clear();
    if (rs.nrows > 0 && fe->cfg.fft_size > 0) {
        int rc = refresh_signal_info(fe);
        if (rc != SDRPP_OK) return rc;
        if (fe->nsig > 0) {
            FE_TRY(fe, launch_signal_info(L, sf, rows_dev, fe->cfg.fft_size, rs.nrows, fe->d_sig_bins, fe->nsig, fe->d_sig));
            if (fe->readback)
                FE_TRY(fe, fft_copy(rs.sig, fe->d_sig, (size_t)rs.nrows * fe->nsig * sizeof(float2)));
            rs.nsig = fe->nsig;
            rs.sig_slot.assign(fe->vfos.size(), -1);
            for (size_t id = 0; id < fe->vfos.size(); id++) if (fe->vfos[id].alive && fe->vfos[id].sig_on) rs.sig_slot[id] = fe->vfos[id].sig_slot;
        }
    }
    if (fe->readback && rs.nrows > 0 && (fe->zoom_keep_raw || fe->zoom_out <= 0))
        FE_TRY(fe, fft_copy(rs.rows, rows_dev, (size_t)rs.nrows * fe->cfg.fft_size * sizeof(float)));
    if (prof) { FE_TRY(fe, L.record(st, fe->pev[2])); fe->ev_fft_valid[0] = fe->ev_fft_valid[1] = false; }
    else if (fft_order == 0) FE_TRY(fe, L.record(sf, fe->ev_fftk));   // the spectrum branch joins the main stream again at the end of the stage-1 graph
    std::vector<Cmd> fft_cmds;
    if (fft_order != 0) {
        L.cur_graph = GRAPH_NONE;
        FE_TRY(fe, L.record(sf, fe->ev_rows[aset]));
        fft_cmds.swap(L.cmds);
        L.cmds.swap(main_cmds);
        if (fft_order == 2) { L.cmds.insert(L.cmds.end(), fft_cmds.begin(), fft_cmds.end()); fft_cmds.clear(); L.cur_graph = GRAPH_S1; }
        else L.cur_graph = graph_before_fft;
    }

    // ---- channelizer: stage 1 of every group (the layout, the fp16 planes and their conversion were planned above) ----
    auto flush_tc = [&](int pi) -> int {
        if (tc_args[pi].ngroups == 0) return SDRPP_OK;
        FE_TRY(fe, launch_s1t(L, st, tc_args[pi], fe->num_sms));
        fe->s1t_launches++;
        tc_args[pi].ngroups = 0;
        return SDRPP_OK;
    };
    // Stage-1 launches of different groups are independent: alternate them between two streams so that the
    // partially filled last wave of one grid is topped up by the next grid's CTAs.
    // The second stream joins in only when a launch actually goes to it (in steady state every group runs in the one
    // tensor-core launch on `st`): saves three runtime calls per block.
    const bool fork = fe->groups.size() > 1;
    bool forked = false;
    if (fork) FE_TRY(fe, L.record(st, fe->ev_s1_fork));
    int s1_launch = 0;
    for (Group& g : fe->groups) {
        const VfoPlan& p = *g.plan;
        const int s1s = (fork && (s1_launch & 1) && !prof) ? SID_S1B : st;
        Stage1Args a{};
        a.ring = ring;
        a.nvfo = (int)g.members.size();
        a.vfos = fe->d_vfos + g.first_dev;
        a.abs_valid = g.st.abs_valid;
        a.out_off = p.s1_off[par];
        int nprev = 0;
        if (p.s1_fir) {
            int off = g.st.s1_offset;
            const int off0 = off;
            advance_decim(off, p.s1_D, n, &nprev);
            g.st.s1_offset = off;
            a.D = p.s1_D; a.A = p.s1_A;
            a.M = nprev; a.G = g.d_G;
            a.abs_first = abs_block - (p.s1_T - 1) + off0;
            // tensor cores when the whole window lies after the group's epoch (history before it reads as zero,
            // which only the FP32 kernel does)
            const int pi = p.s1_D >> 6;
            if (p.tc_ok && tc_args[pi].pl.hi && nprev > 0 && a.abs_first >= g.st.abs_valid && a.abs_first >= tc_args[pi].pl.origin &&
                a.abs_first >= tc_args[pi].pl.valid_from) {
                const int64_t rel_first = a.abs_first - tc_args[pi].pl.origin;
                const int shift = (int)(rel_first % p.s1_D);
                const int A = s1t_A(p.s1_T, p.s1_D, shift);
                const size_t need = s1t_b_bytes(A, p.s1_D, a.nvfo);
                if (need > g.b_cap) {
                    if (g.d_B) { FE_TRY(fe, cudaStreamSynchronize(fe->st)); cudaFree(g.d_B); g.d_B = nullptr; }
                    FE_TRY(fe, dev_alloc(&g.d_B, need, false));
                    g.b_cap = need; g.tc_dirty = true;
                }
                if (g.tc_dirty || shift != g.tc_shift || A != g.tc_A) {
                    FE_TRY(fe, launch_s1t_build_b(L, st, g.d_B, a.vfos, a.nvfo, p.d_s1_taps, p.s1_T, p.s1_D, shift, A, p.tc_escale));
                    g.tc_shift = shift; g.tc_A = A; g.tc_dirty = false;
                }
                if (tc_args[pi].ngroups == kS1TMaxGroups) { int rc = flush_tc(pi); if (rc != SDRPP_OK) return rc; }
                S1TGroupArgs& tg = tc_args[pi].g[tc_args[pi].ngroups++];
                tg = S1TGroupArgs{};
                tg.bblob = g.d_B; tg.vfos = a.vfos; tg.nvfo = a.nvfo; tg.A = A; tg.M = nprev;
                tg.row_first = rel_first / p.s1_D;
                tg.row0 = tg.row_first & ~(int64_t)7;
                tg.n_ttiles = (int)((tg.row_first - tg.row0 + nprev + 119) / 120);
                tg.out_off = a.out_off;
                tg.b_scale_inv = (float)std::ldexp(1.0, -p.tc_escale);
                goto stage1_done;
            }
            // keep the window start even (16-byte aligned in the ring): if it is odd, start one sample earlier
            // and use the tap table with a leading zero
            const int pad = (int)(a.abs_first & 1);
            a.abs_first -= pad;
            a.T = p.s1_T + pad;
            a.tap_off = p.s1_tap_off + pad * p.s1_A * p.s1_D;
            a.ring_first = (uint32_t)((uint64_t)a.abs_first & fe->ring_mask);
            if (s1s != st && !forked) { FE_TRY(fe, L.wait(SID_S1B, fe->ev_s1_fork)); forked = true; }
            FE_TRY(fe, launch_stage1(L, s1s, a));
        } else {
            nprev = n;
            a.D = 1; a.T = 1; a.A = 1; a.tap_off = 0; a.M = n; a.G = nullptr;
            a.abs_first = abs_block;
            a.ring_first = wpos;
            if (s1s != st && !forked) { FE_TRY(fe, L.wait(SID_S1B, fe->ev_s1_fork)); forked = true; }
            FE_TRY(fe, launch_mix_only(L, s1s, a));
        }
        if (nprev > 0) s1_launch++;
    stage1_done:

        TailGroup tg{};
        tg.first_vfo = g.first_dev; tg.nvfo = (int)g.members.size();
        tg.nstages = (int)p.tail.size();
        // a leading decimating FIR still sees 1/D of the input rate per VFO: it gets its own wide launch
        tg.s_begin = (!p.tail.empty() && p.tail[0].type == TAIL_DECFIR && tail_stage0_wide_supported(p.tail[0].T, p.tail[0].D)) ? 1 : 0;
        for (size_t s = 0; s < p.tail.size(); s++) {
            const TailPlanStage& ps = p.tail[s];
            TailStage& ts = tg.st[s];
            ts.type = ps.type; ts.T = ps.T; ts.D = ps.D; ts.interp = ps.interp; ts.taps = ps.d_taps;
            ts.in_off = (s == 0) ? p.s1_off[par] : ps.in_off;
            ts.n_in = nprev; ts.offset = g.st.st_offset[s]; ts.phase = g.st.st_phase[s];
            int nout = 0;
            if (ps.type == TAIL_DECFIR) {
                advance_decim(g.st.st_offset[s], ps.D, nprev, &nout);
            } else if (ps.type == TAIL_FIR) {
                nout = nprev;
            } else {
                // polyphase_resampler.h:75-93 in closed form
                const long long c = (long long)nprev - ts.offset;
                long long cnt = 0;
                if (c > 0) cnt = (c * ps.interp - ts.phase + ps.D - 1) / ps.D;
                const long long P = (long long)ts.phase + cnt * ps.D;
                g.st.st_offset[s] = (int)((long long)ts.offset + P / ps.interp - nprev);
                g.st.st_phase[s] = (int)(P % ps.interp);
                nout = (int)cnt;
            }
            ts.n_out = nout;
            nprev = nout;
        }
        tg.final_off = p.tail.empty() ? p.s1_off[par] : p.final_off;
        tg.carry0_off = p.s1_off[par ^ 1];
        tg.n_final = nprev; tg.demod = g.demod;
        tg.inv_dev = (float)(1.0 / (2.0 * kPi * ((p.bw / 2.0) / p.outSR))); // quadrature.h:21-28 with dev = bw/2 (fm.h:31)
        tg.abs_out = g.st.abs_out;
        g.st.abs_out += nprev;
        g.last_n_final = nprev;
        // Low-latency tail when this block's stage inputs fit in shared memory at once and no member carries a radio IF
        // chain (that lives in the general kernel); both kernels keep the same slab state, so the choice is per block.
        // Measured (profiles/r2e): with fewer VFOs than SMs the tail is bound by one CTA's chain of dependent round trips
        // and the low-latency kernel wins (100 WFM VFOs: 73 -> 58 us per step); with several CTAs per SM both kernels are
        // bound by instruction issue and the general one (tap tables built per segment, fewer instructions) is ahead.
        bool fast_ok = (fe->tail_mode == 0 ? total_vfos_all <= fe->num_sms : fe->tail_mode != 1);
        if (fast_ok) for (int id : g.members) if (fe->vfos[(size_t)id].if_state) { fast_ok = false; break; }
        bool fast = false;
        if (fast_ok) {
            // With the whole SM on one VFO even the first tail stage's input (9600 samples = 77 KB in cfg5) fits beside the
            // other regions: the wide first-stage kernel, its launch and one more round trip through L2 drop out of the chain
            // that every block of a small VFO set waits for (SDRPP_TAIL_FOLD=0: keep the wide kernel).
            static const bool fold = !(getenv("SDRPP_TAIL_FOLD") && getenv("SDRPP_TAIL_FOLD")[0] == '0');
            const int sb = tg.s_begin;
            if (fold && sb == 1 && fast_threads == 1024) { tg.s_begin = 0; fast = tail_fast_fits(tg, nullptr, fast_threads); if (!fast) tg.s_begin = sb; }
            if (!fast) fast = tail_fast_fits(tg, nullptr, fast_threads);
        }
        std::vector<TailArgs>& lst = fast ? tails_fast : tails;
        std::vector<int>& tot = fast ? tail_fast_totals : tail_totals;
        if (lst.empty() || lst.back().ngroups == kTailMaxGroups) {
            TailArgs t{};
            t.vfos = fe->d_vfos; t.arena_iq = fe->d_arena_iq + (size_t)aset * fe->arena_cap; t.arena_demod = fe->d_arena_demod + (size_t)aset * fe->arena_cap;
            lst.push_back(t); tot.push_back(0);
        }
        lst.back().g[lst.back().ngroups++] = tg;
        tot.back() += (int)g.members.size();
    }
    for (int pi = 0; pi < 2; pi++) { int rc = flush_tc(pi); if (rc != SDRPP_OK) return rc; }
    if (forked) {
        FE_TRY(fe, L.record(SID_S1B, fe->ev_s1_join));
        FE_TRY(fe, L.wait(st, fe->ev_s1_join));
    }
    if (int rc = tl_mark(st, 2); rc != SDRPP_OK) return rc;
    if (prof) FE_TRY(fe, L.record(st, fe->pev[3]));
    else {
        if (fft_order == 0) FE_TRY(fe, L.wait(st, fe->ev_fftk));        // join of the spectrum branch: last node of the stage-1 graph
        L.cur_graph = GRAPH_NONE;
        FE_TRY(fe, L.record(st, fe->ev_s1));
        if (fft_order == 0) {
            // the spectrum kernels were a branch of the main stream's own graph: the ring is ordered by the stream itself; the
            // event only serves a later block that runs in another order
            FE_TRY(fe, L.record(st, fe->ev_fft[par]));
            fe->ev_fft_valid[par] = true;
            // rows to the host behind the graph
            FE_TRY(fe, L.wait(sf, fe->ev_s1));
            for (const PendingCopy& c : fft_copies) FE_TRY(fe, L.memcpy_async(sf, c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost));
            FE_TRY(fe, L.record(sf, fe->ev_rows[aset]));
        }
        FE_TRY(fe, L.wait(stl, fe->ev_s1));
        if (fft_order == 1) {
            // the spectrum of this block runs behind its stage 1, beside its tail
            FE_TRY(fe, L.wait(sf, fe->ev_s1));
            L.cmds.insert(L.cmds.end(), fft_cmds.begin(), fft_cmds.end());
        }
        L.cur_graph = GRAPH_TAIL;
    }
    if (int rc = tl_mark(stl, 5); rc != SDRPP_OK) return rc;
    for (int pass = 0; pass < 2; pass++) {
        std::vector<TailArgs>& lst = pass ? tails : tails_fast;
        std::vector<int>& tot = pass ? tail_totals : tail_fast_totals;
        for (size_t i = 0; i < lst.size(); i++) {
            bool wide = false;
            for (int k = 0; k < lst[i].ngroups; k++) wide = wide || lst[i].g[k].s_begin == 1;
            if (tot[i] <= 0) continue;
            const TailArgs* d_args = L.push(lst[i]);
            if (!d_args) return fail(SDRPP_ERR_STATE, "block descriptor overflow");
            if (wide) FE_TRY(fe, launch_tail_stage0_wide(L, stl, lst[i], d_args, tot[i]));
            if (i == 0) { if (int rc = tl_mark(stl, 6); rc != SDRPP_OK) return rc; }
            FE_TRY(fe, pass ? launch_tail(L, stl, lst[i], d_args, tot[i]) : launch_tail_fast(L, stl, lst[i], d_args, tot[i], fast_threads));
        }
    }
    if (int rc = tl_mark(stl, 7); rc != SDRPP_OK) return rc;
    // the post-detector pass below walks every group, whichever tail kernel it took
    tails.insert(tails.end(), tails_fast.begin(), tails_fast.end());
    tail_totals.insert(tail_totals.end(), tail_fast_totals.begin(), tail_fast_totals.end());
    if (fe->post_active > 0) {
        // post-detector stages of the VFOs that have them, on the outputs the tail just wrote
        for (size_t i = 0; i < tails.size(); i++) {
            PostArgs pa{};
            pa.ngroups = tails[i].ngroups;
            bool any = false;
            for (int k = 0; k < pa.ngroups; k++) {
                pa.g[k].first_vfo = tails[i].g[k].first_vfo; pa.g[k].nvfo = tails[i].g[k].nvfo; pa.g[k].n = tails[i].g[k].n_final;
                any = any || pa.g[k].n > 0;
            }
            pa.post = fe->d_post; pa.arena_iq = fe->d_arena_iq + (size_t)aset * fe->arena_cap; pa.arena_demod = fe->d_arena_demod + (size_t)aset * fe->arena_cap;
            pa.arena_audio = fe->d_arena_audio + (size_t)aset * fe->arena_cap;
            pa.arena_audio_r = fe->d_arena_audio_r + (size_t)aset * fe->arena_cap;
            if (any && tail_totals[i] > 0) FE_TRY(fe, launch_post(L, stl, pa, tail_totals[i]));
        }
    }
    // RDS side outputs of the WFM decoders (behind the tail, which wrote the discriminator rows they start from)
    rs.rds_counts.clear();
    if (!fe->rds_ids.empty()) {
        rs.rds_counts.assign(fe->vfos.size(), -1);
        rs.rds_offs.assign(fe->vfos.size(), 0);
        for (size_t k0 = 0; k0 < fe->rds_ids.size(); k0 += kRdsPerLaunch) {
            RdsArgs ra{};
            ra.tab = fe->d_rds_tab; ra.arena_demod = fe->d_arena_demod + (size_t)aset * fe->arena_cap;
            ra.arena_rds = fe->d_arena_rds + (size_t)aset * fe->rds_arena_cap;
            ra.first = (int)k0; ra.count = (int)std::min<size_t>(kRdsPerLaunch, fe->rds_ids.size() - k0);
            bool any = false;
            for (int k = 0; k < ra.count; k++) {
                const int id = fe->rds_ids[k0 + (size_t)k];
                Vfo& v = fe->vfos[(size_t)id];
                RdsBlk& b = ra.blk[k];
                b.n = fe->groups[(size_t)v.group].last_n_final;
                b.phase0 = v.rds_n * v.rds->dphi;
                v.rds_n += (uint64_t)b.n;
                int cur = b.n;
                for (size_t s2 = 0; s2 < v.rds->stages.size(); s2++) {
                    b.off[s2] = v.rds_off[s2];
                    advance_decim(v.rds_off[s2], v.rds->stages[s2].decimation, cur, &b.nout[s2]);
                    cur = b.nout[s2];
                }
                b.pphase = v.rds_pphase; b.poff = v.rds_poff; b.npoly = 0;
                if (v.rds->tpp > 0) {
                    // polyphase_resampler.h:75-93 in closed form
                    const long long interp = v.rds->rp.interp, decim = v.rds->rp.decim;
                    const long long c = (long long)cur - b.poff;
                    long long cnt = 0;
                    if (c > 0) cnt = (c * interp - b.pphase + decim - 1) / decim;
                    const long long P = (long long)b.pphase + cnt * decim;
                    v.rds_poff = (int)((long long)b.poff + P / interp - cur);
                    v.rds_pphase = (int)(P % interp);
                    b.npoly = (int)cnt;
                    cur = (int)cnt;
                }
                b.nfinal = std::min(cur, v.rds_out_cap);
                rs.rds_counts[(size_t)id] = b.nfinal;
                rs.rds_offs[(size_t)id] = v.rds_out_off;
                any = any || b.n > 0;
            }
            if (any) FE_TRY(fe, launch_rds(L, stl, ra));
        }
    }
    L.cur_graph = GRAPH_NONE;
    if (prof) { FE_TRY(fe, L.record(st, fe->pev[4])); fe->pev_valid = true; }

    // ---- results to pinned host memory --------------------------------------------------------------
    rs.counts.assign(fe->vfos.size(), 0);
    rs.offs.assign(fe->vfos.size(), 0);
    rs.demods.assign(fe->vfos.size(), 0);
    rs.has_audio.assign(fe->vfos.size(), 0);
    rs.stereo.assign(fe->vfos.size(), 0);
    for (const Group& g : fe->groups)
        for (int id : g.members) {
            const Vfo& v = fe->vfos[(size_t)id];
            rs.counts[(size_t)id] = g.last_n_final; rs.offs[(size_t)id] = v.out_off; rs.demods[(size_t)id] = v.demod;
            rs.has_audio[(size_t)id] = (v.post.enabled && v.post_state) ? 1 : 0;
            rs.stereo[(size_t)id] = (v.post.enabled && v.post_state && v.post.wfm && v.demod == SDRPP_DEMOD_QUADRATURE) ? 1 : 0;
        }
    // The copies run on their own stream behind the tail, so the tail of the next block does not queue up behind
    // them; the pinned result set and the arena of this parity are free again once the caller has waited for
    // block i (it must, before submitting block i+2).
    const int sd = prof ? st : SID_D2H;
    if (!prof) {
        FE_TRY(fe, L.record(stl, fe->ev_tail[par]));
        fe->ev_tail_valid[par] = true;
        FE_TRY(fe, L.wait(sd, fe->ev_tail[par]));
    } else {
        fe->ev_tail_valid[0] = fe->ev_tail_valid[1] = false;  // everything ran in order on st
    }
    if (fe->readback && fe->arena_used > 0) {
        const size_t ao = (size_t)aset * fe->arena_cap;
        FE_TRY(fe, L.memcpy_async(sd, rs.iq, fe->d_arena_iq + ao, fe->arena_used * sizeof(float2), cudaMemcpyDeviceToHost));
        FE_TRY(fe, L.memcpy_async(sd, rs.demod, fe->d_arena_demod + ao, fe->arena_used * sizeof(float), cudaMemcpyDeviceToHost));
        if (fe->post_active > 0)
            FE_TRY(fe, L.memcpy_async(sd, rs.audio, fe->d_arena_audio + ao, fe->arena_used * sizeof(float), cudaMemcpyDeviceToHost));
        if (fe->post_stereo > 0)
            FE_TRY(fe, L.memcpy_async(sd, rs.audio_r, fe->d_arena_audio_r + ao, fe->arena_used * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (fe->readback && !fe->rds_ids.empty() && fe->rds_arena_used > 0)
        FE_TRY(fe, L.memcpy_async(sd, rs.rds, fe->d_arena_rds + (size_t)aset * fe->rds_arena_cap, fe->rds_arena_used * sizeof(float2), cudaMemcpyDeviceToHost));
    if (!prof) FE_TRY(fe, L.wait(sd, fe->ev_rows[aset])); // the block is done when its rows are on the host too
    if (int rc = tl_mark(sd, 9); rc != SDRPP_OK) return rc;
    FE_TRY(fe, L.record(sd, rs.done));
    fe->launches += L.kernels - kernels0;
    {
        const int rc = execute_block(fe);
        if (rc != SDRPP_OK) return rc;
    }
    rs.pending = true;
    fe->blk++;
    return SDRPP_OK;
}

static int wait_set(sdrpp_cuda_frontend* fe, int idx) {
    ResultSet& rs = fe->rs[idx];
    if (rs.pending) {
        FE_TRY(fe, cudaEventSynchronize(rs.done));
        rs.pending = false;
    }
    return SDRPP_OK;
}

static int submit_common(sdrpp_cuda_frontend* fe, int fmt, const void* in, int count, bool device_src, float scale = 1.0f) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    std::lock_guard<std::mutex> api(fe->api_mtx);
    if (fmt < 0 || (fmt >= SDRPP_FMT_COUNT && fmt != kFmtPcmI8 && fmt != kFmtPcmI16)) return fail(SDRPP_ERR_ARG, "unknown sample format");
    if (count <= 0 || count > fe->cfg.max_block) return fail(SDRPP_ERR_ARG, "count must be in 1..max_block");
    {
        const bool needs_data = !(fe->comm && fe->comm->nranks > 1 && fe->comm->rank != fe->comm_root);
        if (needs_data && !in) return fail(SDRPP_ERR_ARG, "null sample buffer");
        if (!needs_data && in) return fail(SDRPP_ERR_STATE, "this rank receives its blocks from the communicator's root: use sdrpp_cuda_frontend_submit_shared");
    }
    const int slot = (int)(fe->seq % kSets);
    ResultSet& rs = fe->rs[slot];
    rc = wait_set(fe, slot); // the block that used this result set kSets submits ago must be done
    if (rc != SDRPP_OK) return rc;
    const void* d_in = in;
    const bool shared = fe->comm != nullptr && fe->comm->nranks > 1;
    const bool is_root = !shared || fe->comm->rank == fe->comm_root;
    const size_t bytes = (size_t)count * fmt_bytes_per_sample(fmt);
    if (is_root && !device_src) {
        const void* src = in;
        cudaPointerAttributes attr{};
        bool pinned = (cudaPointerGetAttributes(&attr, in) == cudaSuccess) && (attr.type == cudaMemoryTypeHost);
        cudaGetLastError();
        if (fe->consumed_valid[slot]) FE_TRY(fe, cudaStreamWaitEvent(fe->st_copy, fe->ev_consumed[slot], 0));
        if (!pinned) {
            // pageable source: the staging buffer is reusable once its previous H2D copy has finished
            FE_TRY(fe, cudaEventSynchronize(fe->ev_h2d[slot]));
            memcpy(fe->h_stage[slot], in, bytes);
            src = fe->h_stage[slot];
        }
        FE_TRY(fe, cudaMemcpyAsync(fe->d_raw[slot], src, bytes, cudaMemcpyHostToDevice, fe->st_copy));
        FE_TRY(fe, cudaEventRecord(fe->ev_h2d[slot], fe->st_copy));
        if (fe->timeline) FE_TRY(fe, cudaEventRecord(fe->tl_ev[fe->blk % sdrpp_cuda_frontend::kTlBlocks][8], fe->st_copy));
        if (!shared) FE_TRY(fe, cudaStreamWaitEvent(fe->st, fe->ev_h2d[slot], 0));
        d_in = fe->d_raw[slot];
        fe->last_in_slot = slot;
    } else {
        fe->last_in_slot = -1;
    }
    if (shared) {
        // One broadcast of the raw (packed) block per submit, on its own stream so that the broadcast of block i+1 runs
        // beside the kernels of block i. Root: in place out of the staged / caller's device buffer; others: into d_raw.
        cudaStream_t sb = fe->st_bcast;
        void* recv;
        if (is_root) {
            if (!device_src) FE_TRY(fe, cudaStreamWaitEvent(sb, fe->ev_h2d[slot], 0));
            recv = const_cast<void*>(d_in);
        } else {
            if (fe->consumed_valid[slot]) FE_TRY(fe, cudaStreamWaitEvent(sb, fe->ev_consumed[slot], 0));
            recv = fe->d_raw[slot];
            d_in = recv;
        }
        FE_TRY(fe, comm_broadcast(fe->comm, d_in, recv, bytes, fe->comm_root, sb));
        FE_TRY(fe, cudaEventRecord(fe->ev_bcast[slot], sb));
        FE_TRY(fe, cudaStreamWaitEvent(fe->st, fe->ev_bcast[slot], 0));
    }
    fe->blk_device_src = device_src || !is_root;
    rc = process_block(fe, fmt, d_in, count, rs, scale);
    if (rc != SDRPP_OK) return rc;
    if (!device_src || !is_root) {
        // the raw buffer may be overwritten once this block's kernels are done
        FE_TRY(fe, cudaEventRecord(fe->ev_consumed[slot], fe->st));
        fe->consumed_valid[slot] = true;
    }
    fe->seq++;
    return SDRPP_OK;
}

// ---- one-shot context ----------------------------------------------------------------------------
struct OneShot {
    std::mutex mtx;
    cudaStream_t st = nullptr;
    void* d_a = nullptr; size_t cap_a = 0;
    void* d_b = nullptr; size_t cap_b = 0;
    void* d_c = nullptr; size_t cap_c = 0;
    void* d_d = nullptr; size_t cap_d = 0;
    void* d_w = nullptr; size_t cap_w = 0;   // spectrum_device: cached window (natural order)
    std::vector<float> win_host; int win_n = 0;
    int device = -1;
    // immediate-mode launcher of the one-shot operations: pageable host copy of the argument records, device twin used as a ring
    Launcher L;
    std::vector<unsigned char> h_desc;
    unsigned char* d_desc = nullptr;
};
static OneShot g_os;
constexpr size_t kOsDescBytes = 64 * 1024;
static thread_local int g_device = 0;

static cudaError_t os_reserve(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}
static int os_prepare() {
    SDRPP_CUDA_TRY(cudaSetDevice(g_device));
    if (g_os.device != g_device) {
        if (g_os.st) { cudaStreamDestroy(g_os.st); g_os.st = nullptr; }
        g_os.d_a = g_os.d_b = g_os.d_c = g_os.d_d = g_os.d_w = nullptr;
        g_os.cap_a = g_os.cap_b = g_os.cap_c = g_os.cap_d = g_os.cap_w = 0;
        g_os.win_host.clear(); g_os.win_n = 0;
        g_os.d_desc = nullptr;
        g_os.device = g_device;
    }
    if (!g_os.st) SDRPP_CUDA_TRY(cudaStreamCreateWithFlags(&g_os.st, cudaStreamNonBlocking));
    if (!g_os.d_desc) {
        SDRPP_CUDA_TRY(cudaMalloc((void**)&g_os.d_desc, kOsDescBytes));
        g_os.h_desc.assign(kOsDescBytes, 0);
        g_os.L.begin_block(g_os.h_desc.data(), g_os.d_desc, kOsDescBytes);
    }
    return SDRPP_OK;
}
// launcher that runs every command at once on `st`; argument records go round the device ring (a record is overwritten
// thousands of calls later, long after the kernel that read it)
static Launcher& os_launcher(cudaStream_t st) {
    Launcher& L = g_os.L;
    L.immediate(st);
    if (L.desc_used + 8192 > L.desc_cap) L.desc_used = L.desc_flushed = 0;
    return L;
}

} // namespace sdrpp

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* sdrpp_cuda_version(void) { return "sdrpp-b200 0.1 (sm_100a)"; }
const char* sdrpp_cuda_last_error(void) { return g_last_error.c_str(); }

int sdrpp_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sdrpp_cuda_init(int device) {
    int n = sdrpp_cuda_device_count();
    if (n <= 0) return fail(SDRPP_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(SDRPP_ERR_ARG, "device index out of range");
    SDRPP_CUDA_TRY(cudaSetDevice(device));
    g_device = device;
    return SDRPP_OK;
}

void* sdrpp_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_last_error(std::string("cudaMallocHost: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void sdrpp_cuda_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- design maths ----------------------------------------------------------------------------
int sdrpp_cuda_design_window(int type, float* buf, int size, int centered) {
    if (design_window(type, buf, size, centered != 0) != 0) return fail(SDRPP_ERR_ARG, "bad window type or size");
    return SDRPP_OK;
}
int sdrpp_cuda_design_lowpass(double cutoff, double transWidth, double sampleRate, float* out, int cap) {
    if (!(transWidth > 0) || !(sampleRate > 0)) return fail(SDRPP_ERR_ARG, "bad filter spec");
    const int n = lowpass_tap_count(transWidth, sampleRate);
    if (out && cap > 0) {
        std::vector<float> t = design_lowpass(cutoff, transWidth, sampleRate);
        memcpy(out, t.data(), sizeof(float) * (size_t)std::min(n, cap));
    }
    return n;
}
int sdrpp_cuda_design_bandpass_complex(double bandStart, double bandStop, double transWidth, double sampleRate, int oddTapCount, float* out, int cap) {
    if (!(transWidth > 0) || !(sampleRate > 0) || !(bandStop > bandStart)) return fail(SDRPP_ERR_ARG, "bad filter spec");
    const std::vector<float> t = design_bandpass_complex(bandStart, bandStop, transWidth, sampleRate, oddTapCount != 0);
    const int n = (int)t.size() / 2;
    if (out && cap > 0) memcpy(out, t.data(), sizeof(float) * 2 * (size_t)std::min(n, cap));
    return n;
}
int sdrpp_cuda_design_resampler(double inSR, double outSR, int* info, float* taps, int cap) {
    if (!(inSR > 0) || !(outSR > 0) || !info) return fail(SDRPP_ERR_ARG, "bad rates");
    ResamplerPlan p = design_resampler(inSR, outSR);
    info[0] = p.mode; info[1] = p.predec; info[2] = p.interp; info[3] = p.decim;
    info[4] = (int)p.taps.size(); info[5] = p.tpp;
    if (taps && cap > 0) memcpy(taps, p.taps.data(), sizeof(float) * std::min<size_t>(p.taps.size(), (size_t)cap));
    return SDRPP_OK;
}
int sdrpp_cuda_design_decim_plan(int ratio, int* decimation, int* tapcount, const float** taps) {
    std::vector<DecimStage> s = decim_plan(ratio);
    for (size_t i = 0; i < s.size(); i++) {
        if (decimation) decimation[i] = s[i].decimation;
        if (tapcount) tapcount[i] = s[i].ntaps;
        if (taps) taps[i] = s[i].taps;
    }
    return (int)s.size();
}
void sdrpp_cuda_design_reshape(double sampleRate, int fftSize, double fftRate, int* skip, int* nz) {
    reshape_params(sampleRate, fftSize, fftRate, skip, nz);
}

// ---- one-shot operations -----------------------------------------------------------------------
int sdrpp_cuda_convert(int fmt, const void* in, int nsamples, sdrpp_cf32* out) {
    if (fmt < 0 || fmt >= SDRPP_FMT_COUNT || nsamples < 0 || (nsamples > 0 && (!in || !out))) return fail(SDRPP_ERR_ARG, "bad argument");
    if (nsamples == 0) return SDRPP_OK;
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    const size_t inb = (size_t)nsamples * fmt_bytes_per_sample(fmt), outb = (size_t)nsamples * 8;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, inb));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, outb));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, in, inb, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_ingest(os_launcher(g_os.st), SID_MAIN, fmt, g_os.d_a, nsamples, RingRef{ (float2*)g_os.d_b, 0xFFFFFFFFu }, 0, false));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(out, g_os.d_b, outb, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    return SDRPP_OK;
}

// ---- SDR++ server wire packets (dsp/compression/sample_stream_{de,}compressor.h) -------------------
// Packet: u16 compression type (0) | u16 PCMType | f32 scaler | payload (sample_stream_compressor.h:27-34).
struct PcmHeader { int fmt; float divisor; int nsamples; };
// Header arithmetic of SampleStreamDecompressor::process (sample_stream_decompressor.h:14-33). fmt < 0: the
// reference returns 0 samples for an unknown sample type.
static PcmHeader pcm_parse(const void* packet, int nbytes) {
    uint16_t type; float scaler;
    memcpy(&type, (const uint8_t*)packet + 2, 2);
    memcpy(&scaler, (const uint8_t*)packet + 4, 4);
    const int payload = nbytes - 8;
    switch (type) {
    case SDRPP_PCM_F32: return { SDRPP_FMT_CF32, 1.0f, payload / 8 };
    case SDRPP_PCM_I16: return { kFmtPcmI16, 32768.0f / scaler, payload / 4 };
    case SDRPP_PCM_I8:  return { kFmtPcmI8, 128.0f / scaler, payload / 2 };
    }
    return { -1, 1.0f, 0 };
}

int sdrpp_cuda_pcm_decompress(const void* packet, int nbytes, sdrpp_cf32* out) {
    if (!packet || nbytes < 8 || !out) return fail(SDRPP_ERR_ARG, "bad argument");
    const PcmHeader h = pcm_parse(packet, nbytes);
    if (h.fmt < 0 || h.nsamples == 0) return 0;
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    const size_t inb = (size_t)h.nsamples * fmt_bytes_per_sample(h.fmt), outb = (size_t)h.nsamples * 8;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, inb));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, outb));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, (const uint8_t*)packet + 8, inb, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_ingest(os_launcher(g_os.st), SID_MAIN, h.fmt, g_os.d_a, h.nsamples, RingRef{ (float2*)g_os.d_b, 0xFFFFFFFFu }, 0, false, h.divisor));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(out, g_os.d_b, outb, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    return h.nsamples;
}

int sdrpp_cuda_pcm_compress(int pcm_type, const sdrpp_cf32* in, int count, void* packet) {
    if (!in || !packet || count < 0) return fail(SDRPP_ERR_ARG, "bad argument");
    if (pcm_type != SDRPP_PCM_I8 && pcm_type != SDRPP_PCM_I16 && pcm_type != SDRPP_PCM_F32) return fail(SDRPP_ERR_ARG, "unknown PCM type");
    uint8_t* pk = (uint8_t*)packet;
    const uint16_t comp = 0, type = (uint16_t)pcm_type;
    memcpy(pk, &comp, 2);
    memcpy(pk + 2, &type, 2);
    float scaler = 0.0f;
    if (pcm_type == SDRPP_PCM_F32 || count == 0) {
        // float32 needs no device work (sample_stream_compressor.h:37-41)
        memcpy(pk + 4, &scaler, 4);
        if (pcm_type != SDRPP_PCM_F32) return 8;
        memcpy(pk + 8, in, (size_t)count * 8);
        return 8 + count * 8;
    }
    const int bits = pcm_type == SDRPP_PCM_I8 ? 8 : 16;
    const size_t inb = (size_t)count * 8, outb = (size_t)count * 2 * (bits / 8);
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, inb));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, outb));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_c, &g_os.cap_c, 16));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, in, inb, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_pcm_compress(bits, (const float*)g_os.d_a, 2 * count, (unsigned int*)g_os.d_c, g_os.d_b,
                                       (float*)g_os.d_c + 1, g_os.st));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(pk + 8, g_os.d_b, outb, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(&scaler, (float*)g_os.d_c + 1, 4, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    memcpy(pk + 4, &scaler, 4);
    return 8 + (int)outb;
}

int sdrpp_cuda_spectrum(int N, int nz, int fmt, const void* frame, const float* window, float* row, sdrpp_cf32* X) {
    int N1, N2;
    if (spectrum_split(N, &N1, &N2) < 0) return fail(SDRPP_ERR_ARG, "N must be a power of two in 64..4194304");
    if (nz < 1 || nz > N || !frame || !window || fmt < 0 || fmt >= SDRPP_FMT_COUNT) return fail(SDRPP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    const size_t inb = (size_t)nz * fmt_bytes_per_sample(fmt);
    // d_a: raw frame; d_b: [cf32 frame nz | window nz floats]; d_c: four-step intermediate; d_d: [row N floats | X N cf32]
    const size_t woff = ((size_t)nz * 8 + 63) & ~(size_t)63;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, inb));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, woff + (size_t)nz * 4));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_c, &g_os.cap_c, (size_t)N * 8));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_d, &g_os.cap_d, (size_t)N * 12));
    float2* d_frame = (float2*)g_os.d_b;
    float* d_win = (float*)((char*)g_os.d_b + woff);
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, frame, inb, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(d_win, window, (size_t)nz * 4, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_ingest(os_launcher(g_os.st), SID_MAIN, fmt, g_os.d_a, nz, RingRef{ d_frame, 0xFFFFFFFFu }, 0, false));
    SpectrumArgs a{};
    a.in = d_frame; a.ring_mask = 0xFFFFFFFFu; a.start = 0; a.frame_stride = 0; a.nz = nz; a.window = d_win;
    a.inter = (float2*)g_os.d_c;
    a.rows = row ? (float*)g_os.d_d : nullptr;
    a.X = X ? (float2*)((char*)g_os.d_d + (size_t)N * 4) : nullptr;
    a.frames = 1;
    SDRPP_CUDA_TRY(launch_spectrum(os_launcher(g_os.st), SID_MAIN, N, a, nullptr));
    if (row) SDRPP_CUDA_TRY(cudaMemcpyAsync(row, a.rows, (size_t)N * 4, cudaMemcpyDeviceToHost, g_os.st));
    if (X) SDRPP_CUDA_TRY(cudaMemcpyAsync(X, a.X, (size_t)N * 8, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    return SDRPP_OK;
}

int sdrpp_cuda_spectrum_device(int N, int nz, int frames, long long frame_stride, const sdrpp_cf32* dev_in, const float* window,
                               float* dev_rows, void* stream) {
    int N1, N2;
    if (spectrum_split(N, &N1, &N2) < 0) return fail(SDRPP_ERR_ARG, "N must be a power of two in 64..4194304");
    if (nz < 1 || nz > N || frames < 1 || frame_stride < 0 || !dev_in || !dev_rows) return fail(SDRPP_ERR_ARG, "bad argument");
    if ((long long)(frames - 1) * frame_stride + nz > 0x7FFFFFFFLL) return fail(SDRPP_ERR_ARG, "input span exceeds 2^31 samples");
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : g_os.st;
    // window table (re-uploaded only when it changes) and an L2-sized four-step intermediate
    if (!window && (g_os.win_n != N || g_os.win_host.size() != (size_t)nz)) return fail(SDRPP_ERR_STATE, "no cached window of this size: pass the window table");
    if (window && (g_os.win_n != N || g_os.win_host.size() != (size_t)nz || memcmp(g_os.win_host.data(), window, (size_t)nz * 4) != 0)) {
        SDRPP_CUDA_TRY(cudaStreamSynchronize(st));
        SDRPP_CUDA_TRY(os_reserve(&g_os.d_w, &g_os.cap_w, (size_t)nz * 4));
        SDRPP_CUDA_TRY(upload_sync(g_os.d_w, window, (size_t)nz * 4));
        g_os.win_host.assign(window, window + nz);
        g_os.win_n = N;
    }
    // frames per launch: enough CTAs for several waves (the kernels' load and compute phases only overlap once
    // CTAs are out of step); the intermediate need not stay in L2 for that to pay (profiles/README.md)
    const size_t inter_cap = getenv("SDRPP_FFT_INTER_MB") ? (size_t)atoi(getenv("SDRPP_FFT_INTER_MB")) << 20 : ((size_t)256 << 20);
    const int group = N1 > 1 ? std::max(1, std::min(frames, (int)(inter_cap / ((size_t)N * 8)))) : frames;
    if (N1 > 1 && g_os.cap_c < (size_t)group * N * 8) {
        SDRPP_CUDA_TRY(cudaStreamSynchronize(st));
        SDRPP_CUDA_TRY(os_reserve(&g_os.d_c, &g_os.cap_c, (size_t)group * N * 8));
    }
    for (int f0 = 0; f0 < frames; f0 += group) {
        SpectrumArgs a{};
        a.in = dev_in + (size_t)f0 * (size_t)frame_stride; a.ring_mask = 0xFFFFFFFFu; a.start = 0; a.frame_stride = (uint32_t)frame_stride;
        a.nz = nz; a.window = (const float*)g_os.d_w; a.inter = (float2*)g_os.d_c;
        a.rows = dev_rows + (size_t)f0 * N; a.X = nullptr; a.frames = std::min(group, frames - f0);
        SDRPP_CUDA_TRY(launch_spectrum(os_launcher(st), SID_MAIN, N, a, nullptr));
    }
    if (!stream) SDRPP_CUDA_TRY(cudaStreamSynchronize(st));
    return SDRPP_OK;
}

int sdrpp_cuda_fft_zoom(int N, const float* row, double viewOffset, double viewBandwidth, double wholeBandwidth, int outSize,
                        float* out, int* idx_out) {
    if (N < 1 || !row || outSize < 1 || !(viewBandwidth > 0) || !(wholeBandwidth > 0) || !out) return fail(SDRPP_ERR_ARG, "bad argument");
    std::vector<int> idx;
    const bool ranged = zoom_indices(viewOffset, viewBandwidth, wholeBandwidth, N, outSize, &idx);
    if (idx_out) memcpy(idx_out, idx.data(), idx.size() * sizeof(int));
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, (size_t)N * 4));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, idx.size() * 4));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_c, &g_os.cap_c, (size_t)outSize * 4));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, row, (size_t)N * 4, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_b, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_fft_zoom(os_launcher(g_os.st), SID_MAIN, (const float*)g_os.d_a, N, 1, (const int*)g_os.d_b, outSize, ranged, (float*)g_os.d_c));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(out, g_os.d_c, (size_t)outSize * 4, cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    return SDRPP_OK;
}

int sdrpp_cuda_signal_info(int N, const float* row, int nvfo, const double* centerOffset, const double* bandwidth, double wholeBandwidth,
                           float* strength, float* snr) {
    if (N < 2 || !row || nvfo < 1 || !centerOffset || !bandwidth || !(wholeBandwidth > 0)) return fail(SDRPP_ERR_ARG, "bad argument");
    std::vector<int4> bins((size_t)nvfo);
    for (int v = 0; v < nvfo; v++) {
        int b[4];
        signal_info_bins(centerOffset[v], bandwidth[v], wholeBandwidth, N, b);
        bins[(size_t)v] = make_int4(b[0], b[1], b[2], b[3]);
    }
    std::lock_guard<std::mutex> lck(g_os.mtx);
    int rc = os_prepare();
    if (rc != SDRPP_OK) return rc;
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_a, &g_os.cap_a, (size_t)N * 4));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_b, &g_os.cap_b, (size_t)nvfo * sizeof(int4)));
    SDRPP_CUDA_TRY(os_reserve(&g_os.d_c, &g_os.cap_c, (size_t)nvfo * sizeof(float2)));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_a, row, (size_t)N * 4, cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(cudaMemcpyAsync(g_os.d_b, bins.data(), (size_t)nvfo * sizeof(int4), cudaMemcpyHostToDevice, g_os.st));
    SDRPP_CUDA_TRY(launch_signal_info(os_launcher(g_os.st), SID_MAIN, (const float*)g_os.d_a, N, 1, (const int4*)g_os.d_b, nvfo, (float2*)g_os.d_c));
    std::vector<float2> out((size_t)nvfo);
    SDRPP_CUDA_TRY(cudaMemcpyAsync(out.data(), g_os.d_c, (size_t)nvfo * sizeof(float2), cudaMemcpyDeviceToHost, g_os.st));
    SDRPP_CUDA_TRY(cudaStreamSynchronize(g_os.st));
    for (int v = 0; v < nvfo; v++) { if (strength) strength[v] = out[(size_t)v].x; if (snr) snr[v] = out[(size_t)v].y; }
    return SDRPP_OK;
}

// ---- front end ---------------------------------------------------------------------------------
sdrpp_cuda_frontend* sdrpp_cuda_frontend_create(const sdrpp_cuda_frontend_cfg* cfg) {
    if (!cfg) { set_last_error("null cfg"); return nullptr; }
    if (!(cfg->sample_rate > 0)) { set_last_error("sample_rate must be positive"); return nullptr; }
    if (cfg->decim_ratio != 0 && cfg->decim_ratio != 1 && (!is_pow2(cfg->decim_ratio) || cfg->decim_ratio > 8192)) {
        set_last_error("decim_ratio must be 1 or a power of two <= 8192"); return nullptr;
    }
    if (sdrpp_cuda_device_count() <= 0) { set_last_error("no CUDA device: this library has no CPU fallback"); return nullptr; }
    auto* fe = new sdrpp_cuda_frontend();
    fe->device = g_device;
    fe->cfg = *cfg;
    if (fe->cfg.decim_ratio < 1) fe->cfg.decim_ratio = 1;
    if (fe->cfg.max_block <= 0) fe->cfg.max_block = 1000000;
    auto bail = [&](const std::string& why) -> sdrpp_cuda_frontend* {
        if (!why.empty()) set_last_error(why);
        sdrpp_cuda_frontend_destroy(fe);
        return nullptr;
    };
    if (cudaSetDevice(fe->device) != cudaSuccess) return bail("cudaSetDevice failed");
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, fe->device) == cudaSuccess && sms > 0) fe->num_sms = sms;
        // SMs left free for other work on the device while the persistent stage-1 kernel runs (a multi-GPU deployment
        // sets this so that the NCCL broadcast of the next block is not queued behind it)
        const char* rsv = getenv("SDRPP_RESERVE_SMS");
        if (rsv) { const int r = atoi(rsv); if (r > 0 && r < fe->num_sms) fe->num_sms -= r; }
        const char* tm = getenv("SDRPP_TAIL_MODE");
        fe->tail_mode = (tm && (!strcmp(tm, "general") || !strcmp(tm, "1"))) ? 1 : (tm && (!strcmp(tm, "fast") || !strcmp(tm, "2"))) ? 2
                      : (tm && (!strcmp(tm, "narrow") || !strcmp(tm, "3"))) ? 3 : 0;
        const char* m = getenv("SDRPP_S1_MODE");
        fe->s1_mode = (m && (!strcmp(m, "fp32") || !strcmp(m, "1"))) ? 1 : 0;
    }
    // ring: history for the longest filter + a whole spectrum frame + three blocks (the spectrum of block i-1 may still
    // read while block i is written, see process_block), rounded up to a power of two
    {
        long long need = (long long)fe->cfg.max_block * 3 + std::max(fe->cfg.fft_size, 0) * 2LL + 8192;
        int lg = 16;
        while ((1LL << lg) < need) lg++;
        if (fe->cfg.ring_log2 > 0) lg = fe->cfg.ring_log2;
        if (lg < 12 || lg > 30) return bail("ring_log2 out of range");
        fe->ring_log2 = lg;
        fe->ring_mask = (uint32_t)((1u << lg) - 1u);
    }
    if (cudaStreamCreateWithFlags(&fe->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_fft, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_tail, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_d2h, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_s1b, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_desc, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&fe->st_bcast, cudaStreamNonBlocking) != cudaSuccess) return bail("stream creation failed");
    if (cudaEventCreateWithFlags(&fe->ev_ingest, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_s1, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_fft[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_fft[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_s1_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_s1_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_tail[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_tail[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_desc, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_fftk, cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
    for (cudaEvent_t& e : fe->ev_rows) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
    fe->L.streams[SID_MAIN] = fe->st; fe->L.streams[SID_FFT] = fe->st_fft; fe->L.streams[SID_TAIL] = fe->st_tail;
    fe->L.streams[SID_S1B] = fe->st_s1b; fe->L.streams[SID_D2H] = fe->st_d2h;
    { const char* g = getenv("SDRPP_GRAPHS"); fe->graphs_on = !(g && g[0] == '0'); }
    if (const char* g = getenv("SDRPP_TIMELINE"); g && g[0] == '1') {
        fe->timeline = true; fe->graphs_on = false;
        for (auto& row : fe->tl_ev) for (cudaEvent_t& e : row) { if (cudaEventCreate(&e) != cudaSuccess) return bail("event creation failed"); cudaEventRecord(e, fe->st); }
    }
    { const char* g = getenv("SDRPP_FUSE_INGEST"); fe->fuse_ingest = !(g && g[0] == '0'); }
    { const char* g = getenv("SDRPP_FFT_ORDER"); if (g && g[0] >= '0' && g[0] <= '2') fe->fft_order = g[0] - '0'; }
    for (int i = 0; i < kSets; i++) {
        if (cudaMallocHost((void**)&fe->h_desc[i], kDescBytes) != cudaSuccess) return bail("descriptor allocation failed");
        if (cudaMalloc((void**)&fe->d_desc[i], kDescBytes) != cudaSuccess) return bail("descriptor allocation failed");
        memset(fe->h_desc[i], 0, kDescBytes);
    }
    if (dev_alloc(&fe->ring, (size_t)1 << fe->ring_log2) != cudaSuccess) return bail("ring allocation failed");
    fe->raw_cap = (size_t)fe->cfg.max_block * 16; // widest input format: complex f64
    for (int i = 0; i < kSets; i++) {
        if (cudaMallocHost(&fe->h_stage[i], fe->raw_cap) != cudaSuccess) return bail("pinned staging allocation failed");
        if (cudaMalloc(&fe->d_raw[i], fe->raw_cap) != cudaSuccess) return bail("raw buffer allocation failed");
        if (cudaEventCreateWithFlags(&fe->ev_h2d[i], cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
        if (cudaEventCreateWithFlags(&fe->ev_consumed[i], cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
        if (cudaEventCreateWithFlags(&fe->rs[i].done, cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
        if (cudaEventCreateWithFlags(&fe->ev_bcast[i], cudaEventDisableTiming) != cudaSuccess) return bail("event creation failed");
        cudaEventRecord(fe->ev_h2d[i], fe->st_copy);
    }
    for (int i = 0; i < 5; i++) if (cudaEventCreate(&fe->pev[i]) != cudaSuccess) return bail("event creation failed");
    if (configure_preproc(fe) != SDRPP_OK) return bail("");
    if (configure_fft(fe) != SDRPP_OK) return bail("");
    if (cudaDeviceSynchronize() != cudaSuccess) return bail("device sync failed");
    return fe;
}

int sdrpp_cuda_frontend_destroy(sdrpp_cuda_frontend* fe) {
    if (!fe) return SDRPP_OK;
    cudaSetDevice(fe->device);
    if (fe->st) cudaStreamSynchronize(fe->st);
    if (fe->st_copy) cudaStreamSynchronize(fe->st_copy);
    if (fe->st_fft) cudaStreamSynchronize(fe->st_fft);
    if (fe->st_tail) cudaStreamSynchronize(fe->st_tail);
    if (fe->st_d2h) cudaStreamSynchronize(fe->st_d2h);
    if (fe->st_s1b) cudaStreamSynchronize(fe->st_s1b);
    if (fe->st_bcast) cudaStreamSynchronize(fe->st_bcast);
    if (fe->st_desc) cudaStreamSynchronize(fe->st_desc);
    for (Vfo& v : fe->vfos) { if (v.slab) cudaFree(v.slab); cudaFree(v.post_state); cudaFree(v.post_taps); cudaFree(v.post_taps2); cudaFree(v.if_state); cudaFree(v.rds_state); v.rds.reset(); }
    fe->rds_plans.clear(); cudaFree(fe->d_rds_tab); cudaFree(fe->d_arena_rds);
    for (Group& g : fe->groups) { if (g.d_G) cudaFree(g.d_G); if (g.d_B) cudaFree(g.d_B); }
    for (int i = 0; i < 2; i++) { cudaFree(fe->tc_planes[i].hi); cudaFree(fe->tc_planes[i].lo); cudaFree(fe->tc_planes[i].sinv); }
    fe->vfos.clear(); fe->groups.clear(); fe->plan_cache.clear();
    for (float* t : fe->fe_taps) cudaFree(t);
    for (float2* b : fe->fe_buf) cudaFree(b);
    cudaFree(fe->dc_in); cudaFree(fe->dc_state); cudaFree(fe->dc_scratch);
    cudaFree(fe->ring); cudaFree(fe->d_window); cudaFree(fe->d_inter); cudaFree(fe->d_rows);
    cudaFree(fe->d_zoom_idx); cudaFree(fe->d_zoom); cudaFree(fe->d_disp); cudaFree(fe->d_sig_bins); cudaFree(fe->d_sig);
    cudaFree(fe->d_vfos); cudaFree(fe->d_post); cudaFree(fe->d_arena_iq); cudaFree(fe->d_arena_demod); cudaFree(fe->d_arena_audio); cudaFree(fe->d_arena_audio_r);
    for (int i = 0; i < kSets; i++) {
        if (fe->h_stage[i]) cudaFreeHost(fe->h_stage[i]);
        if (fe->d_raw[i]) cudaFree(fe->d_raw[i]);
        if (fe->ev_h2d[i]) cudaEventDestroy(fe->ev_h2d[i]);
        if (fe->ev_consumed[i]) cudaEventDestroy(fe->ev_consumed[i]);
        if (fe->rs[i].done) cudaEventDestroy(fe->rs[i].done);
        if (fe->ev_bcast[i]) cudaEventDestroy(fe->ev_bcast[i]);
        if (fe->rs[i].iq) cudaFreeHost(fe->rs[i].iq);
        if (fe->rs[i].demod) cudaFreeHost(fe->rs[i].demod);
        if (fe->rs[i].audio) cudaFreeHost(fe->rs[i].audio);
        if (fe->rs[i].audio_r) cudaFreeHost(fe->rs[i].audio_r);
        if (fe->rs[i].rds) cudaFreeHost(fe->rs[i].rds);
        if (fe->rs[i].rows) cudaFreeHost(fe->rs[i].rows);
        if (fe->rs[i].zoom) cudaFreeHost(fe->rs[i].zoom);
        if (fe->rs[i].hold) cudaFreeHost(fe->rs[i].hold);
        if (fe->rs[i].sig) cudaFreeHost(fe->rs[i].sig);
    }
    for (int i = 0; i < 5; i++) if (fe->pev[i]) cudaEventDestroy(fe->pev[i]);
    for (int k = 0; k < GRAPH_KINDS; k++) { for (auto& e : fe->gcache[k]) if (e.second.exec) cudaGraphExecDestroy(e.second.exec); fe->gcache[k].clear(); }
    for (int i = 0; i < kSets; i++) { if (fe->h_desc[i]) cudaFreeHost(fe->h_desc[i]); if (fe->d_desc[i]) cudaFree(fe->d_desc[i]); }
    if (fe->ev_fftk) cudaEventDestroy(fe->ev_fftk);
    if (fe->ev_desc) cudaEventDestroy(fe->ev_desc);
    for (cudaEvent_t e : fe->ev_rows) if (e) cudaEventDestroy(e);
    if (fe->st_desc) cudaStreamDestroy(fe->st_desc);
    if (fe->st) cudaStreamDestroy(fe->st);
    if (fe->st_copy) cudaStreamDestroy(fe->st_copy);
    if (fe->st_fft) cudaStreamDestroy(fe->st_fft);
    if (fe->st_tail) cudaStreamDestroy(fe->st_tail);
    if (fe->st_d2h) cudaStreamDestroy(fe->st_d2h);
    if (fe->st_s1b) cudaStreamDestroy(fe->st_s1b);
    if (fe->st_bcast) cudaStreamDestroy(fe->st_bcast);
    if (fe->ev_join) cudaEventDestroy(fe->ev_join);
    for (cudaEvent_t e : { fe->ev_ingest, fe->ev_s1, fe->ev_fft[0], fe->ev_fft[1], fe->ev_tail[0], fe->ev_tail[1], fe->ev_s1_fork, fe->ev_s1_join }) if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete fe;
    return SDRPP_OK;
}

static int fe_quiesce(sdrpp_cuda_frontend* fe) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    debug_stale("quiesce entry");
    std::lock_guard<std::mutex> api(fe->api_mtx);
    FE_TRY(fe, cudaStreamSynchronize(fe->st_copy));
    FE_TRY(fe, cudaStreamSynchronize(fe->st));
    FE_TRY(fe, cudaStreamSynchronize(fe->st_fft));
    FE_TRY(fe, cudaStreamSynchronize(fe->st_tail));
    FE_TRY(fe, cudaStreamSynchronize(fe->st_d2h));
    FE_TRY(fe, cudaStreamSynchronize(fe->st_s1b));
    FE_TRY(fe, cudaStreamSynchronize(fe->st_bcast));
    fe->ev_tail_valid[0] = fe->ev_tail_valid[1] = false;
    fe->ev_fft_valid[0] = fe->ev_fft_valid[1] = false;
    // Everything submitted so far is complete. Blocks the caller has not waited for yet stay queued: wait() still hands
    // them back one by one, in order (their result sets remember the row offsets of their own VFO layout), so a control
    // call between submit and wait loses no output -- like the reference, which never drops samples on a retune.
    for (int i = 0; i < kSets; i++) fe->rs[i].pending = false;
    return SDRPP_OK;
}

// A sample-rate / decimation change re-plans every VFO (IQFrontEnd::setSampleRate, iq_frontend.cpp:55-80)
static int replan_all(sdrpp_cuda_frontend* fe) {
    for (size_t id = 0; id < fe->vfos.size(); id++) {
        Vfo& v = fe->vfos[id];
        if (!v.alive) continue;
        remove_from_group(fe, (int)id);
        std::shared_ptr<VfoPlan> plan;
        int rc = get_plan(fe, v.outSR, v.bw, &plan);
        if (rc != SDRPP_OK) return rc;
        if (v.if_state && plan->cap_final > kIfMaxBlock) {
            // the chain cannot follow the new input rate (its staging area holds kIfMaxBlock output samples per block): it is
            // switched off rather than silently bypassed
            cudaFree(v.if_state); v.if_state = nullptr;
            set_last_error("IF chain of a VFO disabled: its block output no longer fits the chain's staging area at the new rate");
        }
        v.plan = plan;
        if (v.slab) cudaFree(v.slab);
        v.slab = nullptr;
        FE_TRY(fe, dev_alloc(&v.slab, plan->slab_elems));
        set_nco(fe, v, v.offset, true);
        join_group(fe, (int)id, fe->abs_pos);
        if (v.post.enabled && (rc = apply_post(fe, v)) != SDRPP_OK) return rc;
    }
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_set_sample_rate(sdrpp_cuda_frontend* fe, double sampleRate) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (!(sampleRate > 0)) return fail(SDRPP_ERR_ARG, "sample_rate must be positive");
    fe->cfg.sample_rate = sampleRate;
    fe->eff_sr = sampleRate / fe->cfg.decim_ratio;
    if ((rc = configure_fft(fe)) != SDRPP_OK) return rc;
    return replan_all(fe);
}
int sdrpp_cuda_frontend_set_decimation(sdrpp_cuda_frontend* fe, int ratio) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (ratio != 1 && (!is_pow2(ratio) || ratio > 8192)) return fail(SDRPP_ERR_ARG, "ratio must be 1 or a power of two <= 8192");
    fe->cfg.decim_ratio = ratio;
    if ((rc = configure_preproc(fe)) != SDRPP_OK) return rc;
    if ((rc = configure_fft(fe)) != SDRPP_OK) return rc;
    return replan_all(fe);
}
int sdrpp_cuda_frontend_set_dc_blocking(sdrpp_cuda_frontend* fe, int enabled) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    fe->cfg.dc_blocking = enabled != 0;
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_set_invert_iq(sdrpp_cuda_frontend* fe, int enabled) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    fe->cfg.invert_iq = enabled != 0;
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_set_fft_size(sdrpp_cuda_frontend* fe, int size) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    const int old = fe->cfg.fft_size;
    fe->cfg.fft_size = size;
    if ((rc = configure_fft(fe)) != SDRPP_OK) { fe->cfg.fft_size = old; configure_fft(fe); return rc; }
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_set_fft_rate(sdrpp_cuda_frontend* fe, double rate) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    const double old = fe->cfg.fft_rate;
    fe->cfg.fft_rate = rate;
    if ((rc = configure_fft(fe)) != SDRPP_OK) { fe->cfg.fft_rate = old; configure_fft(fe); return rc; }
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_set_fft_window(sdrpp_cuda_frontend* fe, int window) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    const int old = fe->cfg.fft_window;
    fe->cfg.fft_window = window;
    if ((rc = configure_fft(fe)) != SDRPP_OK) { fe->cfg.fft_window = old; configure_fft(fe); return rc; }
    return SDRPP_OK;
}
double sdrpp_cuda_frontend_effective_samplerate(sdrpp_cuda_frontend* fe) { return fe ? fe->eff_sr : 0.0; }

int sdrpp_cuda_vfo_create(sdrpp_cuda_frontend* fe, double outSR, double bandwidth, double offset, int demod) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (demod < SDRPP_DEMOD_NONE || demod > SDRPP_DEMOD_DSB) return fail(SDRPP_ERR_ARG, "unknown demod");
    std::shared_ptr<VfoPlan> plan;
    if ((rc = get_plan(fe, outSR, bandwidth, &plan)) != SDRPP_OK) return rc;
    int id = -1;
    for (size_t i = 0; i < fe->vfos.size(); i++) if (!fe->vfos[i].alive) { id = (int)i; break; }
    if (id < 0) { fe->vfos.emplace_back(); id = (int)fe->vfos.size() - 1; }
    Vfo& v = fe->vfos[(size_t)id];
    v = Vfo();
    v.alive = true; v.outSR = outSR; v.bw = bandwidth; v.demod = demod; v.plan = plan;
    FE_TRY(fe, dev_alloc(&v.slab, plan->slab_elems));
    set_nco(fe, v, offset, false);
    join_group(fe, id, fe->abs_pos);
    return id;
}

static int vfo_get(sdrpp_cuda_frontend* fe, int id, Vfo** out) {
    if (!fe || id < 0 || id >= (int)fe->vfos.size() || !fe->vfos[(size_t)id].alive) return fail(SDRPP_ERR_STATE, "unknown VFO id");
    *out = &fe->vfos[(size_t)id];
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_destroy(sdrpp_cuda_frontend* fe, int id) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    remove_from_group(fe, id);
    if (v->slab) cudaFree(v->slab);
    cudaFree(v->post_state); cudaFree(v->post_taps); cudaFree(v->post_taps2); cudaFree(v->if_state); v->if_state = nullptr;
    cudaFree(v->rds_state); v->rds_state = nullptr; v->rds.reset(); v->rds_slot = -1;
    *v = Vfo();
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_set_offset(sdrpp_cuda_frontend* fe, int id, double offset) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    set_nco(fe, *v, offset, true);
    fe->groups[(size_t)v->group].g_dirty = true;
    fe->layout_dirty = true;
    return SDRPP_OK;
}

static int vfo_replan(sdrpp_cuda_frontend* fe, int id, double outSR, double bw, bool new_epoch) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    std::shared_ptr<VfoPlan> plan;
    if ((rc = get_plan(fe, outSR, bw, &plan)) != SDRPP_OK) return rc;
    if (v->if_state && plan->cap_final > kIfMaxBlock)
        return fail(SDRPP_ERR_ARG, "the new output rate makes the VFO's block output exceed the IF chain's staging area: disable the chain first");
    remove_from_group(fe, id);
    v->outSR = outSR; v->bw = bw; v->plan = plan;
    if (v->slab) cudaFree(v->slab);
    v->slab = nullptr;
    FE_TRY(fe, dev_alloc(&v->slab, plan->slab_elems));
    set_nco(fe, *v, v->offset, true);
    (void)new_epoch;
    join_group(fe, id, fe->abs_pos);
    if (v->post.enabled) return apply_post(fe, *v); // the demodulator's filter follows the bandwidth / rate
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_set_bandwidth(sdrpp_cuda_frontend* fe, int id, double bw) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    // RxVFO::setBandwidth (rx_vfo.h:60-70) only swaps the channel filter's taps: xlator, resampler and the
    // filter's history carry on (FIR::setTaps keeps the newest samples and zero-fills when the filter grows,
    // fir.h:31-52). Possible in place when the filter exists before and after; otherwise the VFO restarts.
    std::shared_ptr<VfoPlan> np;
    if ((rc = get_plan(fe, v->outSR, bw, &np)) != SDRPP_OK) return rc;
    const std::shared_ptr<VfoPlan> op = v->plan;
    const bool in_place = op->filter_needed && np->filter_needed && op->tail.size() == np->tail.size() &&
                          op->slab_elems == np->slab_elems && op->final_off == np->final_off && v->group >= 0;
    if (!in_place) return vfo_replan(fe, id, v->outSR, bw, false);
    const TailPlanStage& of = op->tail.back();
    const TailPlanStage& nf = np->tail.back();
    const uint32_t offs[2] = { op->tail.size() == 1 ? op->s1_off[0] : nf.in_off, op->tail.size() == 1 ? op->s1_off[1] : nf.in_off };
    if (nf.T > of.T) {
        for (int r = 0; r < (op->tail.size() == 1 ? 2 : 1); r++)
            FE_TRY(fe, memset_sync(v->slab + offs[r] - (nf.T - 1), 0, sizeof(float2) * (size_t)(nf.T - of.T)));
    }
    const GroupState st = fe->groups[(size_t)v->group].st;
    remove_from_group(fe, id);
    v->bw = bw; v->plan = np;
    set_nco(fe, *v, v->offset, true); // SSB translation follows the bandwidth; NCO phase continues
    join_group_with_state(fe, id, st);
    if (v->post.enabled) return apply_post(fe, *v);
    return SDRPP_OK;
}
int sdrpp_cuda_vfo_set_out_samplerate(sdrpp_cuda_frontend* fe, int id, double outSR, double bw) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    return vfo_replan(fe, id, outSR, bw, true);
}
int sdrpp_cuda_vfo_reset(sdrpp_cuda_frontend* fe, int id) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    remove_from_group(fe, id);
    FE_TRY(fe, memset_sync(v->slab, 0, v->plan->slab_elems * sizeof(float2)));
    // xlator.reset(): phase = 1 at the next sample (frequency_xlator.h:31-34)
    v->phi_ref = 0; v->n_ref = fe->abs_pos;
    join_group(fe, id, fe->abs_pos);
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_info(sdrpp_cuda_frontend* fe, int id, int* info) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    if (!info) return fail(SDRPP_ERR_ARG, "null info");
    const VfoPlan& p = *v->plan;
    info[0] = p.rp.mode; info[1] = p.rp.predec; info[2] = p.rp.interp; info[3] = p.rp.decim;
    info[4] = (int)p.rp.taps.size(); info[5] = p.rp.tpp; info[6] = p.filter_needed ? p.chan_taps : 0;
    info[7] = p.s1_fir ? p.s1_D : 1; info[8] = p.s1_fir ? p.s1_T : 0;
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_submit(sdrpp_cuda_frontend* fe, int fmt, const void* in, int count) {
    return submit_common(fe, fmt, in, count, false);
}
int sdrpp_cuda_frontend_submit_device(sdrpp_cuda_frontend* fe, int fmt, const void* dev_in, int count) {
    return submit_common(fe, fmt, dev_in, count, true);
}
int sdrpp_cuda_frontend_join_streams(sdrpp_cuda_frontend* fe) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    std::lock_guard<std::mutex> api(fe->api_mtx);
    if (!fe->ev_join) FE_TRY(fe, cudaEventCreateWithFlags(&fe->ev_join, cudaEventDisableTiming));
    for (cudaStream_t s : { fe->st_fft, fe->st_tail, fe->st_d2h, fe->st_s1b, fe->st_bcast, fe->st_copy }) {
        FE_TRY(fe, cudaEventRecord(fe->ev_join, s));
        FE_TRY(fe, cudaStreamWaitEvent(fe->st, fe->ev_join, 0));
    }
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_submit_shared(sdrpp_cuda_frontend* fe, int fmt, int count) {
    if (!fe || !fe->comm || fe->comm->nranks < 2 || fe->comm->rank == fe->comm_root)
        return fail(SDRPP_ERR_STATE, "submit_shared is for the non-root ranks of an attached communicator");
    return submit_common(fe, fmt, nullptr, count, true);
}
int sdrpp_cuda_frontend_set_comm(sdrpp_cuda_frontend* fe, sdrpp_cuda_comm* c, int root) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (c && (root < 0 || root >= c->nranks)) return fail(SDRPP_ERR_ARG, "root out of range");
    if (c && c->device != fe->device) return fail(SDRPP_ERR_ARG, "communicator and front end live on different devices");
    fe->comm = c; fe->comm_root = root;
    // the persistent stage-1 kernel takes one CTA per SM with all of its shared memory: leave a few SMs to the NCCL
    // kernel, or the broadcast of block i+1 queues behind stage 1 of block i
    if (c && c->nranks > 1 && !getenv("SDRPP_RESERVE_SMS")) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, fe->device) == cudaSuccess && sms > 16) fe->num_sms = sms - 8;
    }
    return SDRPP_OK;
}
int sdrpp_cuda_frontend_submit_pcm(sdrpp_cuda_frontend* fe, const void* packet, int nbytes) {
    if (!packet || nbytes < 8) return fail(SDRPP_ERR_ARG, "bad argument");
    const PcmHeader h = pcm_parse(packet, nbytes);
    if (h.fmt < 0 || h.nsamples == 0) return 0; // the decompressor emits nothing (sample_stream_decompressor.h:35,43-48)
    const int rc = submit_common(fe, h.fmt, (const uint8_t*)packet + 8, h.nsamples, false, h.divisor);
    return rc == SDRPP_OK ? h.nsamples : rc;
}

int sdrpp_cuda_frontend_wait(sdrpp_cuda_frontend* fe) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    std::unique_lock<std::mutex> api(fe->api_mtx);
    if (fe->seq == 0) return fail(SDRPP_ERR_STATE, "nothing submitted");
    // oldest block not yet waited for
    if (fe->waited < fe->seq - kSets) fe->waited = fe->seq - kSets;
    if (fe->waited >= fe->seq) fe->waited = fe->seq - 1;
    const int slot = (int)(fe->waited % kSets);
    ResultSet& rs = fe->rs[slot];
    if (rs.pending) {
        // block on the completion event without the lock: the submitting thread may enqueue the next blocks meanwhile
        cudaEvent_t done = rs.done;
        api.unlock();
        cudaError_t e = cudaEventSynchronize(done);
        api.lock();
        if (e != cudaSuccess) {
            fe->sticky = std::string("cudaEventSynchronize(done): ") + cudaGetErrorString(e);
            set_last_error(fe->sticky);
            return SDRPP_ERR_CUDA;
        }
        rs.pending = false;
    }
    fe->cur = slot;
    fe->waited++;
    if (fe->profiling && fe->pev_valid && fe->waited == fe->seq) {
        for (int i = 0; i < 4; i++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, fe->pev[i], fe->pev[i + 1]) == cudaSuccess) fe->kernel_ms[i] = ms;
            else cudaGetLastError();
        }
    }
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_wait_input(sdrpp_cuda_frontend* fe) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> api(fe->api_mtx);
        if (fe->last_in_slot >= 0) ev = fe->ev_h2d[fe->last_in_slot];
    }
    if (ev) FE_TRY(fe, cudaEventSynchronize(ev));
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_drain(sdrpp_cuda_frontend* fe) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    std::lock_guard<std::mutex> api(fe->api_mtx);
    if (fe->seq > 0) { fe->waited = fe->seq; fe->cur = (int)((fe->seq - 1) % kSets); }
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_pending(sdrpp_cuda_frontend* fe) {
    if (!fe) return fail(SDRPP_ERR_ARG, "null front end");
    std::lock_guard<std::mutex> api(fe->api_mtx);
    long long w = fe->waited;
    if (w < fe->seq - kSets) w = fe->seq - kSets;
    return (int)(fe->seq - w);
}

int sdrpp_cuda_frontend_set_readback(sdrpp_cuda_frontend* fe, int enabled) {
    int rc = sdrpp_cuda_frontend_drain(fe); // a measurement control, not a reference setter: starts from a clean slate
    if (rc != SDRPP_OK) return rc;
    fe->readback = enabled != 0;
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_output(sdrpp_cuda_frontend* fe, int id, const sdrpp_cf32** iq, const float** demod) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    const ResultSet& rs = fe->rs[fe->cur];
    const bool have = (size_t)id < rs.counts.size();   // false: the VFO did not exist when that block was submitted
    const int n = have ? rs.counts[(size_t)id] : 0;
    const uint32_t off = have ? rs.offs[(size_t)id] : 0;
    if (iq) *iq = (rs.iq && have) ? rs.iq + off : nullptr;
    if (demod) *demod = (have && rs.demods[(size_t)id] != SDRPP_DEMOD_NONE && rs.demod) ? rs.demod + off : nullptr;
    return n;
}

int sdrpp_cuda_vfo_set_post(sdrpp_cuda_frontend* fe, int id, const sdrpp_cuda_post_cfg* cfg) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    if (!cfg) return fail(SDRPP_ERR_ARG, "null cfg");
    if (cfg->enabled && (!(cfg->agc_attack >= 0) || !(cfg->agc_decay >= 0) || !(cfg->dc_block_rate >= 0))) return fail(SDRPP_ERR_ARG, "negative coefficient");
    const sdrpp_cuda_post_cfg old = v->post;
    v->post = *cfg;
    if ((rc = apply_post(fe, *v)) != SDRPP_OK) { v->post = old; v->post.enabled = 0; apply_post(fe, *v); return rc; }
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_set_if_chain(sdrpp_cuda_frontend* fe, int id, const sdrpp_cuda_if_cfg* cfg) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    if (!cfg) return fail(SDRPP_ERR_ARG, "null cfg");
    if (v->plan->cap_final > kIfMaxBlock) return fail(SDRPP_ERR_ARG, "IF chain: the VFO's block output exceeds the tail kernel's staging area");
    if (cfg->fmif_bins != 0 && (cfg->fmif_bins < 2 || cfg->fmif_bins > kIfMaxBins)) return fail(SDRPP_ERR_ARG, "fmif_bins must be 0 or 2..64");
    float rec[IF_FLOATS] = { 0 };
    if (!v->if_state) {
        if (!cfg->nb_enabled && !cfg->squelch_enabled && !cfg->fmif_bins) return SDRPP_OK;
        FE_TRY(fe, dev_alloc(&v->if_state, (size_t)IF_FLOATS));
        rec[IF_NB_AMP] = 1.0f; // noise_blanker.h:77; Squelch: _isMute = false (squelch.h:79)
        fe->layout_dirty = true;
    } else {
        FE_TRY(fe, cudaMemcpy(rec, v->if_state, sizeof(rec), cudaMemcpyDeviceToHost));
    }
    // float members assigned from doubles: _rate = rate; _invRate = 1.0f - _rate; _level = level (noise_blanker.h:13-16)
    rec[IF_NB_ON] = cfg->nb_enabled ? 1.0f : 0.0f;
    rec[IF_NB_RATE] = (float)cfg->nb_rate;
    rec[IF_NB_INVRATE] = 1.0f - rec[IF_NB_RATE];
    rec[IF_NB_LEVEL] = (float)cfg->nb_level;
    rec[IF_SQ_ON] = cfg->squelch_enabled ? 1.0f : 0.0f;
    rec[IF_SQ_LEVEL] = (float)cfg->squelch_level;
    if ((int)rec[IF_FMIF_BINS] != cfg->fmif_bins) {
        // setBins: destroyBuffers + initBuffers (fm_if.h:26-35,91-115) -- cleared history, Nuttall window over bins - 1
        const int n = cfg->fmif_bins;
        for (int i = IF_HIST; i < IF_FLOATS; i++) rec[i] = 0.0f;
        for (int i = 0; i < n; i++) {
            rec[IF_WIN + i] = (float)window_nuttall((double)i, (double)(n - 1));
            const double a = -2.0 * 3.14159265358979323846 * (double)i / (double)n;
            rec[IF_TW + 2 * i] = (float)cos(a); rec[IF_TW + 2 * i + 1] = (float)sin(a);
        }
        rec[IF_FMIF_BINS] = (float)n;
    }
    FE_TRY(fe, upload_sync(v->if_state, rec, sizeof(rec)));
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_squelch_state(sdrpp_cuda_frontend* fe, int id, int* muted, float* level_db) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    if (!v->if_state) return fail(SDRPP_ERR_STATE, "no IF chain on this VFO");
    float rec[IF_FLOATS];
    FE_TRY(fe, cudaMemcpy(rec, v->if_state, sizeof(rec), cudaMemcpyDeviceToHost));
    if (muted) *muted = rec[IF_SQ_MUTE] != 0.0f;
    if (level_db) *level_db = rec[IF_SQ_LAST_DB];
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_audio(sdrpp_cuda_frontend* fe, int id, const float** audio) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    if (!v->post.enabled) return fail(SDRPP_ERR_STATE, "post-detector stages are not enabled for this VFO");
    const ResultSet& rs = fe->rs[fe->cur];
    const bool have = (size_t)id < rs.counts.size() && rs.has_audio[(size_t)id];
    if (audio) *audio = (rs.audio && have) ? rs.audio + rs.offs[(size_t)id] : nullptr;
    return have ? rs.counts[(size_t)id] : 0;
}

int sdrpp_cuda_vfo_rds(sdrpp_cuda_frontend* fe, int id, const sdrpp_cf32** rds) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    if (!(v->post.enabled && v->post.wfm && v->post.wfm_rds)) return fail(SDRPP_ERR_STATE, "the RDS output is not enabled for this VFO");
    const ResultSet& rs = fe->rs[fe->cur];
    const bool have = (size_t)id < rs.rds_counts.size() && rs.rds_counts[(size_t)id] >= 0 && rs.rds;
    if (rds) *rds = have ? reinterpret_cast<const sdrpp_cf32*>(rs.rds + rs.rds_offs[(size_t)id]) : nullptr;
    return have ? rs.rds_counts[(size_t)id] : 0;
}

int sdrpp_cuda_vfo_audio_stereo(sdrpp_cuda_frontend* fe, int id, const float** left, const float** right) {
    const float* l = nullptr;
    const int n = sdrpp_cuda_vfo_audio(fe, id, &l);
    if (n < 0) return n;
    const ResultSet& rs = fe->rs[fe->cur];
    const bool st = (size_t)id < rs.stereo.size() && rs.stereo[(size_t)id] && rs.audio_r && l;
    if (left) *left = l;
    if (right) *right = st ? rs.audio_r + rs.offs[(size_t)id] : l;
    return n;
}

int sdrpp_cuda_fft_rows(sdrpp_cuda_frontend* fe, const float** rows) {
    if (!fe) return fail(SDRPP_ERR_ARG, "null front end");
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    const ResultSet& rs = fe->rs[fe->cur];
    if (rows) *rows = rs.rows;
    return rs.nrows;
}

int sdrpp_cuda_frontend_set_fft_zoom(sdrpp_cuda_frontend* fe, double viewOffset, double viewBandwidth, double wholeBandwidth,
                                     int outSize, int keep_raw) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (outSize < 0 || outSize > (1 << 20) || (outSize > 0 && (!(viewBandwidth > 0) || !(wholeBandwidth > 0)))) return fail(SDRPP_ERR_ARG, "bad zoom view");
    fe->zoom_out = outSize; fe->zoom_keep_raw = keep_raw != 0;
    fe->zoom_view[0] = viewOffset; fe->zoom_view[1] = viewBandwidth; fe->zoom_view[2] = wholeBandwidth;
    return configure_zoom(fe);
}

int sdrpp_cuda_frontend_set_fft_display(sdrpp_cuda_frontend* fe, int smoothing, float smoothingSpeed, int hold, float holdSpeed) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (fe->zoom_out <= 0 || !fe->d_disp) return fail(SDRPP_ERR_STATE, "smoothing / peak hold work on the zoomed row: call sdrpp_cuda_frontend_set_fft_zoom first");
    const int W = fe->zoom_out;
    if (smoothing && !fe->disp_smoothing) {
        // setFFTSmoothing(true): the smoothing buffer starts as a copy of the latest row (waterfall.cpp:1197-1201)
        FE_TRY(fe, cudaMemcpy(fe->d_disp, fe->d_disp + 2 * W, (size_t)W * sizeof(float), cudaMemcpyDeviceToDevice));
        FE_TRY(fe, cudaStreamSynchronize(cudaStreamLegacy));
    }
    if (hold && !fe->disp_hold) {
        // setFFTHold(true): the hold row restarts at -1000 dB (waterfall.cpp:1169-1177)
        std::vector<float> init((size_t)W, -1000.0f);
        FE_TRY(fe, upload_sync(fe->d_disp + W, init.data(), init.size() * sizeof(float)));
    }
    fe->disp_smoothing = smoothing != 0; fe->disp_alpha = smoothingSpeed;
    fe->disp_hold = hold != 0; fe->disp_hold_speed = holdSpeed;
    return SDRPP_OK;
}

int sdrpp_cuda_fft_hold_row(sdrpp_cuda_frontend* fe, const float** row) {
    if (!fe) return fail(SDRPP_ERR_ARG, "null front end");
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    const ResultSet& rs = fe->rs[fe->cur];
    if (row) *row = rs.hold;
    return (fe->disp_hold && rs.hold && rs.nrows > 0) ? fe->zoom_out : 0;
}

int sdrpp_cuda_vfo_set_signal_info(sdrpp_cuda_frontend* fe, int id, int enabled) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    Vfo* v;
    if ((rc = vfo_get(fe, id, &v)) != SDRPP_OK) return rc;
    if (enabled && fe->cfg.fft_size <= 0) return fail(SDRPP_ERR_STATE, "the level / SNR read-out needs the spectrum branch (fft_size > 0)");
    v->sig_on = enabled != 0;
    if (!enabled) { v->sig_snr = 0.0f; v->sig_levels.clear(); }
    fe->sig_dirty = true;
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_set_snr_smoothing(sdrpp_cuda_frontend* fe, int enabled, float speed) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    fe->snr_smoothing = enabled != 0; fe->snr_alpha = speed;
    return SDRPP_OK;
}

int sdrpp_cuda_vfo_signal_info(sdrpp_cuda_frontend* fe, int id, float* strength, float* snr, float* level_max, int cap) {
    Vfo* v;
    int rc = vfo_get(fe, id, &v);
    if (rc != SDRPP_OK) return rc;
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    ResultSet& rs = fe->rs[fe->cur];
    if ((size_t)id >= rs.sig_slot.size() || rs.sig_slot[(size_t)id] < 0 || rs.nsig <= 0 || !rs.sig) return 0;
    const int slot = rs.sig_slot[(size_t)id];
    const int n = std::min(rs.nrows, std::max(cap, 0));
    // the scalar recurrences of WaterFall::pushFFT (waterfall.cpp:927-949) on the per-row device results: optional SNR
    // smoothing and the maximum of the last ten levels. Applied once per block, when its results are first read.
    if (!rs.sig_done.count(id)) {
        for (int r = 0; r < rs.nrows; r++) {
            float2& e = rs.sig[(size_t)r * rs.nsig + slot];
            if (fe->snr_smoothing) { v->sig_snr = ((1.0f - fe->snr_alpha) * v->sig_snr) + (fe->snr_alpha * e.y); e.y = v->sig_snr; }
            else v->sig_snr = e.y;
            v->sig_levels.push_back(e.x);
            if (v->sig_levels.size() > 10) v->sig_levels.erase(v->sig_levels.begin());
        }
        rs.sig_done.insert(id);
    }
    for (int r = 0; r < n; r++) {
        const float2 e = rs.sig[(size_t)r * rs.nsig + slot];
        if (strength) strength[r] = e.x;
        if (snr) snr[r] = e.y;
    }
    if (level_max) {
        float m = -INFINITY;
        for (float l : v->sig_levels) m = std::max(m, l);
        *level_max = m;
    }
    return rs.nrows;
}

int sdrpp_cuda_fft_zoomed_rows(sdrpp_cuda_frontend* fe, const float** rows) {
    if (!fe) return fail(SDRPP_ERR_ARG, "null front end");
    if (fe->cur < 0) return fail(SDRPP_ERR_STATE, "no completed block");
    const ResultSet& rs = fe->rs[fe->cur];
    if (rows) *rows = rs.zoom;
    return (fe->zoom_out > 0 && rs.zoom) ? rs.nrows : 0;
}

int sdrpp_cuda_frontend_read_iq(sdrpp_cuda_frontend* fe, sdrpp_cf32* out, int cap) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    const int n = std::min(cap, fe->last_count);
    if (n <= 0 || !out) return 0;
    const uint32_t start = (uint32_t)((uint64_t)(fe->abs_pos - fe->last_count) & fe->ring_mask);
    const uint32_t len = fe->ring_mask + 1u;
    const uint32_t first = std::min<uint32_t>((uint32_t)n, len - start);
    FE_TRY(fe, cudaMemcpy(out, fe->ring + start, (size_t)first * 8, cudaMemcpyDeviceToHost));
    if (first < (uint32_t)n) FE_TRY(fe, cudaMemcpy(out + first, fe->ring, (size_t)(n - first) * 8, cudaMemcpyDeviceToHost));
    return n;
}

long long sdrpp_cuda_frontend_launches(sdrpp_cuda_frontend* fe) { return fe ? fe->launches : 0; }
long long sdrpp_cuda_frontend_stage1_tensor_launches(sdrpp_cuda_frontend* fe) { return fe ? fe->s1t_launches : 0; }
int sdrpp_cuda_frontend_set_stage1_mode(sdrpp_cuda_frontend* fe, int mode) {
    int rc = fe_quiesce(fe);
    if (rc != SDRPP_OK) return rc;
    if (mode != 0 && mode != 1) return fail(SDRPP_ERR_ARG, "stage-1 mode must be 0 (tensor cores where possible) or 1 (FP32 only)");
    fe->s1_mode = mode;
    return SDRPP_OK;
}
void* sdrpp_cuda_frontend_stream(sdrpp_cuda_frontend* fe) { return fe ? (void*)fe->st : nullptr; }
int sdrpp_cuda_frontend_set_graphs(sdrpp_cuda_frontend* fe, int enabled) {
    int rc = fe_check(fe);
    if (rc != SDRPP_OK) return rc;
    std::lock_guard<std::mutex> api(fe->api_mtx);
    fe->graphs_on = enabled != 0;
    return SDRPP_OK;
}

// Timeline probe: for each of the last blocks, milliseconds of the eight marks relative to mark 0 of the OLDEST block kept.
// out[b * 11 + 0] = block number, out[b * 11 + 1 + k] = time of mark k (0 main start, 1 ingest/split done, 2 stage 1 done, 3 spectrum
// start, 4 spectrum done, 5 tail start, 6 wide stage done, 7 tail done, 8 host-to-device copy done, 9 results on the host).
// Returns the number of blocks written.
extern "C" __attribute__((visibility("default"))) int sdrpp_cuda_debug_timeline(sdrpp_cuda_frontend* fe, float* out, int cap_blocks) {
    if (!fe || !fe->timeline || !out) return 0;
    cudaDeviceSynchronize();
    constexpr int NB = sdrpp_cuda_frontend::kTlBlocks, NM = sdrpp_cuda_frontend::kTlMarks;
    int oldest = 0;
    for (int i = 1; i < NB; i++) if (fe->tl_blk[i] < fe->tl_blk[oldest]) oldest = i;
    int n = 0;
    for (int k = 0; k < NB && n < cap_blocks; k++) {
        const int slot = (oldest + k) % NB;
        out[n * 11] = (float)fe->tl_blk[slot];
        for (int m = 0; m < NM; m++) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, fe->tl_ev[oldest][0], fe->tl_ev[slot][m]) != cudaSuccess) { cudaGetLastError(); ms = -1.0f; }
            out[n * 11 + 1 + m] = ms;
        }
        n++;
    }
    return n;
}

int sdrpp_cuda_frontend_graph_stats(sdrpp_cuda_frontend* fe, long long* out4) {
    if (!fe || !out4) return fail(SDRPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> api(fe->api_mtx);
    out4[0] = fe->graph_replays; out4[1] = fe->graph_captures; out4[2] = fe->direct_runs; out4[3] = fe->capture_us;
    return SDRPP_OK;
}

int sdrpp_cuda_frontend_set_profiling(sdrpp_cuda_frontend* fe, int enabled) {
    int rc = sdrpp_cuda_frontend_drain(fe);
    if (rc != SDRPP_OK) return rc;
    fe->profiling = enabled != 0;
    fe->pev_valid = false;
    return SDRPP_OK;
}
float sdrpp_cuda_frontend_kernel_ms(sdrpp_cuda_frontend* fe, int idx) {
    if (!fe || idx < 0 || idx > 3) return 0.f;
    return fe->kernel_ms[idx];
}

} // extern "C"
