// Kernel launch interfaces shared between the .cu files and the engine.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "launcher.h"

namespace sdrpp {

// ---------------------------------------------------------------------------------------------
// Device IQ ring: cf32 samples at (index & mask). mask = 0xFFFFFFFF addresses a linear buffer.
// ---------------------------------------------------------------------------------------------
struct RingRef {
    float2* base;
    uint32_t mask;
};

// ---------------------------------------------------------------------------------------------
// Ingest / pre-processing (convert.cu, preproc.cu)
// ---------------------------------------------------------------------------------------------
// raw samples of format fmt -> cf32 at dst[(pos+i)&mask], optional conjugate (dsp/math/conjugate.h:12-15)
// (L, sid): the command goes to stream `sid` of the launcher; per-block arguments travel in the launcher's descriptor.
cudaError_t launch_ingest(Launcher& L, int sid, int fmt, const void* raw, int count, RingRef dst, uint32_t pos, bool conj, float scale = 1.0f);
// SampleStreamCompressor (sample_stream_compressor.h:26-60): block maximum, then int8/int16 packing.
// key: one device word of scratch; scaler_out: device float that receives the packet's scaler.
cudaError_t launch_pcm_compress(int bits, const float* in, int nscalars, unsigned int* key, void* out, float* scaler_out,
                                cudaStream_t st);

// One decimating FIR stage of the front-end PowerDecimator (dsp/filter/decimating_fir.h:45-68):
// out[m] = sum_k buf[offset + m*D + k] * taps[k], buf = [T-1 history | count new samples] (linear).
cudaError_t launch_decim_stage(const float2* buf, const float* taps, int T, int D, int offset, int nout,
                               RingRef dst, uint32_t pos, bool conj, cudaStream_t st);
// Move the last `hist` samples of buf[0 .. hist+count) to the front (history carry, fir.h:80).
cudaError_t launch_shift_history(float2* buf, int hist, int count, cudaStream_t st);
// DC blocker (dsp/correction/dc_blocker.h:54-60) as a chunked linear-recurrence scan.
// state: one float2 (the offset estimate) in device memory, updated in place. scratch: >= 2*nchunks float2.
cudaError_t launch_dc_block(const float2* in, int count, float rate, float2* state, float2* scratch,
                            RingRef dst, uint32_t pos, bool conj, cudaStream_t st, long long* launches);
int dc_block_chunks(int count);

// ---------------------------------------------------------------------------------------------
// Spectrum (fft.cu)
// ---------------------------------------------------------------------------------------------
struct SpectrumArgs {
    const void* in;        // cf32 samples (ring or linear)
    uint32_t ring_mask;
    uint32_t start;        // index of the first sample of frame 0
    uint32_t frame_stride; // samples between frame starts
    int nz;                // windowed samples per frame (rest of the N-point input is zero)
    const float* window;   // nz floats
    float2* inter;         // four-step intermediate, frames*N (unused for N <= 4096)
    float* rows;           // frames*N dB rows (may be null)
    float2* X;             // frames*N complex spectra (may be null; parity/debug)
    int frames;
};
int spectrum_split(int N, int* N1, int* N2);
cudaError_t launch_spectrum(Launcher& L, int sid, int N, const SpectrumArgs& a, long long* launches);

// Waterfall zoom / max-decimation of dB rows (gui/widgets/fft_scaler.h:41-64): out[r][i] = max over bins
// [idx[i], idx[i+1]) (ranged) or rows[r][idx[i]] (point sampling). Reads are clamped to the row.
cudaError_t launch_fft_zoom(Launcher& L, int sid, const float* rows, int N, int nrows, const int* idx, int outSize, bool ranged, float* out);

// WaterFall::calculateVFOSignalInfo (gui/widgets/waterfall.cpp:563-603) for nsig VFOs x nrows rows: out[r*nsig + v] =
// (strength, snr); bins[v] = (minSide, min, max, maxSide) bin indices (design.h: signal_info_bins).
cudaError_t launch_signal_info(Launcher& L, int sid, const float* rows, int N, int nrows, const int4* bins, int nsig, float2* out);
// FFT smoothing and peak hold of the zoomed rows, in place (WaterFall::pushFFT, waterfall.cpp:918-956); state buffers of W floats.
cudaError_t launch_fft_display(Launcher& L, int sid, float* zoom, int W, int nrows, bool smoothing, float alpha, float* smooth, bool hold_on,
                               float hold_speed, float* hold, float* latest);

// ---------------------------------------------------------------------------------------------
// Channelizer (channelizer.cu)
// ---------------------------------------------------------------------------------------------
// Per-VFO device record.
struct VfoDev {
    float2* slab;        // this VFO's stage buffers
    uint64_t phi_ref;    // NCO phase (turns * 2^64) of absolute input sample n_ref
    int64_t n_ref;
    uint64_t dphi;       // NCO phase step per input sample (turns * 2^64), from the fp32-quantised increment
    uint64_t dphi2;      // SSB second translation, per output sample
    uint32_t out_off;    // offset of this VFO's rows in the output arenas (samples)
    uint32_t pad;
    float* ifs;          // radio IF chain record (IF_* below) between the VFO output and the demod front end; null = off
};

// Radio IF chain record, one per VFO that has a block of the chain enabled (decoder_modules/radio/src/radio_module.h:
// 73-78: NoiseBlanker -> Squelch -> FMIF in front of the demodulator). Configuration and state: 16 scalars, then the
// FM IF noise reduction's input history (bins - 1 samples), window (bins floats) and DFT twiddles e^{-2 pi j m / bins}.
constexpr int kIfMaxBins = 64;
constexpr int kIfMaxBlock = 2176; // VFO output samples per block an IF chain can take (tail kernel staging area)
enum { IF_NB_ON = 0, IF_NB_RATE, IF_NB_INVRATE, IF_NB_LEVEL, IF_NB_AMP, IF_SQ_ON, IF_SQ_LEVEL, IF_SQ_MUTE, IF_SQ_CNT,
       IF_PREV_RE, IF_PREV_IM, IF_SQ_LAST_DB, IF_FMIF_BINS /* 0 = off */,
       IF_HIST = 16, IF_WIN = IF_HIST + 2 * kIfMaxBins, IF_TW = IF_WIN + kIfMaxBins, IF_FLOATS = IF_TW + 2 * kIfMaxBins };

constexpr int kStage1Warps = 8;     // warps per CTA, each owning R consecutive input rows

// Stage 1 of a group of VFOs sharing one plan: NCO translation + the first decimating FIR
// (dsp/channel/frequency_xlator.h:43-50 + dsp/filter/decimating_fir.h:45-68).
struct Stage1Args {
    RingRef ring;
    uint32_t ring_first;   // ring index of the first sample of the dot product of output 0
    int64_t abs_first;     // absolute index of that sample
    int64_t abs_valid;     // samples with absolute index < abs_valid read as zero (reset / stream start)
    int D, T, A;           // decimation, taps (+1 when the shifted form is used), rows of the tap matrix
    int tap_off;           // offset of this plan's zero-padded taps (plain or shifted form) in the constant tap pool
    int M;                 // outputs this block
    int opc;               // outputs per CTA (set by the launcher)
    int nvfo;              // VFOs in the group
    const float4* G;       // per-VFO phasor table F[p] = e^{j w p}: [vb][p/2][lane] = (F[p], F[p+1])
    const VfoDev* vfos;    // group members, contiguous
    uint32_t out_off;      // slab offset (in float2) where output 0 goes
};
bool stage1_supported(int A, int D);
int stage1_A(int T, int D);                                // rows of the tap matrix, ceil((T+1)/D)
int stage1_tap_offset(int ratio);                          // also uploads the pool to the current device
size_t stage1_g_elems(int A, int D, int nvfo);          // float4 elements of G
void stage1_g_index(int A, int D, int v, int p, size_t* idx4, int* half); // where F[p] of VFO v lives
cudaError_t launch_stage1(Launcher& L, int sid, const Stage1Args& a);
// D = 1, T = 1 (no pre-decimation): pure translation.
cudaError_t launch_mix_only(Launcher& L, int sid, const Stage1Args& a);

// ---------------------------------------------------------------------------------------------
// Stage 1 on the tensor cores (channelizer_tc.cu): same result as launch_stage1 for first-stage decimations of
// 32 and 64, computed as X (samples, fp16 hi/lo) * B (shifted taps x per-VFO phasors, fp16 hi/lo) with tcgen05.
// ---------------------------------------------------------------------------------------------
constexpr int kS1TVfosPerTile = 16;
constexpr int kS1TMaxGroups = 6;
// fp16 hi/lo copies of the IQ ring in the UMMA shared-memory image: [k-half][group of 8 rows][1024 B] per plane,
// row = D consecutive samples starting at an absolute index == origin (mod D); sinv[group] = 2^-e of the group's block scale.
struct S1TPlanes {
    int D;
    int origin;            // rows start at absolute sample indices == origin (mod D); multiple of 4, < D
    uint32_t group_mask;   // groups in the ring - 1
    uint8_t* hi;
    uint8_t* lo;
    float* sinv;
    int64_t valid_from;    // absolute sample index from which the planes hold converted samples (host bookkeeping)
};
struct S1TGroupArgs {
    const uint8_t* bblob;  // B images of the group's VFO tiles (s1t_b_bytes)
    const VfoDev* vfos;    // group members, contiguous
    int nvfo;
    int A;                 // rows of the shifted tap matrix, ceil((T + shift) / D)
    int n_ttiles;          // time tiles (120 outputs each) of this block
    int M;                 // outputs this block
    int64_t row0;          // absolute row of time tile 0's first row (multiple of 8)
    int64_t row_first;     // absolute row in which the window of output 0 starts
    uint32_t out_off;      // slab offset (in float2) where output 0 goes
    float b_scale_inv;     // 2^-(B block exponent)
    int n_vtiles, cta_per_vtile, cta_begin; // set by the launcher
};
struct S1TArgs {
    S1TPlanes pl;
    int ngroups;
    int nchunks;           // set by the launcher
    S1TGroupArgs g[kS1TMaxGroups];
};
bool s1t_supported(int T, int D);
int s1t_A(int T, int D, int shift);
int s1t_b_exponent(const float* taps, int T);
size_t s1t_b_bytes(int A, int D, int nvfo);
// samples [abs_begin, abs_end) were just written to the ring: (re)build the 8-row groups they touch
cudaError_t launch_s1t_split(Launcher& L, int sid, RingRef ring, const S1TPlanes& pl, int64_t abs_begin, int64_t abs_end,
                             const void* raw = nullptr, int64_t abs_block = 0, bool conj = false);
bool s1t_split_covers(const S1TPlanes& pl, int64_t abs_begin, int64_t abs_block, int64_t abs_end);
cudaError_t launch_s1t_build_b(Launcher& L, int sid, uint8_t* blob, const VfoDev* vfos, int nvfo, const float* d_taps, int T, int D, int shift,
                               int A, int escale);
cudaError_t launch_s1t(Launcher& L, int sid, S1TArgs& a, int num_sms);

enum { TAIL_DECFIR = 0, TAIL_POLY = 1, TAIL_FIR = 2 };
struct TailStage {
    int type;
    int T;          // taps (per phase for POLY)
    int D;          // decimation (DECFIR), polyphase decim (POLY), 1 (FIR)
    int interp;     // POLY
    const float* taps; // FIR taps or polyphase bank [interp][T]
    uint32_t in_off;   // slab offset of the data area of this stage's input; history sits just before it
    int n_in, n_out;
    int offset, phase; // integer state at the start of this block
};
constexpr int kTailMaxStages = 6;
constexpr int kTailMaxGroups = 6;
struct TailGroup {
    int first_vfo, nvfo;
    int nstages;
    int s_begin;         // first stage the tail kernel runs: 1 when stage 0 ran in tail_stage0_wide_kernel
    TailStage st[kTailMaxStages];
    uint32_t final_off;  // slab offset of the final output data area (one sample of history before it)
    uint32_t carry0_off; // data-area offset of the OTHER stage-1 region: receives the history carry of stage 0
                         // (or of the final output when there is no tail stage) for the next block
    int n_final;         // output samples this block
    int demod;
    float inv_dev;       // Quadrature: 1/(2*pi*dev/sr)
    int64_t abs_out;     // absolute index of the first output sample of this block (SSB phase)
};
struct TailArgs {
    int ngroups;
    TailGroup g[kTailMaxGroups];
    const VfoDev* vfos;
    float2* arena_iq;
    float* arena_demod;
};
// the three tail launchers take the DEVICE copy of the arguments (Launcher::push) beside the host copy they size the grid from
cudaError_t launch_tail(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos);
// the low-latency form for groups whose stage inputs of this block fit in shared memory at once (tail_fast_fits)
bool tail_fast_fits(const TailGroup& g, int* samples, int threads = 1024, int* tap_floats = nullptr);
cudaError_t launch_tail_fast(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos, int threads = 1024);
// stage 0 of the groups with s_begin == 1 on a wide grid (a decimating FIR); launch before launch_tail
cudaError_t launch_tail_stage0_wide(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos);
bool tail_stage0_wide_supported(int T, int D);

// Post-detector stages of the three demodulators: FM low-pass (dsp/demod/fm.h:86-103), AM
// [carrier AGC] -> magnitude -> DC block -> [audio AGC] -> low-pass (dsp/demod/am.h:114-146), SSB AGC
// (dsp/demod/ssb.h:90-101); dsp::loop::AGC (dsp/loop/agc.h:87-147). One CTA per VFO, after the tail.
enum { POST_NONE = 0, POST_FM = 1, POST_AM = 2, POST_SSB = 3, POST_WFM = 4 };
struct PostDev {
    int kind;            // POST_*
    int mode;            // FM: 1 = low-pass; AM: AGCMode (0 OFF, 1 CARRIER, 2 AUDIO); SSB: 1 = AGC enabled
    int ntaps;           // float FIR taps (0: no filter)
    int hist_pad;        // floats reserved for the FIR history in front of the work area (>= ntaps - 1)
    const float* taps;
    float* state;        // [0] audio amp [1] audio gain [2] carrier amp [3] carrier gain [4] dc offset, ... [16 ..] history | work
    uint32_t out_off;    // offset of this VFO's rows in the output arenas (samples)
    float attack, inv_attack, decay, inv_decay;
    float dc_rate, set_point, max_gain, max_out;
    // POST_WFM (dsp::demod::BroadcastFM, demod/broadcast_fm.h): mode bit 0 = stereo, bit 1 = low-pass; taps = the 19 kHz pilot
    // band-pass (complex, interleaved, ntaps of them), taps2 = the 15 kHz audio low-pass (ntaps2); PLL coefficients and limits
    const float* taps2;
    int ntaps2, delay, cap;
    float pll_alpha, pll_beta, pll_min_freq, pll_max_freq;
    float* out_r;        // right channel rows (arena), indexed like out_off
};
struct PostArgs {
    int ngroups;
    struct { int first_vfo, nvfo, n; } g[kTailMaxGroups];
    const PostDev* post;          // indexed like the VfoDev table
    const float2* arena_iq;
    const float* arena_demod;
    float* arena_audio;
    float* arena_audio_r;        // right channel of stereo demodulators (POST_WFM); mono kinds leave it alone
};
cudaError_t launch_post(Launcher& L, int sid, const PostArgs& a, int total_vfos);

// RDS side output of dsp::demod::BroadcastFM (demod/broadcast_fm.h:168-175,188-198): the discriminator output as (mpx, 0),
// translated by -57 kHz (FrequencyXlator, 64-bit phase accumulator on the fp32-quantised increment) and resampled to 5 kS/s by
// the decoder's RationalResampler (PowerDecimator stages, then the polyphase resampler). One CTA per VFO with the output on.
constexpr int kRdsMaxStages = 4;
constexpr int kRdsPerLaunch = 64;
struct RdsDev {
    uint32_t in_off;                 // the VFO's row in the demod arena
    uint32_t out_off;                // its row in the RDS arena (float2)
    int nstages;                     // PowerDecimator stages (0: RESAMP_ONLY / NONE)
    int T[kRdsMaxStages], D[kRdsMaxStages];
    const float* taps[kRdsMaxStages];
    uint32_t buf_off[kRdsMaxStages]; // float2 offset of [T-1 history | block] in state
    int interp, decim, tpp;          // polyphase resampler (tpp = 0: none)
    const float* bank;               // [interp][tpp]
    uint32_t pbuf_off;               // [tpp-1 history | block]
    float2* state;
    uint64_t dphi;
};
// per block and VFO: every integer of the reference's state machines, computed on the host (decimating_fir.h:51-62,
// polyphase_resampler.h:75-93) -- the device keeps sample history only
struct RdsBlk {
    int n;                           // discriminator samples of this block
    int off[kRdsMaxStages], nout[kRdsMaxStages];
    int pphase, poff, npoly;         // polyphase state at the start of the block, outputs of the block
    int nfinal;
    uint64_t phase0;                 // NCO phase of the block's first sample
};
struct RdsArgs {
    const RdsDev* tab;               // table of the VFOs with the output on
    const float* arena_demod;
    float2* arena_rds;
    int first, count;                // table range of this launch
    RdsBlk blk[kRdsPerLaunch];
};
cudaError_t launch_rds(Launcher& L, int sid, const RdsArgs& a);

} // namespace sdrpp
