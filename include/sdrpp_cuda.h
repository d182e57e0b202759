/* sdrpp_cuda.h -- C ABI of the B200-native SDR++ signal-path hot loop.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. The host-side C++
 * mirror of the reference's dsp:: / sigpath:: interface (include/sdrpp/...) forwards to these
 * entry points; SDR++ modules never see them. Each entry point cites the reference interface it
 * replaces (paths relative to the reference tree's core/src unless noted).
 *
 * Conventions: every int-returning call returns >= 0 on success and < 0 on error
 * (SDRPP_ERR_*); the message of the last error on the calling thread is
 * sdrpp_cuda_last_error(). There is NO CPU fallback: without a CUDA device every compute call
 * fails with SDRPP_ERR_CUDA.
 */
#ifndef SDRPP_CUDA_H
#define SDRPP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SDRPP_API __attribute__((visibility("default")))
#else
#define SDRPP_API
#endif

/* dsp::complex_t (dsp/types.h:6-91): interleaved {float re, im} */
typedef struct { float re, im; } sdrpp_cf32;

enum {
    SDRPP_OK = 0,
    SDRPP_ERR_ARG = -1,      /* bad argument (the reference asserts or returns NULL) */
    SDRPP_ERR_CUDA = -2,     /* CUDA runtime error, sticky per handle */
    SDRPP_ERR_STATE = -3,    /* call not valid in this state (e.g. duplicate / unknown id) */
    SDRPP_ERR_NOMEM = -4
};

/* Input sample formats = the per-source conversion formulas on the reference's input edge. */
enum {
    SDRPP_FMT_CF32 = 0,      /* already dsp::complex_t */
    SDRPP_FMT_U8_RTL = 1,    /* (u8-128+0.5f)/(128.0f-0.5f): source_modules/rtl_sdr_source/src/main.cpp:526-527, file_source/src/main.cpp:489 */
    SDRPP_FMT_U8_TCP = 2,    /* (float)(((double)u8-128.0)/128.0): source_modules/rtl_tcp_source/src/rtl_tcp_client.cpp:86-87 */
    SDRPP_FMT_I8 = 3,        /* volk_8i_s32f_convert_32f(..,128.0f): source_modules/hackrf_source/src/main.cpp:386 */
    SDRPP_FMT_I16_FILE = 4,  /* (i16+0.5f)/(32768.0f-0.5f): source_modules/file_source/src/main.cpp:506 */
    SDRPP_FMT_I16_VOLK = 5,  /* volk_16i_s32f_convert_32f(..,32768): bladerf_source/src/main.cpp:587, plutosdr_source/src/main.cpp:261-265 */
    SDRPP_FMT_I24_FILE = 6,  /* packed LE 24-bit, (i24+0.5f)/(8388608.0f-0.5f): source_modules/file_source/src/main.cpp:525 */
    SDRPP_FMT_I32_FILE = 7,  /* (float)((i32+0.5)/(2147483648.0-0.5)): source_modules/file_source/src/main.cpp:542 */
    SDRPP_FMT_F64 = 8,       /* volk_64f_convert_32f: source_modules/file_source/src/main.cpp:475 */
    SDRPP_FMT_COUNT = 9
};

/* dsp::window::windowType (dsp/window/window.h:28-36) -- persisted as an int in the config */
enum {
    SDRPP_WIN_RECTANGULAR = 0, SDRPP_WIN_HAMMING, SDRPP_WIN_HANN, SDRPP_WIN_BLACKMAN,
    SDRPP_WIN_NUTTALL, SDRPP_WIN_BLACKMAN_HARRIS4, SDRPP_WIN_BLACKMAN_HARRIS7, SDRPP_WIN_COUNT
};

/* Demodulator front end fused behind a VFO. */
enum {
    SDRPP_DEMOD_NONE = 0,
    SDRPP_DEMOD_QUADRATURE = 1, /* dsp::demod::Quadrature::process (dsp/demod/quadrature.h:41-56), deviation = bw/2 (demod/fm.h:31) */
    SDRPP_DEMOD_AM = 2,         /* volk_32fc_magnitude_32f in dsp::demod::AM::process (dsp/demod/am.h:122) */
    SDRPP_DEMOD_USB = 3,        /* dsp::demod::SSB::process xlate(+bw/2) + ComplexToReal (dsp/demod/ssb.h:90-101,119-126) */
    SDRPP_DEMOD_LSB = 4,        /* same, -bw/2 */
    SDRPP_DEMOD_DSB = 5         /* same, 0 */
};

/* ---------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------ */
SDRPP_API const char* sdrpp_cuda_version(void);
SDRPP_API const char* sdrpp_cuda_last_error(void);
SDRPP_API int sdrpp_cuda_device_count(void);
/* Select the CUDA device used by objects created afterwards on this thread. */
SDRPP_API int sdrpp_cuda_init(int device);
/* Pinned host memory for IQ blocks (replaces volk_malloc in dsp::stream, dsp/stream.h:28-29,125-126). */
SDRPP_API void* sdrpp_cuda_host_alloc(size_t bytes);
SDRPP_API void sdrpp_cuda_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * Host-side design maths (double precision, no GPU needed). These define filter shapes and index
 * arithmetic and must agree exactly with the reference.
 * ------------------------------------------------------------------------------------------ */
/* dsp::window::createWindow (dsp/window/window.h:38-64). buf holds size+1 floats. */
SDRPP_API int sdrpp_cuda_design_window(int type, float* buf, int size, int centered);
/* dsp::taps::lowPass (dsp/taps/low_pass.h:7-11). Returns the tap count; writes min(count,cap). */
SDRPP_API int sdrpp_cuda_design_lowpass(double cutoff, double transWidth, double sampleRate, float* out, int cap);
/* dsp::taps::bandPass<dsp::complex_t> (dsp/taps/band_pass.h:10-25): complex band-pass taps, interleaved (re, im); the pilot
 * filter of the WFM stereo decoder is bandPass(18750, 19250, 3000, sampleRate, true). Returns the tap count. */
SDRPP_API int sdrpp_cuda_design_bandpass_complex(double bandStart, double bandStop, double transWidth, double sampleRate,
                                                 int oddTapCount, float* out, int cap);
/* dsp::multirate::RationalResampler::reconfigure (dsp/multirate/rational_resampler.h:121-167).
 * info[0]=mode (0 BOTH,1 DECIM_ONLY,2 RESAMP_ONLY,3 NONE) [1]=predec ratio [2]=interp [3]=decim
 * [4]=tap count [5]=taps per phase; taps (optional) are already scaled by interp. */
SDRPP_API int sdrpp_cuda_design_resampler(double inSR, double outSR, int* info, float* taps, int cap);
/* dsp::multirate::decim::plans (dsp/multirate/decim/plans.h:126-140). Returns the stage count. */
SDRPP_API int sdrpp_cuda_design_decim_plan(int ratio, int* decimation, int* tapcount, const float** taps);
/* IQFrontEnd::genReshapeParams (signal_path/iq_frontend.h:56-60). */
SDRPP_API void sdrpp_cuda_design_reshape(double sampleRate, int fftSize, double fftRate, int* skip, int* nz);

/* ---------------------------------------------------------------------------------------------
 * One-shot block operations on HOST buffers (H2D + kernel + D2H inside the call)
 * ------------------------------------------------------------------------------------------ */
/* Source conversions (see SDRPP_FMT_*). in: nsamples interleaved I,Q pairs. Bit-exact. */
SDRPP_API int sdrpp_cuda_convert(int fmt, const void* in, int nsamples, sdrpp_cf32* out);
/* SDR++ server wire packets (core/src/dsp/compression/): u16 compression type (0) | u16 PCMType |
 * f32 scaler | payload. PCMType values: core/src/dsp/compression/pcm_type.h:4-8. */
enum { SDRPP_PCM_I8 = 0, SDRPP_PCM_I16 = 1, SDRPP_PCM_F32 = 2 };
/* dsp::compression::SampleStreamDecompressor::process (sample_stream_decompressor.h:13-36), the
 * client side (source_modules/sdrpp_server_source): packet of nbytes -> cf32. Returns the number of
 * complex samples written (0 for an unknown sample type, like the reference), < 0 on error.
 * I8/I16 payloads are divided by 128.0f/scaler resp. 32768.0f/scaler in fp32. Bit-exact. */
SDRPP_API int sdrpp_cuda_pcm_decompress(const void* packet, int nbytes, sdrpp_cf32* out);
/* dsp::compression::SampleStreamCompressor::process (sample_stream_compressor.h:26-60), the server
 * side (core/src/server.cpp): count cf32 samples -> packet (capacity 8 + count*8 bytes is always
 * enough). scaler = the largest SIGNED scalar of the block, payload = saturated rintf(x*(128|32768)
 * /scaler). Returns the packet size in bytes, < 0 on error. Bit-exact for finite input whose
 * maximum is > 0 (NaN input and an all-non-positive block are undefined in the reference too). */
SDRPP_API int sdrpp_cuda_pcm_compress(int pcm_type, const sdrpp_cf32* in, int count, void* packet);
/* IQFrontEnd::handler (signal_path/iq_frontend.cpp:230-249): window * frame -> zero-padded forward
 * DFT of size N -> 10*log10|X|^2. frame: nz samples of format fmt; window: nz floats.
 * row: N floats (may be NULL); X: N complex FFT outputs (may be NULL; parity/debug). */
SDRPP_API int sdrpp_cuda_spectrum(int N, int nz, int fmt, const void* frame, const float* window,
                                  float* row, sdrpp_cf32* X);

/* Same transform on DEVICE buffers, for callers whose samples already live on the GPU: `frames` frames of nz
 * cf32 samples, frame f starting at dev_in + f*frame_stride; window: nz floats in HOST memory, or NULL to reuse
 * the table of the previous call with the same N and nz (skips the comparison of a large table); dev_rows:
 * frames*N floats in device memory. Asynchronous on `stream` (a cudaStream_t, NULL = the library's own stream,
 * in which case the call synchronises before returning). */
SDRPP_API int sdrpp_cuda_spectrum_device(int N, int nz, int frames, long long frame_stride, const sdrpp_cf32* dev_in,
                                         const float* window, float* dev_rows, void* stream);

/* Waterfall zoom / max-decimation of one dB row to outSize pixels: fft_scaler(viewOffset, viewBandwidth,
 * wholeBandwidth, N, outSize).doZoom (gui/widgets/fft_scaler.h:28-64, used by WaterFall::pushFFT,
 * gui/widgets/waterfall.cpp:900-904). idx (optional, outSize+1 ints) receives the bin boundaries. */
SDRPP_API int sdrpp_cuda_fft_zoom(int N, const float* row, double viewOffset, double viewBandwidth,
                                  double wholeBandwidth, int outSize, float* out, int* idx);

/* ---------------------------------------------------------------------------------------------
 * Front end: the device-resident signal path (sigpath::iqFrontEnd + sigpath::vfoManager's
 * dsp::channel::RxVFO set + demod front ends) of one GPU.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrpp_cuda_frontend sdrpp_cuda_frontend;

typedef struct {
    double sample_rate;   /* IQFrontEnd::init sampleRate */
    int decim_ratio;      /* decimRatio: 1 or 2^k (dsp/multirate/power_decimator.h:22-25) */
    int dc_blocking;      /* dcBlocking */
    int invert_iq;        /* setInvertIQ */
    int fft_size;         /* fftSize, power of two 2^6..2^22; 0 disables the spectrum branch */
    double fft_rate;      /* fftRate (lines/s) */
    int fft_window;       /* SDRPP_WIN_* */
    int max_block;        /* largest sample count of one submit (<= 1e6 in the reference, dsp/stream.h:9); 0 = 1000000 */
    int ring_log2;        /* log2 of the device IQ ring length in samples; 0 = auto */
    int max_fft_rows;     /* spectrum rows buffered per block; 0 = auto */
} sdrpp_cuda_frontend_cfg;

SDRPP_API sdrpp_cuda_frontend* sdrpp_cuda_frontend_create(const sdrpp_cuda_frontend_cfg* cfg);
SDRPP_API int sdrpp_cuda_frontend_destroy(sdrpp_cuda_frontend* fe);

/* IQFrontEnd setters (signal_path/iq_frontend.cpp:51-174). Applied at the next block boundary. */
SDRPP_API int sdrpp_cuda_frontend_set_sample_rate(sdrpp_cuda_frontend* fe, double sampleRate);
SDRPP_API int sdrpp_cuda_frontend_set_decimation(sdrpp_cuda_frontend* fe, int ratio);
SDRPP_API int sdrpp_cuda_frontend_set_dc_blocking(sdrpp_cuda_frontend* fe, int enabled);
SDRPP_API int sdrpp_cuda_frontend_set_invert_iq(sdrpp_cuda_frontend* fe, int enabled);
SDRPP_API int sdrpp_cuda_frontend_set_fft_size(sdrpp_cuda_frontend* fe, int size);
SDRPP_API int sdrpp_cuda_frontend_set_fft_rate(sdrpp_cuda_frontend* fe, double rate);
SDRPP_API int sdrpp_cuda_frontend_set_fft_window(sdrpp_cuda_frontend* fe, int window);
SDRPP_API double sdrpp_cuda_frontend_effective_samplerate(sdrpp_cuda_frontend* fe); /* getEffectiveSamplerate */

/* IQFrontEnd::addVFO / VFOManager::createVFO (signal_path/iq_frontend.cpp:122-142, vfo_manager.cpp:95-103)
 * -> dsp::channel::RxVFO(in, effectiveSr, outSR, bw, offset) (+ demod front end). Returns the VFO id (>=0). */
SDRPP_API int sdrpp_cuda_vfo_create(sdrpp_cuda_frontend* fe, double outSR, double bandwidth, double offset, int demod);
SDRPP_API int sdrpp_cuda_vfo_destroy(sdrpp_cuda_frontend* fe, int vfo);                       /* removeVFO */
SDRPP_API int sdrpp_cuda_vfo_set_offset(sdrpp_cuda_frontend* fe, int vfo, double offset);     /* RxVFO::setOffset, rx_vfo.h:72-77 */
SDRPP_API int sdrpp_cuda_vfo_set_bandwidth(sdrpp_cuda_frontend* fe, int vfo, double bw);      /* RxVFO::setBandwidth, rx_vfo.h:60-70 */
SDRPP_API int sdrpp_cuda_vfo_set_out_samplerate(sdrpp_cuda_frontend* fe, int vfo, double outSR, double bw); /* rx_vfo.h:46-58 */
SDRPP_API int sdrpp_cuda_vfo_reset(sdrpp_cuda_frontend* fe, int vfo);                         /* RxVFO::reset, rx_vfo.h:79-87 */
/* info[0..5] as sdrpp_cuda_design_resampler, info[6] = channel filter taps (0 if bypassed),
 * info[7] = stage-1 decimation fused with the NCO, info[8] = stage-1 taps. */
SDRPP_API int sdrpp_cuda_vfo_info(sdrpp_cuda_frontend* fe, int vfo, int* info);

/* One IQ block through the whole path (the work of threads A..L of SURVEY 3.2 for one
 * stream.swap()): conversion -> [PowerDecimator] -> [DCBlocker] -> [Conjugate] -> device ring ->
 * spectrum frames that complete in this block + every VFO + demod front ends -> pinned host
 * results. Asynchronous: returns after enqueueing; results are valid after _wait().
 * `in` is host memory (pinned memory from sdrpp_cuda_host_alloc avoids a staging copy). */
SDRPP_API int sdrpp_cuda_frontend_submit(sdrpp_cuda_frontend* fe, int fmt, const void* in, int count);
/* Same, with `in` already in device memory on this GPU (e.g. the target of an NCCL broadcast). */
SDRPP_API int sdrpp_cuda_frontend_submit_device(sdrpp_cuda_frontend* fe, int fmt, const void* dev_in, int count);
/* Same as _submit for one SDR++ server wire packet (SampleStreamDecompressor::run feeding the
 * front end, source_modules/sdrpp_server_source): the 8-byte header is read on the host, the payload
 * is converted on the device inside the ingest kernel. Returns the number of samples submitted
 * (0 = nothing submitted: unknown sample type or empty payload), < 0 on error. */
SDRPP_API int sdrpp_cuda_frontend_submit_pcm(sdrpp_cuda_frontend* fe, const void* packet, int nbytes);
/* Block until the OLDEST block not yet waited for has its results on the host. Up to five blocks may be in flight
 * (submit x5, wait, submit, wait, ...): with blocks submitted ahead, the host-to-device copy of a block never waits for an
 * earlier block's results to reach the host, and the end-to-end rate is no longer bound by one block's latency
 * (copy in + kernels + copy out) divided by the blocks in flight. A sixth submit blocks until the oldest is done. */
SDRPP_API int sdrpp_cuda_frontend_wait(sdrpp_cuda_frontend* fe);
/* Block until the host-to-device copy of the LAST submitted block has left the caller's buffer, i.e. until that buffer
 * may be reused (dsp::stream::flush() of the input stream, dsp/stream.h:88-96) -- long before the block's results exist.
 * Immediate for pageable memory (staged inside submit) and for device sources. */
SDRPP_API int sdrpp_cuda_frontend_wait_input(sdrpp_cuda_frontend* fe);
/* Blocks submitted and not yet waited for (0..5). */
SDRPP_API int sdrpp_cuda_frontend_pending(sdrpp_cuda_frontend* fe);
/* Wait for everything in flight and DISCARD the results not yet waited for: the next wait() returns the next submit.
 * (Control calls -- setters, VFO create/destroy -- never discard: blocks submitted before them are still handed back by
 * wait() in order, with the row layout they were computed with.)
 * Threading: one thread may submit (submit*, wait_input) while another waits and reads results (wait, vfo_output,
 * fft_rows, ...); all other calls must be serialised by the caller against both. */
SDRPP_API int sdrpp_cuda_frontend_drain(sdrpp_cuda_frontend* fe);
/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, the VFO set sharded across them (no VFO needs another's data), every IQ block
 * broadcast from the ingest rank over NVLink. Replaces the fan-out of dsp::routing::Splitter::run
 * (dsp/routing/splitter.h:46-60: memcpy + swap per consumer). The library owns the NCCL communicator and issues ONE
 * ncclBroadcast of the raw, still packed samples per block on a stream of its own; conversion and pre-processing then
 * run redundantly on every GPU, so every rank sees bit-identical samples. NCCL is loaded at run time (libnccl.so.2).
 *
 *   rank 0:      sdrpp_cuda_comm_unique_id(id)  -> hand the 128 bytes to the other processes (file, socket, MPI ...)
 *   every rank:  c = sdrpp_cuda_comm_create(id, rank, nranks, device);  sdrpp_cuda_frontend_set_comm(fe, c, root)
 *   per block:   root:   sdrpp_cuda_frontend_submit(fe, fmt, host_block, count)   (or _submit_device)
 *                others: sdrpp_cuda_frontend_submit_shared(fe, fmt, count)        (same fmt/count, same order)
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrpp_cuda_comm sdrpp_cuda_comm;
SDRPP_API int sdrpp_cuda_comm_unique_id(void* id128);
SDRPP_API sdrpp_cuda_comm* sdrpp_cuda_comm_create(const void* id128, int rank, int nranks, int device);
SDRPP_API int sdrpp_cuda_comm_destroy(sdrpp_cuda_comm* c);
/* rank / size / NCCL version code / broadcasts issued and bytes broadcast since creation (any pointer may be NULL) */
SDRPP_API int sdrpp_cuda_comm_info(sdrpp_cuda_comm* c, int* rank, int* nranks, int* nccl_version, long long* broadcasts, long long* bytes);
/* Attach a front end to a communicator (NULL detaches). From then on every submit on `root` broadcasts the block and
 * every other rank must mirror it with _submit_shared. */
SDRPP_API int sdrpp_cuda_frontend_set_comm(sdrpp_cuda_frontend* fe, sdrpp_cuda_comm* c, int root);
/* Non-root ranks: take part in the broadcast of the block the root is submitting (count samples of format fmt) and run
 * this rank's share of the path on it. */
SDRPP_API int sdrpp_cuda_frontend_submit_shared(sdrpp_cuda_frontend* fe, int fmt, int count);

/* Skip the device->host copies of results (kernel-only timing); default 1 = copy. */
SDRPP_API int sdrpp_cuda_frontend_set_readback(sdrpp_cuda_frontend* fe, int enabled);

/* Results of the last waited block. Pointers are into pinned host memory owned by the front end
 * and stay valid until the fifth submit after the block's own (five result sets rotate). */
/* RxVFO::out for this block: returns the output sample count; *iq -> cf32[count];
 * *demod -> float[count] (NULL when demod == NONE). */
SDRPP_API int sdrpp_cuda_vfo_output(sdrpp_cuda_frontend* fe, int vfo, const sdrpp_cf32** iq, const float** demod);
/* Post-detector stages of the demodulator behind a VFO, run on the device after the front end:
 *   QUADRATURE: dsp::demod::FM<float>  -- optional low-pass FIR lowPass(bw/2, bw/20, outSR) (dsp/demod/fm.h:86-103,117-145)
 *   AM:         dsp::demod::AM<float>  -- [carrier AGC] -> magnitude -> DC block -> [audio AGC] -> low-pass (dsp/demod/am.h:27-44,114-146)
 *   USB/LSB/DSB: dsp::demod::SSB<float> -- AGC on the real output (dsp/demod/ssb.h:21-36,90-101)
 *   QUADRATURE with wfm = 1: dsp::demod::BroadcastFM (dsp/demod/broadcast_fm.h:35-65,147-214), the WFM stereo decoder of the
 *               radio module (decoder_modules/radio/src/demodulators/wfm.h): 19 kHz pilot band-pass -> PLL -> L-R down-conversion
 *               -> L/R matrix -> 15 kHz low-pass; deviation = bandwidth / 2. Stereo output: sdrpp_cuda_vfo_audio_stereo. With
 *               wfm_rds = 1 the decoder's RDS side output (rdsOut, broadcast_fm.h:50-51,168-175,188-198) is produced as well:
 *               (mpx, 0) translated by -57 kHz and resampled to 5 kS/s; read it with sdrpp_cuda_vfo_rds.
 * with dsp::loop::AGC (dsp/loop/agc.h:87-147: setPoint 1, maxGain 10e6, maxOutputAmp 10, initGain INFINITY) and
 * dsp::correction::DCBlocker<float> (dsp/correction/dc_blocker.h:54-60). The radio module passes attack/decay/rate
 * already divided by the IF sample rate (decoder_modules/radio/src/demodulators/am.h:38, usb.h:40). */
typedef struct {
    int enabled;           /* 0: front end only (sdrpp_cuda_vfo_audio is then an error) */
    int fm_lowpass;        /* FM: _lowPass */
    int am_agc_mode;       /* AM: dsp::demod::AM::AGCMode 0 OFF, 1 CARRIER, 2 AUDIO */
    int ssb_agc;           /* SSB: agcEnabled */
    double agc_attack;     /* per-sample attack coefficient */
    double agc_decay;      /* per-sample decay coefficient */
    double dc_block_rate;  /* AM: dcBlockRate */
    float agc_gain;        /* > 0: setAGCGain(agc_gain) after init (the fixed gain when the audio AGC is off) */
    int wfm;               /* QUADRATURE only: 1 = dsp::demod::BroadcastFM (stereo decoder) instead of dsp::demod::FM */
    int wfm_stereo;        /* BroadcastFM _stereo (setStereo); the low-pass switch is fm_lowpass (setLowPass) */
    int wfm_rds;           /* BroadcastFM _rdsOut (setRDSOut): the 5 kS/s complex RDS baseband beside the audio */
} sdrpp_cuda_post_cfg;
/* (Re)initialises the VFO's post-detector objects; also redone when the VFO's bandwidth or rate changes. */
SDRPP_API int sdrpp_cuda_vfo_set_post(sdrpp_cuda_frontend* fe, int vfo, const sdrpp_cuda_post_cfg* cfg);
/* Demodulated audio of the last waited block: returns the sample count; *audio -> float[count]. */
SDRPP_API int sdrpp_cuda_vfo_audio(sdrpp_cuda_frontend* fe, int vfo, const float** audio);
/* Stereo demodulators (wfm = 1): left and right channel rows of the last waited block (dsp::stereo_t de-interleaved);
 * for mono demodulators *right = *left. Returns the sample count. */
SDRPP_API int sdrpp_cuda_vfo_audio_stereo(sdrpp_cuda_frontend* fe, int vfo, const float** left, const float** right);
/* BroadcastFM::rdsOut (dsp/demod/broadcast_fm.h:229) of the last waited block: returns the sample count (the reference swaps
 * the stream only when it is non-zero); *rds -> cf32[count] at 5 kS/s. SDRPP_ERR_STATE unless wfm = wfm_rds = 1. */
SDRPP_API int sdrpp_cuda_vfo_rds(sdrpp_cuda_frontend* fe, int vfo, const sdrpp_cf32** rds);

/* Radio IF chain between the VFO output and the demodulator front end: the order is the radio
 * module's, NoiseBlanker -> Squelch -> FMIF (decoder_modules/radio/src/radio_module.h:73-78). With a block enabled the VFO's
 * demod (and audio) results are computed from the chain's output; the iq result stays the raw VFO output (what the
 * reference exposes as vfo->output).
 *   dsp::noise_reduction::NoiseBlanker::init(in, rate, level) / process (core/src/dsp/noise_reduction/noise_blanker.h:12-17,39-59)
 *   dsp::noise_reduction::Squelch::init(in, level) / process (core/src/dsp/noise_reduction/squelch.h:19-26,34-64): mean
 *   magnitude of the block in dB against `level`, 1 dB hysteresis, 10 blocks above the level before unmuting. The
 *   reference's block counter is a function-local static shared by every Squelch in the process; here it is per VFO.
 *   dsp::noise_reduction::FMIF::init(in, bins) / process (core/src/dsp/noise_reduction/fm_if.h:20-24,45-74): per sample,
 *   Nuttall-windowed DFT of the last `bins` samples, strongest bin only, element bins/2 of the backward DFT. */
typedef struct {
    int nb_enabled;        /* ifChain.enableBlock(&nb) */
    double nb_rate;        /* radio: 500.0 / ifSamplerate (radio_module.h:428) */
    double nb_level;       /* amplitude ratio above the running mean at which a sample is scaled down */
    int squelch_enabled;   /* ifChain.enableBlock(&squelch) */
    double squelch_level;  /* dB */
    int fmif_bins;         /* 0: off; else FMIF::init(in, bins) / setBins, 2..64 (radio presets: 9, 15, 31, 32; radio_module.h:30-35) */
} sdrpp_cuda_if_cfg;
/* Creates the chain's objects on first use (NoiseBlanker amp = 1, Squelch unmuted); later calls change rate/level/enable
 * and keep the state, like setRate/setLevel/enableBlock. Needs the VFO's block output to fit the tail kernel's staging
 * area (<= 2176 samples per block), SDRPP_ERR_ARG otherwise. Changing fmif_bins clears the FMIF history (setBins). */
SDRPP_API int sdrpp_cuda_vfo_set_if_chain(sdrpp_cuda_frontend* fe, int vfo, const sdrpp_cuda_if_cfg* cfg);
/* Squelch state after the last submitted block: *muted, *level_db = 20*log10(mean magnitude) of that block. */
SDRPP_API int sdrpp_cuda_vfo_squelch_state(sdrpp_cuda_frontend* fe, int vfo, int* muted, float* level_db);

/* Spectrum rows completed in this block (each fft_size floats, the buffer handed to
 * acquireFFTBuffer/releaseFFTBuffer in the reference): returns the row count. */
SDRPP_API int sdrpp_cuda_fft_rows(sdrpp_cuda_frontend* fe, const float** rows);
/* Zoomed rows on the device: every spectrum row is also reduced to outSize pixels for the given view
 * (outSize = 0 disables). keep_raw = 0 stops the device->host copy of the full rows (4 B x fft_size per row),
 * leaving 4 B x outSize per row -- what the waterfall widget draws and the scanner module reads. */
SDRPP_API int sdrpp_cuda_frontend_set_fft_zoom(sdrpp_cuda_frontend* fe, double viewOffset, double viewBandwidth,
                                               double wholeBandwidth, int outSize, int keep_raw);
/* Zoomed rows completed in the last waited block (each outSize floats): returns the row count. */
SDRPP_API int sdrpp_cuda_fft_zoomed_rows(sdrpp_cuda_frontend* fe, const float** rows);
/* The waterfall's per-line display state on the ZOOMED row (WaterFall::pushFFT, gui/widgets/waterfall.cpp:918-956), kept on
 * the device: FFT smoothing latest = speed*latest + (1-speed)*smoothingBuf (setFFTSmoothing / setFFTSmoothingSpeed,
 * waterfall.cpp:1183-1211; the zoomed rows handed back are then the smoothed ones) and peak hold hold[i] = max(latest[i],
 * hold[i] - holdSpeed) for i >= 1 (setFFTHold / setFFTHoldSpeed, waterfall.cpp:1169-1181). Needs set_fft_zoom. */
SDRPP_API int sdrpp_cuda_frontend_set_fft_display(sdrpp_cuda_frontend* fe, int smoothing, float smoothingSpeed, int hold, float holdSpeed);
/* The peak-hold row after the last waited block: returns its width (0 when hold is off or the block completed no row). */
SDRPP_API int sdrpp_cuda_fft_hold_row(sdrpp_cuda_frontend* fe, const float** row);
/* WaterFall::calculateVFOSignalInfo (gui/widgets/waterfall.cpp:563-603) for this VFO on every RAW spectrum row, on the
 * device: strength = largest bin inside the VFO's bandwidth, snr = strength - mean of the half-bandwidth shoulders on
 * both sides (what the scanner module reads, misc_modules/scanner/src/main.cpp:164). centerOffset and bandwidth are the
 * VFO's own, wholeBandwidth the effective sample rate. */
SDRPP_API int sdrpp_cuda_vfo_set_signal_info(sdrpp_cuda_frontend* fe, int vfo, int enabled);
/* setSNRSmoothing / setSNRSmoothingSpeed (waterfall.cpp:1213-1220): snr = (1-speed)*snr + speed*new, row by row. */
SDRPP_API int sdrpp_cuda_frontend_set_snr_smoothing(sdrpp_cuda_frontend* fe, int enabled, float speed);
/* Per-row results of the last waited block (up to cap rows): returns the row count. level_max (optional) = the maximum of
 * the last ten levels (selectedVFO_LevelMax, waterfall.cpp:937-948). */
SDRPP_API int sdrpp_cuda_vfo_signal_info(sdrpp_cuda_frontend* fe, int vfo, float* strength, float* snr, float* level_max, int cap);
/* One-shot form on a host row of N floats, for nvfo (centerOffset, bandwidth) pairs. */
SDRPP_API int sdrpp_cuda_signal_info(int N, const float* row, int nvfo, const double* centerOffset, const double* bandwidth,
                                     double wholeBandwidth, float* strength, float* snr);
/* The post-preprocessing IQ block as the Splitter would hand it to bound streams
 * (IQFrontEnd::bindIQStream, signal_path/iq_frontend.cpp:114-116; recorder tap). Copies up to cap
 * samples of the last block to `out` (device->host); returns the count. */
SDRPP_API int sdrpp_cuda_frontend_read_iq(sdrpp_cuda_frontend* fe, sdrpp_cf32* out, int cap);

/* Kernel launch counter (kernels of this library launched since creation) and the CUDA stream
 * (cudaStream_t) the front end enqueues on, for device-side timing with events. */
SDRPP_API long long sdrpp_cuda_frontend_launches(sdrpp_cuda_frontend* fe);
SDRPP_API void* sdrpp_cuda_frontend_stream(sdrpp_cuda_frontend* fe);
/* Make that stream wait for everything enqueued so far on the front end's other streams (spectrum, tail, result copies,
 * broadcast): an event recorded on it afterwards covers the whole of the blocks submitted up to here. No host sync. */
SDRPP_API int sdrpp_cuda_frontend_join_streams(sdrpp_cuda_frontend* fe);
/* Device-time of the kernels of the last waited block, by kernel family, measured with CUDA
 * events on the front end's stream when profiling is enabled (ms). idx: 0 ingest/preproc,
 * 1 spectrum, 2 channelizer stage 1, 3 channelizer tail. */
SDRPP_API int sdrpp_cuda_frontend_set_profiling(sdrpp_cuda_frontend* fe, int enabled);
SDRPP_API float sdrpp_cuda_frontend_kernel_ms(sdrpp_cuda_frontend* fe, int idx);
/* Channelizer stage 1 (FrequencyXlator + first DecimatingFIR, frequency_xlator.h:43-50 + decimating_fir.h:45-68)
 * has two device implementations with the same results contract: mode 0 (default) runs it on the tensor cores
 * (tcgen05, fp16 hi/lo split operands, fp32 accumulate) for first-stage decimations of 32 and 64 and on the FP32
 * FMA kernel otherwise; mode 1 uses the FP32 FMA kernel only. The environment variable SDRPP_S1_MODE=fp32 sets
 * mode 1 at creation. _stage1_tensor_launches counts the tensor-core stage-1 launches since creation. */
SDRPP_API int sdrpp_cuda_frontend_set_stage1_mode(sdrpp_cuda_frontend* fe, int mode);
/* CUDA graphs. A block is planned into a command list; the kernels of its stage-1 part (descriptor upload, ingest, spectrum,
 * fp16 split, stage 1) and of its tail part are replayed as one instantiated graph each whenever exactly the same command
 * sequence has run before (everything that changes per block travels in a device-resident descriptor, so graphs are
 * never updated). A sequence is instantiated the second time it is seen; a steady stream settles on about 30 graphs
 * within ~150 blocks. _set_graphs(0) runs every block command by command (also: environment SDRPP_GRAPHS=0); results are
 * bit-identical either way. _graph_stats: out[0] tagged runs replayed as a graph, out[1] graphs instantiated, out[2] tagged
 * runs executed command by command, out[3] microseconds spent instantiating. */
SDRPP_API int sdrpp_cuda_frontend_set_graphs(sdrpp_cuda_frontend* fe, int enabled);
SDRPP_API int sdrpp_cuda_frontend_graph_stats(sdrpp_cuda_frontend* fe, long long* out4);
SDRPP_API long long sdrpp_cuda_frontend_stage1_tensor_launches(sdrpp_cuda_frontend* fe);

#ifdef __cplusplus
}
#endif
#endif /* SDRPP_CUDA_H */
