// Host-side design maths of the signal path: window tables, windowed-sinc taps, the
// RationalResampler plan and the PowerDecimator stage tables. All double precision, stored to
// float exactly as the reference does, because these define the filter shapes and the integer
// index arithmetic that results must reproduce.
#pragma once
#include <cstdint>
#include <vector>

namespace sdrpp {

constexpr double kPi = 3.14159265358979323846; // DB_M_PI, dsp/math/constants.h:2

// dsp::window::createWindow (dsp/window/window.h:38-64). buf needs size+1 floats.
int design_window(int type, float* buf, int size, bool centered);

// dsp::taps::estimateTapCount / lowPass (dsp/taps/estimate_tap_count.h:4-6, low_pass.h:7-11)
int lowpass_tap_count(double transWidth, double sampleRate);
std::vector<float> design_lowpass(double cutoff, double transWidth, double sampleRate);

// One stage of a PowerDecimator plan (dsp/multirate/decim/plans.h:36-140)
struct DecimStage {
    int decimation;
    int ntaps;
    const float* taps;
    int fir_id; // index into the distinct FIR table (13 filters)
};
// Stage list for ratio = 2^k, k in 1..13; empty if the ratio is invalid.
std::vector<DecimStage> decim_plan(int ratio);

// dsp::multirate::RationalResampler::reconfigure (dsp/multirate/rational_resampler.h:121-167)
struct ResamplerPlan {
    int mode = 3;   // 0 BOTH, 1 DECIM_ONLY, 2 RESAMP_ONLY, 3 NONE
    int predec = 1; // power-of-two pre-decimation ratio
    int interp = 1, decim = 1;
    int tpp = 0;               // taps per phase
    std::vector<float> taps;   // polyphase prototype, already scaled by interp
};
ResamplerPlan design_resampler(double inSR, double outSR);

// dsp::multirate::PolyphaseBank (dsp/multirate/polyphase_bank.h:15-48): bank[phase][j], tpp each
std::vector<float> build_polyphase_bank(const std::vector<float>& taps, int interp, int* tpp);

// dsp::window::nuttall (window/nuttall.h:5-8)
double window_nuttall(double n, double N);

// IQFrontEnd::genReshapeParams (signal_path/iq_frontend.h:56-60)
void reshape_params(double sampleRate, int size, double rate, int* skip, int* nz);

// fft_scaler (gui/widgets/fft_scaler.h:28-64): bin boundaries of the waterfall zoom / max-decimation.
// idx has outSize+1 entries; pixel i covers bins [idx[i], max(idx[i]+1, idx[i+1])) when ranged, bin idx[i] otherwise.
bool zoom_indices(double viewOffset, double viewBandwidth, double wholeBandwidth, int fftSize, int outSize, std::vector<int>* idx);

// dsp::taps::bandPass<complex_t> (taps/band_pass.h): interleaved (re, im) taps. PhaseControlLoop<float>::criticallyDamped.
std::vector<float> design_bandpass_complex(double bandStart, double bandStop, double transWidth, double sampleRate, bool oddTapCount);
void pll_critically_damped(float bandwidth, float* alpha, float* beta);

// Bin ranges of WaterFall::calculateVFOSignalInfo (gui/widgets/waterfall.cpp:567-574): out = (minSide, min, max, maxSide).
void signal_info_bins(double centerOffset, double bandwidth, double wholeBandwidth, int rawFFTSize, int out[4]);

// FrequencyXlator increment (dsp/channel/frequency_xlator.h:15-23): the fp32-quantised phasor
// (cos w, sin w) and the frequency it actually realises, in turns per sample.
void xlator_increment(double offsetHz, double sampleRate, float* inc_re, float* inc_im, double* turns_eff);

} // namespace sdrpp
