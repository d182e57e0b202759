// Host-side design maths (see design.h). Pure C++, no CUDA.
#include "design.h"
#include <cmath>
#include <cstring>
#include <mutex>

namespace sdrpp {

// ---------------------------------------------------------------------------------------------
// Windows: dsp/window/cosine.h:7-16 + coefficient headers
// ---------------------------------------------------------------------------------------------
static double cosine_sum(double n, double N, const double* c, int cnt) {
    double win = 0.0, sign = 1.0;
    for (int i = 0; i < cnt; i++) {
        win += sign * c[i] * cos((double)i * 2.0 * kPi * n / N);
        sign = -sign;
    }
    return win;
}

static const double kRect[] = { 1.0 };
static const double kHamming[] = { 0.53836, 0.46164 };                       // hamming.h:6
static const double kHann[] = { 0.5, 0.5 };                                  // hann.h:6
static const double kBlackman[] = { 0.42, 0.5, 0.08 };                       // blackman.h:6
static const double kNuttall[] = { 0.355768, 0.487396, 0.144232, 0.012604 }; // nuttall.h:6
static const double kBH4[] = { 0.35875, 0.48829, 0.14128, 0.01168 };         // blackman_harris4.h:6
static const double kBH7[] = { 0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803,
                               0.01081174209837, 0.00077658482522, 0.00001388721735 }; // blackman_harris7.h:22-30

static int window_coefs(int type, const double** c) {
    switch (type) { // enum order of dsp::window::windowType, window.h:28-36
    case 0: *c = kRect; return 1;
    case 1: *c = kHamming; return 2;
    case 2: *c = kHann; return 2;
    case 3: *c = kBlackman; return 3;
    case 4: *c = kNuttall; return 4;
    case 5: *c = kBH4; return 4;
    case 6: *c = kBH7; return 7;
    }
    return 0;
}

int design_window(int type, float* buf, int size, bool centered) {
    const double* c;
    const int cnt = window_coefs(type, &c);
    if (!cnt || size <= 0 || !buf) return -1;
    // evaluated in double, stored to float first, then summed back in double (window.h:43-52)
    for (int i = 0; i < size; i++) buf[i] = (float)cosine_sum((double)i, (double)size, c, cnt);
    double wscale = 0.0;
    for (int i = 0; i < size; i++) wscale += buf[i];
    wscale = 1.0 / wscale;
    if (!centered) {
        for (int i = 0; i < size; i++) buf[i] = (float)((double)buf[i] * wscale);
    } else {
        // (-,+,-,+...) alternation does the FFT shift (window.h:57-62); writes buf[size] when size is odd
        for (int i = 0; i < size; i += 2) {
            buf[i] = (float)((double)buf[i] * -wscale);
            buf[i + 1] = (float)((double)buf[i + 1] * wscale);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dsp::window::nuttall(n, N) (window/nuttall.h:5-8), as called by FMIF::initBuffers (noise_reduction/fm_if.h:111)
double window_nuttall(double n, double N) { return cosine_sum(n, N, kNuttall, 4); }

// Taps: dsp/taps/windowed_sinc.h:9-29 with window::nuttall, math/sinc.h, math/hz_to_rads.h
// ---------------------------------------------------------------------------------------------
int lowpass_tap_count(double transWidth, double sampleRate) {
    return (int)(3.8 * sampleRate / transWidth);
}

std::vector<float> design_lowpass(double cutoff, double transWidth, double sampleRate) {
    const int count = lowpass_tap_count(transWidth, sampleRate);
    std::vector<float> taps((size_t)(count > 0 ? count : 0));
    const double omega = 2.0 * kPi * (cutoff / sampleRate);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / kPi;
    for (int i = 0; i < count; i++) {
        const double t = (double)i - half + 0.5;
        const double x = t * omega;
        const double sinc = (x == 0.0) ? 1.0 : (sin(x) / x);
        taps[(size_t)i] = (float)(sinc * cosine_sum(t - half, (double)count, kNuttall, 4) * corr);
    }
    return taps;
}

// ---------------------------------------------------------------------------------------------
// PowerDecimator plans. The coefficient tables are numeric data of the reference that cannot be
// regenerated (dsp/multirate/decim/plans.h:17-22); they are embedded from
// sdrpp_b200/data/decim_plans.bin (tools/extract_decim_plans.py) at build time.
// ---------------------------------------------------------------------------------------------
static const uint32_t kPlanBlob[] = {
#include "decim_plans_blob.inc"
};

namespace {
struct PlanFir { uint32_t len, off; };
struct PlanEnt { uint32_t ratio, nstages; struct { uint32_t decim, fir; } st[4]; };
struct PlanTable {
    uint32_t nfirs = 0, nplans = 0, pool_len = 0;
    const PlanFir* firs = nullptr;
    const PlanEnt* plans = nullptr;
    const float* pool = nullptr;
    bool ok = false;
    PlanTable() {
        const uint32_t* w = kPlanBlob;
        if (sizeof(kPlanBlob) < 20 || w[0] != 0x50445053u || w[1] != 1) return;
        nfirs = w[2]; nplans = w[3]; pool_len = w[4];
        firs = reinterpret_cast<const PlanFir*>(w + 5);
        plans = reinterpret_cast<const PlanEnt*>(w + 5 + 2 * nfirs);
        pool = reinterpret_cast<const float*>(w + 5 + 2 * nfirs + 10 * nplans);
        ok = (5 + 2 * nfirs + 10 * nplans + pool_len) * 4 == sizeof(kPlanBlob);
    }
};
const PlanTable& plan_table() { static PlanTable t; return t; }
} // namespace

std::vector<DecimStage> decim_plan(int ratio) {
    std::vector<DecimStage> out;
    const PlanTable& t = plan_table();
    if (!t.ok) return out;
    for (uint32_t i = 0; i < t.nplans; i++) {
        if ((int)t.plans[i].ratio != ratio) continue;
        for (uint32_t s = 0; s < t.plans[i].nstages; s++) {
            const PlanFir& f = t.firs[t.plans[i].st[s].fir];
            out.push_back({ (int)t.plans[i].st[s].decim, (int)f.len, t.pool + f.off, (int)t.plans[i].st[s].fir });
        }
        break;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// RationalResampler plan: dsp/multirate/rational_resampler.h:121-167
// ---------------------------------------------------------------------------------------------
static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a < 0 ? -a : a; }

ResamplerPlan design_resampler(double inSR, double outSR) {
    ResamplerPlan p;
    const int maxRatio = 1 << 13; // PowerDecimator::getMaxRatio, power_decimator.h:28-30
    int predecPower = (int)floor(log2(inSR / outSR));
    if (predecPower > maxRatio) predecPower = maxRatio; // :123 clamps the exponent against 8192
    int predecRatio = (predecPower >= 0 && predecPower < 31) ? (1 << predecPower) : ((predecPower < 0) ? 0 : maxRatio);
    if (predecRatio > maxRatio) predecRatio = maxRatio; // :124
    const bool useDecim = (inSR > outSR && predecPower > 0);
    double intSR = inSR;
    if (useDecim) intSR = inSR / (double)predecRatio;
    const int IntSR = (int)round(intSR), OutSR = (int)round(outSR);
    const int g = gcd_i(IntSR, OutSR);
    const int interp = OutSR / g, decim = IntSR / g;
    p.predec = useDecim ? predecRatio : 1;
    if (interp == decim) {
        p.mode = useDecim ? 1 : 3;
        return p;
    }
    const double tapSR = intSR * (double)interp;
    const double tapBW = (inSR < outSR ? inSR : outSR) / 2.0;
    const double tapTW = tapBW * 0.1;
    p.taps = design_lowpass(tapBW, tapTW, tapSR);
    for (float& t : p.taps) t *= (float)interp; // :160
    p.mode = useDecim ? 0 : 2;
    p.interp = interp; p.decim = decim;
    p.tpp = ((int)p.taps.size() + interp - 1) / interp;
    return p;
}

std::vector<float> build_polyphase_bank(const std::vector<float>& taps, int interp, int* tpp_out) {
    const int ntaps = (int)taps.size();
    const int tpp = (ntaps + interp - 1) / interp; // polyphase_bank.h:24
    std::vector<float> bank((size_t)interp * (size_t)tpp, 0.0f);
    const int tot = interp * tpp;
    for (int i = 0; i < tot; i++) // polyphase_bank.h:31-34
        bank[(size_t)((interp - 1) - (i % interp)) * (size_t)tpp + (size_t)(i / interp)] = (i < ntaps) ? taps[(size_t)i] : 0.0f;
    if (tpp_out) *tpp_out = tpp;
    return bank;
}

void reshape_params(double sampleRate, int size, double rate, int* skip, int* nz) {
    const int interval = (int)round(sampleRate / rate);
    *nz = interval < size ? interval : size;
    *skip = interval - *nz;
}

bool zoom_indices(double viewOffset, double viewBandwidth, double wholeBandwidth, int fftSize, int outSize, std::vector<int>* idx) {
    // the running sum f0 += factor is kept in double and rounded through float exactly as doZoom does, so the
    // bin boundaries are the reference's own
    const double offsetRatio = viewOffset / (wholeBandwidth / 2.0);
    double width = (viewBandwidth / wholeBandwidth) * fftSize;
    double offset = (((double)fftSize / 2.0) * (offsetRatio + 1)) - (width / 2);
    if (offset < 0) offset = 0;
    if (width > fftSize - offset) width = fftSize - offset;
    const double factor = width / outSize;
    idx->assign((size_t)outSize + 1, 0);
    double f0 = offset;
    if (factor <= 1.0) {
        for (int i = 0; i < outSize; i++) { (*idx)[(size_t)i] = (int)roundf((float)f0); f0 = f0 + factor; }
        (*idx)[(size_t)outSize] = -1;
        return false;
    }
    int i0 = (int)roundf((float)f0);
    for (int i = 0; i < outSize; i++) {
        const double f1 = f0 + factor;
        (*idx)[(size_t)i] = i0;
        i0 = (int)roundf((float)f1);
        f0 = f1;
    }
    (*idx)[(size_t)outSize] = i0;
    return true;
}

// dsp::taps::bandPass<complex_t>(bandStart, bandStop, transWidth, sampleRate, oddTapCount) (taps/band_pass.h:10-25 over
// taps/windowed_sinc.h:9-29), operation for operation: the sinc in double, the window = phasor(-offsetOmega * (float)n)
// (cosf / sinf of a float) times the Nuttall window cast to float, the correction factor cast to float.
std::vector<float> design_bandpass_complex(double bandStart, double bandStop, double transWidth, double sampleRate, bool oddTapCount) {
    const float offsetOmega = (float)(2.0 * kPi * (((bandStart + bandStop) / 2.0) / sampleRate));
    int count = (int)(3.8 * sampleRate / transWidth);
    if (oddTapCount && !(count % 2)) count++;
    const double omega = 2.0 * kPi * (((bandStop - bandStart) / 2.0) / sampleRate);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / kPi;
    std::vector<float> taps((size_t)2 * (size_t)std::max(count, 0));
    for (int i = 0; i < count; i++) {
        const double t = (double)i - half + 0.5;
        const double x = t * omega;
        const float sc = (float)((x == 0.0) ? 1.0 : (sin(x) / x));
        const double n = t - half;
        const float ph = -offsetOmega * (float)n;
        const float wn = (float)window_nuttall(n, (double)count);
        const float wre = cosf(ph) * wn, wim = sinf(ph) * wn;          // complex_t * double: each part times (float)b
        const float re = (sc * wre) - (0.0f * wim), im = (0.0f * wre) + (sc * wim); // complex_t * complex_t, types.h:23-25
        taps[2 * (size_t)i] = re * (float)corr;
        taps[2 * (size_t)i + 1] = im * (float)corr;
    }
    return taps;
}

// PhaseControlLoop<float>::criticallyDamped (loop/phase_control_loop.h:31-36): the float T makes every intermediate a float
// only where the reference's expression does (sqrt(2.0)/2.0 is a double rounded into the float dampningFactor).
void pll_critically_damped(float bandwidth, float* alpha, float* beta) {
    const float dampningFactor = (float)(sqrt(2.0) / 2.0);
    const float denominator = (float)(1.0 + 2.0 * dampningFactor * bandwidth + bandwidth * bandwidth);
    *alpha = (4 * dampningFactor * bandwidth) / denominator;
    *beta = (4 * bandwidth * bandwidth) / denominator;
}

void signal_info_bins(double centerOffset, double bandwidth, double wholeBandwidth, int rawFFTSize, int out[4]) {
    // gui/widgets/waterfall.cpp:567-574, operation for operation (double arithmetic, truncation to int, clamp to [0, size])
    const double f[4] = { centerOffset - bandwidth, centerOffset - (bandwidth / 2.0), centerOffset + (bandwidth / 2.0), centerOffset + bandwidth };
    for (int i = 0; i < 4; i++) {
        const int v = (int)(((f[i] / (wholeBandwidth / 2.0)) * (double)(rawFFTSize / 2)) + (rawFFTSize / 2));
        out[i] = v < 0 ? 0 : (v > rawFFTSize ? rawFFTSize : v);
    }
}

void xlator_increment(double offsetHz, double sampleRate, float* inc_re, float* inc_im, double* turns_eff) {
    const double w = 2.0 * kPi * (offsetHz / sampleRate); // math/hz_to_rads.h:6-8
    const float re = (float)cos(w), im = (float)sin(w);   // frequency_xlator.h:17-19
    if (inc_re) *inc_re = re;
    if (inc_im) *inc_im = im;
    // the recurrence phase *= inc advances by arg(inc) per sample (w_eff)
    if (turns_eff) *turns_eff = atan2((double)im, (double)re) / (2.0 * kPi);
}

} // namespace sdrpp
