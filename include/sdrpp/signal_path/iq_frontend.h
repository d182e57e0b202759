// IQFrontEnd -- drop-in replacement of core/src/signal_path/iq_frontend.h (reference: iq_frontend.h:14-104). Same
// public interface, member for member; the implementation (sdrpp_b200/host/iq_frontend.cpp, which replaces
// core/src/signal_path/iq_frontend.cpp) drives the CUDA library through its C ABI (sdrpp_cuda.h) instead of the
// reference's thread-per-block chain: the pre-processing chain, the splitter fan-out, the spectrum branch and every
// bound RxVFO run as one stream-ordered launch sequence per IQ block on the GPU.
//
// Host threads per front end: `ingest` reads the input stream and submits blocks (up to four ahead); `deliver` waits for
// finished blocks in order and, with a few helpers, copies + swaps VFO blocks into their RxVFO::out streams; `spectrum`
// hands finished rows to the acquire/release pair (the reference's FFT sink thread). The H2D copy and the kernels of
// block i+1 run while block i is being delivered.
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
// the same dsp/ headers the reference's iq_frontend.h includes: modules rely on them arriving through signal_path.h
// (e.g. decoder_modules/radio/src/demodulators/wfm.h uses dsp::sink::Handler and dsp::buffer::Reshaper that way)
#include "../dsp/buffer/frame_buffer.h"
#include "../dsp/buffer/reshaper.h"
#include "../dsp/multirate/power_decimator.h"
#include "../dsp/correction/dc_blocker.h"
#include "../dsp/chain.h"
#include "../dsp/routing/splitter.h"
#include "../dsp/channel/rx_vfo.h"
#include "../dsp/sink/handler_sink.h"
#include "../dsp/math/conjugate.h"
#include "../dsp/window/window.h"
#include <sdrpp_cuda.h>

class IQFrontEnd {
public:
    ~IQFrontEnd();

    void init(dsp::stream<dsp::complex_t>* in, double sampleRate, bool buffering, int decimRatio, bool dcBlocking, int fftSize, double fftRate, dsp::window::windowType fftWindow, float* (*acquireFFTBuffer)(void* ctx), void (*releaseFFTBuffer)(void* ctx), void* fftCtx);

    void updateFFTSize();

    void setInput(dsp::stream<dsp::complex_t>* in);
    void setSampleRate(double sampleRate);
    inline double getSampleRate() { return _sampleRate / _decimRatio; }

    void setBuffering(bool enabled);
    void setDecimation(int ratio);
    void setInvertIQ(bool enabled);
    void setDCBlocking(bool enabled);

    void bindIQStream(dsp::stream<dsp::complex_t>* stream);
    void unbindIQStream(dsp::stream<dsp::complex_t>* stream);

    dsp::channel::RxVFO* addVFO(std::string name, double sampleRate, double bandwidth, double offset);
    void removeVFO(std::string name);

    void setFFTSize(int size);
    void setFFTRate(double rate);
    void setFFTWindow(dsp::window::windowType fftWindow);

    void flushInputBuffer();

    void start();
    void stop();

    double getEffectiveSamplerate();

    // Extension (not in the reference): the engine handle, for sources that hand raw device samples over instead of
    // converting on the CPU (sdrpp_cuda_frontend_submit with SDRPP_FMT_U8_RTL ...), and the
    // blocks delivered so far.
    sdrpp_cuda_frontend* engine() { return fe; }
    long long blocksDelivered();

protected:
    void ingestLoop();
    void deliverLoop();
    void deliverBlock();
    void spectrumLoop();
    void helperLoop(int k);
    void deliverShard(int k);
    void updateFFTPath(bool updateWaterfall = false);

    static constexpr int kMaxAhead = 4; // blocks submitted and not yet delivered (the engine keeps five result sets)

    dsp::stream<dsp::complex_t>* _in = nullptr;
    sdrpp_cuda_frontend* fe = nullptr;
    std::recursive_mutex mtx;   // control surface + every engine call except the blocking waits
    std::mutex swapMtx;         // held by the deliver thread while it swaps blocks into the output streams
    std::thread ingestThread, deliverThread, spectrumThread;
    std::vector<std::thread> helpers;
    std::mutex flowMtx;
    std::condition_variable flowCv;
    long long submitted = 0, delivered = 0;
    bool running = false, stopping = false;
    // rows of the block being delivered, handed to the spectrum thread (at most one block behind the deliver thread, so
    // the engine's result set is still intact when it copies)
    std::deque<std::pair<const float*, int>> rowsQueue;
    long long rowsPosted = 0, rowsDone = 0;
    // VFO blocks of the block being delivered, sharded over the deliver thread + helpers
    struct OutItem { dsp::channel::RxVFO* vfo; const sdrpp_cf32* iq; int n; };
    std::vector<OutItem> outItems;
    std::vector<dsp::stream<dsp::complex_t>*> outTaps;
    int outRaw = 0;
    long long shardGen = 0;
    int shardsLeft = 0;

    // VFOs and raw IQ taps
    std::map<std::string, dsp::channel::RxVFO*> vfos;
    std::vector<dsp::stream<dsp::complex_t>*> bound;

    // Parameters
    double _sampleRate = 0;
    double _decimRatio = 1;
    int _fftSize = 0;
    double _fftRate = 0;
    dsp::window::windowType _fftWindow = dsp::window::windowType::NUTTALL;
    float* (*_acquireFFTBuffer)(void* ctx) = nullptr;
    void (*_releaseFFTBuffer)(void* ctx) = nullptr;
    void* _fftCtx = nullptr;

    double effectiveSr = 0;

    bool _init = false;
};
