// IQFrontEnd -- host-side mirror of sigpath::iqFrontEnd (reference: core/src/signal_path/iq_frontend.h:14-104,
// iq_frontend.cpp:15-296). Same public methods; the pre-processing chain, the splitter fan-out, the spectrum
// branch and every bound RxVFO run as one stream-ordered launch sequence per IQ block on the GPU
// (sdrpp_cuda_frontend_*), driven by a single worker thread instead of a thread per block.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "../dsp/channel/rx_vfo.h"
#include "../dsp/window/window.h"
#include "../../sdrpp_cuda.h"

class IQFrontEnd {
public:
    ~IQFrontEnd() {
        if (!_init) { return; }
        stop();
        for (auto& [name, vfo] : vfos) { delete vfo; }
        if (fe) { sdrpp_cuda_frontend_destroy(fe); }
    }

    void init(dsp::stream<dsp::complex_t>* in, double sampleRate, bool buffering, int decimRatio, bool dcBlocking, int fftSize,
              double fftRate, dsp::window::windowType fftWindow, float* (*acquireFFTBuffer)(void* ctx),
              void (*releaseFFTBuffer)(void* ctx), void* fftCtx) {
        _in = in; _sampleRate = sampleRate; _decimRatio = decimRatio; _fftSize = fftSize; _fftRate = fftRate; _fftWindow = fftWindow;
        _acquireFFTBuffer = acquireFFTBuffer; _releaseFFTBuffer = releaseFFTBuffer; _fftCtx = fftCtx;
        (void)buffering; // the 32-deep SampleFrameBuffer is replaced by the library's pinned/device double buffering
        sdrpp_cuda_frontend_cfg cfg{};
        cfg.sample_rate = sampleRate; cfg.decim_ratio = decimRatio; cfg.dc_blocking = dcBlocking; cfg.invert_iq = 0;
        cfg.fft_size = fftSize; cfg.fft_rate = fftRate; cfg.fft_window = (int)fftWindow; cfg.max_block = STREAM_BUFFER_SIZE;
        fe = sdrpp_cuda_frontend_create(&cfg);
        if (!fe) { fprintf(stderr, "[IQFrontEnd] %s\n", sdrpp_cuda_last_error()); }
        effectiveSr = _sampleRate / _decimRatio;
        _init = true;
    }

    void setInput(dsp::stream<dsp::complex_t>* in) {
        const bool was = running;
        stop();
        _in = in;
        if (was) { start(); }
    }

    void setSampleRate(double sampleRate) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        _sampleRate = sampleRate;
        effectiveSr = _sampleRate / _decimRatio;
        if (fe) { sdrpp_cuda_frontend_set_sample_rate(fe, sampleRate); }
    }
    inline double getSampleRate() { return _sampleRate / _decimRatio; }
    void setBuffering(bool enabled) { (void)enabled; }
    void setDecimation(int ratio) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        _decimRatio = ratio;
        effectiveSr = _sampleRate / _decimRatio;
        if (fe) { sdrpp_cuda_frontend_set_decimation(fe, ratio); }
    }
    void setInvertIQ(bool enabled) { std::lock_guard<std::recursive_mutex> lck(mtx); if (fe) { sdrpp_cuda_frontend_set_invert_iq(fe, enabled); } }
    void setDCBlocking(bool enabled) { std::lock_guard<std::recursive_mutex> lck(mtx); if (fe) { sdrpp_cuda_frontend_set_dc_blocking(fe, enabled); } }

    // Raw (post-preprocessing) IQ taps, as the Splitter hands them out (recorder, iq_frontend.cpp:114-120)
    void bindIQStream(dsp::stream<dsp::complex_t>* stream) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        if (std::find(bound.begin(), bound.end(), stream) != bound.end()) { throw std::runtime_error("[IQFrontEnd] stream already bound"); }
        bound.push_back(stream);
    }
    void unbindIQStream(dsp::stream<dsp::complex_t>* stream) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        auto it = std::find(bound.begin(), bound.end(), stream);
        if (it == bound.end()) { throw std::runtime_error("[IQFrontEnd] stream not bound"); }
        bound.erase(it);
    }

    dsp::channel::RxVFO* addVFO(std::string name, double sampleRate, double bandwidth, double offset) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        if (vfos.find(name) != vfos.end()) {
            fprintf(stderr, "[IQFrontEnd] Tried to add VFO with existing name.\n");
            return NULL;
        }
        if (!fe) { return NULL; }
        const int id = sdrpp_cuda_vfo_create(fe, sampleRate, bandwidth, offset, SDRPP_DEMOD_NONE);
        if (id < 0) { fprintf(stderr, "[IQFrontEnd] %s\n", sdrpp_cuda_last_error()); return NULL; }
        auto* vfo = new dsp::channel::RxVFO();
        vfo->attach(fe, id, &mtx, effectiveSr, sampleRate, bandwidth, offset);
        vfos[name] = vfo;
        return vfo;
    }
    void removeVFO(std::string name) {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        auto it = vfos.find(name);
        if (it == vfos.end()) {
            fprintf(stderr, "[IQFrontEnd] Tried to remove a VFO that doesn't exist.\n");
            return;
        }
        it->second->out.stopWriter();
        sdrpp_cuda_vfo_destroy(fe, it->second->vfoId);
        delete it->second;
        vfos.erase(it);
    }

    void setFFTSize(int size) { std::lock_guard<std::recursive_mutex> lck(mtx); _fftSize = size; if (fe) { sdrpp_cuda_frontend_set_fft_size(fe, size); } }
    void setFFTRate(double rate) { std::lock_guard<std::recursive_mutex> lck(mtx); _fftRate = rate; if (fe) { sdrpp_cuda_frontend_set_fft_rate(fe, rate); } }
    void setFFTWindow(dsp::window::windowType w) { std::lock_guard<std::recursive_mutex> lck(mtx); _fftWindow = w; if (fe) { sdrpp_cuda_frontend_set_fft_window(fe, (int)w); } }
    void flushInputBuffer() {}

    void start() {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        if (running || !_init || !_in) { return; }
        running = true;
        worker = std::thread(&IQFrontEnd::workerLoop, this);
    }
    void stop() {
        {
            std::lock_guard<std::recursive_mutex> lck(mtx);
            if (!running) { return; }
            running = false;
        }
        if (_in) { _in->stopReader(); }
        for (auto& [name, vfo] : vfos) { vfo->out.stopWriter(); }
        for (auto* s : bound) { s->stopWriter(); }
        if (worker.joinable()) { worker.join(); }
        if (_in) { _in->clearReadStop(); }
        for (auto& [name, vfo] : vfos) { vfo->out.clearWriteStop(); }
        for (auto* s : bound) { s->clearWriteStop(); }
    }

    double getEffectiveSamplerate() { return effectiveSr; }
    sdrpp_cuda_frontend* engine() { return fe; }

protected:
    void workerLoop() {
        while (true) {
            const int count = _in->read();
            if (count < 0) { return; }
            std::unique_lock<std::recursive_mutex> lck(mtx);
            if (!fe || sdrpp_cuda_frontend_submit(fe, SDRPP_FMT_CF32, _in->readBuf, count) < 0 || sdrpp_cuda_frontend_wait(fe) < 0) {
                fprintf(stderr, "[IQFrontEnd] %s\n", sdrpp_cuda_last_error());
                _in->flush();
                continue;
            }
            _in->flush();
            // spectrum rows: acquire/release are always called as a pair, once per line (iq_frontend.cpp:239-248)
            const float* rows = nullptr;
            const int nrows = sdrpp_cuda_fft_rows(fe, &rows);
            for (int r = 0; r < nrows; r++) {
                float* dst = _acquireFFTBuffer ? _acquireFFTBuffer(_fftCtx) : nullptr;
                if (dst) { memcpy(dst, rows + (size_t)r * _fftSize, sizeof(float) * (size_t)_fftSize); }
                if (_releaseFFTBuffer) { _releaseFFTBuffer(_fftCtx); }
            }
            // VFO outputs into each RxVFO::out (swap blocks until the consumer flushed the previous block)
            std::vector<dsp::channel::RxVFO*> live;
            for (auto& [name, vfo] : vfos) { live.push_back(vfo); }
            std::vector<dsp::stream<dsp::complex_t>*> taps = bound;
            const sdrpp_cf32* iq = nullptr;
            for (auto* vfo : live) {
                const int n = sdrpp_cuda_vfo_output(fe, vfo->vfoId, &iq, nullptr);
                if (n > 0) { memcpy(vfo->out.writeBuf, iq, sizeof(dsp::complex_t) * (size_t)n); }
                vfo->pendingOut = n;
            }
            int nraw = 0;
            if (!taps.empty()) {
                nraw = sdrpp_cuda_frontend_read_iq(fe, (sdrpp_cf32*)taps[0]->writeBuf, STREAM_BUFFER_SIZE);
                for (size_t i = 1; i < taps.size(); i++) { memcpy(taps[i]->writeBuf, taps[0]->writeBuf, sizeof(dsp::complex_t) * (size_t)std::max(nraw, 0)); }
            }
            lck.unlock(); // never hold the control mutex while blocked on a consumer
            for (auto* vfo : live) { if (vfo->pendingOut > 0) { vfo->out.swap(vfo->pendingOut); } }
            for (auto* s : taps) { if (nraw > 0) { s->swap(nraw); } }
        }
    }

    dsp::stream<dsp::complex_t>* _in = nullptr;
    sdrpp_cuda_frontend* fe = nullptr;
    std::recursive_mutex mtx;
    std::thread worker;
    bool running = false;
    std::map<std::string, dsp::channel::RxVFO*> vfos;
    std::vector<dsp::stream<dsp::complex_t>*> bound;

    double _sampleRate = 0;
    double _decimRatio = 1;
    int _fftSize = 0;
    double _fftRate = 0;
    dsp::window::windowType _fftWindow = dsp::window::NUTTALL;
    float* (*_acquireFFTBuffer)(void* ctx) = nullptr;
    void (*_releaseFFTBuffer)(void* ctx) = nullptr;
    void* _fftCtx = nullptr;
    double effectiveSr = 0;
    bool _init = false;
};
