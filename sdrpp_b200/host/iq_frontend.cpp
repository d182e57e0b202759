// IQFrontEnd on the CUDA library -- replaces core/src/signal_path/iq_frontend.cpp (reference: iq_frontend.cpp:15-296).
// Built into sdrpp_core in place of that file (tools/make_overlay.py maps signal_path/iq_frontend.cpp here); with
// -DSDRPP_HEADLESS it builds without the GUI (tests/cpp/mirror_demo.cpp, no waterfall to resize).
#include "iq_frontend.h"
#include <utils/flog.h>
#include <algorithm>
#include <cstring>
#include <stdexcept>
#ifndef SDRPP_HEADLESS
#include <gui/gui.h>
#include <core.h>
#endif

IQFrontEnd::~IQFrontEnd() {
    if (!_init) { return; }
    stop();
    for (auto& [name, vfo] : vfos) { delete vfo; }
    if (fe) { sdrpp_cuda_frontend_destroy(fe); }
}

void IQFrontEnd::init(dsp::stream<dsp::complex_t>* in, double sampleRate, bool buffering, int decimRatio, bool dcBlocking, int fftSize, double fftRate, dsp::window::windowType fftWindow, float* (*acquireFFTBuffer)(void* ctx), void (*releaseFFTBuffer)(void* ctx), void* fftCtx) {
    _in = in;
    _sampleRate = sampleRate;
    _decimRatio = decimRatio;
    _fftSize = fftSize;
    _fftRate = fftRate;
    _fftWindow = fftWindow;
    _acquireFFTBuffer = acquireFFTBuffer;
    _releaseFFTBuffer = releaseFFTBuffer;
    _fftCtx = fftCtx;
    (void)buffering; // SampleFrameBuffer's 32-slot FIFO (frame_buffer.h:52-94) is replaced by the blocks in flight on the device

    effectiveSr = _sampleRate / _decimRatio;

    sdrpp_cuda_frontend_cfg cfg{};
    cfg.sample_rate = sampleRate;
    cfg.decim_ratio = decimRatio;
    cfg.dc_blocking = dcBlocking;
    cfg.invert_iq = 0;
    cfg.fft_size = fftSize;
    cfg.fft_rate = fftRate;
    cfg.fft_window = (int)fftWindow;
    cfg.max_block = STREAM_BUFFER_SIZE;
    fe = sdrpp_cuda_frontend_create(&cfg);
    if (!fe) { flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error()); }

    _init = true;
}

void IQFrontEnd::updateFFTSize() {
    // window table + transform plan for the current (_fftSize, _fftRate, _fftWindow): all inside the library
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (fe && sdrpp_cuda_frontend_set_fft_size(fe, _fftSize) < 0) { flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error()); }
}

void IQFrontEnd::updateFFTPath(bool updateWaterfall) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (fe) {
        if (sdrpp_cuda_frontend_set_fft_rate(fe, _fftRate) < 0 || sdrpp_cuda_frontend_set_fft_window(fe, (int)_fftWindow) < 0) {
            flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error());
        }
    }
    updateFFTSize();
#ifndef SDRPP_HEADLESS
    if (updateWaterfall) { gui::waterfall.setRawFFTSize(_fftSize); }
#else
    (void)updateWaterfall;
#endif
}

void IQFrontEnd::setInput(dsp::stream<dsp::complex_t>* in) {
    const bool was = running;
    stop();
    _in = in;
    if (was) { start(); }
}

void IQFrontEnd::setSampleRate(double sampleRate) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    _sampleRate = sampleRate;
    effectiveSr = _sampleRate / _decimRatio;
    // re-plans every VFO for the new input rate and reconfigures the spectrum (iq_frontend.cpp:55-80)
    if (fe && sdrpp_cuda_frontend_set_sample_rate(fe, sampleRate) < 0) { flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error()); }
    for (auto& [name, vfo] : vfos) { vfo->noteInSamplerate(effectiveSr); }
}

void IQFrontEnd::setBuffering(bool enabled) { (void)enabled; }

void IQFrontEnd::setDecimation(int ratio) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    _decimRatio = ratio;
    effectiveSr = _sampleRate / _decimRatio;
    if (fe && sdrpp_cuda_frontend_set_decimation(fe, ratio) < 0) { flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error()); }
    for (auto& [name, vfo] : vfos) { vfo->noteInSamplerate(effectiveSr); }
#ifndef SDRPP_HEADLESS
    core::setInputSampleRate(_sampleRate);
#endif
}

void IQFrontEnd::setDCBlocking(bool enabled) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (fe) { sdrpp_cuda_frontend_set_dc_blocking(fe, enabled); }
}

void IQFrontEnd::setInvertIQ(bool enabled) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (fe) { sdrpp_cuda_frontend_set_invert_iq(fe, enabled); }
}

// Raw (post-preprocessing) IQ taps, as the Splitter hands them out (recorder, iq_frontend.cpp:114-120;
// Splitter::bindStream / unbindStream throw on misuse, splitter.h:19,36)
void IQFrontEnd::bindIQStream(dsp::stream<dsp::complex_t>* stream) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (std::find(bound.begin(), bound.end(), stream) != bound.end()) {
        throw std::runtime_error("[Splitter] Tried to bind stream to that is already bound");
    }
    bound.push_back(stream);
}

void IQFrontEnd::unbindIQStream(dsp::stream<dsp::complex_t>* stream) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    auto it = std::find(bound.begin(), bound.end(), stream);
    if (it == bound.end()) {
        throw std::runtime_error("[Splitter] Tried to unbind stream to that isn't bound");
    }
    stream->stopWriter();
    { std::lock_guard<std::mutex> sw(swapMtx); bound.erase(it); }
    stream->clearWriteStop();
}

dsp::channel::RxVFO* IQFrontEnd::addVFO(std::string name, double sampleRate, double bandwidth, double offset) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    // Make sure no other VFO with that name already exists
    if (vfos.find(name) != vfos.end()) {
        flog::error("[IQFrontEnd] Tried to add VFO with existing name.");
        return NULL;
    }
    if (!fe) { return NULL; }
    const int id = sdrpp_cuda_vfo_create(fe, sampleRate, bandwidth, offset, SDRPP_DEMOD_NONE);
    if (id < 0) {
        flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error());
        return NULL;
    }
    dsp::channel::RxVFO* vfo = new dsp::channel::RxVFO();
    vfo->attach(fe, id, &mtx, effectiveSr, sampleRate, bandwidth, offset);
    vfos[name] = vfo;
    return vfo;
}

void IQFrontEnd::removeVFO(std::string name) {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    auto it = vfos.find(name);
    if (it == vfos.end()) {
        flog::error("[IQFrontEnd] Tried to remove a VFO that doesn't exist.");
        return;
    }
    dsp::channel::RxVFO* vfo = it->second;
    vfo->out.stopWriter();                       // a delivery blocked on this VFO's consumer returns
    {
        std::lock_guard<std::mutex> sw(swapMtx); // the deliver thread is not (and cannot get) inside its swap phase
        sdrpp_cuda_vfo_destroy(fe, vfo->vfoId);
        vfos.erase(it);
    }
    delete vfo;
}

void IQFrontEnd::setFFTSize(int size) {
    _fftSize = size;
    updateFFTPath(true);
}

void IQFrontEnd::setFFTRate(double rate) {
    _fftRate = rate;
    updateFFTPath();
}

void IQFrontEnd::setFFTWindow(dsp::window::windowType fftWindow) {
    _fftWindow = fftWindow;
    updateFFTPath();
}

void IQFrontEnd::flushInputBuffer() {}

void IQFrontEnd::start() {
    std::lock_guard<std::recursive_mutex> lck(mtx);
    if (running || !_init || !_in) { return; }
    running = true;
    stopping = false;
    if (fe) { sdrpp_cuda_frontend_drain(fe); }
    submitted = delivered = 0;
    rowsPosted = rowsDone = 0; shardGen = 0; shardsLeft = 0; rowsQueue.clear();
    ingestThread = std::thread(&IQFrontEnd::ingestLoop, this);
    const unsigned hw = std::thread::hardware_concurrency();
    const int nhelp = hw >= 16 ? 3 : hw >= 8 ? 1 : 0;
    for (int k = 0; k < nhelp; k++) { helpers.emplace_back(&IQFrontEnd::helperLoop, this, k + 1); }
    spectrumThread = std::thread(&IQFrontEnd::spectrumLoop, this);
    deliverThread = std::thread(&IQFrontEnd::deliverLoop, this);
}

void IQFrontEnd::stop() {
    {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        if (!running) { return; }
        running = false;
    }
    // like block::doStop (block.h:78-97): stop the reader side of the input and the writer side of every output, join
    if (_in) { _in->stopReader(); }
    { std::lock_guard<std::mutex> f(flowMtx); stopping = true; }
    flowCv.notify_all();
    {
        std::lock_guard<std::recursive_mutex> lck(mtx);
        for (auto& [name, vfo] : vfos) { vfo->out.stopWriter(); }
        for (auto* s : bound) { s->stopWriter(); }
    }
    if (ingestThread.joinable()) { ingestThread.join(); }
    if (deliverThread.joinable()) { deliverThread.join(); }
    if (spectrumThread.joinable()) { spectrumThread.join(); }
    for (auto& h : helpers) { if (h.joinable()) { h.join(); } }
    helpers.clear();
    if (_in) { _in->clearReadStop(); }
    std::lock_guard<std::recursive_mutex> lck(mtx);
    for (auto& [name, vfo] : vfos) { vfo->out.clearWriteStop(); }
    for (auto* s : bound) { s->clearWriteStop(); }
    if (fe) { sdrpp_cuda_frontend_drain(fe); } // blocks still in flight at a stop are dropped, as in the reference
}

double IQFrontEnd::getEffectiveSamplerate() {
    return effectiveSr;
}

long long IQFrontEnd::blocksDelivered() {
    std::lock_guard<std::mutex> f(flowMtx);
    return delivered;
}

// input stream -> device. The input buffer is handed back (flush) as soon as its H2D copy has left it; the block's
// kernels and the delivery of earlier blocks overlap the next read.
void IQFrontEnd::ingestLoop() {
    while (true) {
        const int count = _in->read();
        if (count < 0) { return; }
        {
            // at most kMaxAhead blocks ahead of the deliver thread; with raw IQ taps bound the tap reads the LAST block,
            // so the pipeline runs one block deep then
            std::unique_lock<std::mutex> f(flowMtx);
            flowCv.wait(f, [&] { return stopping || submitted - delivered < (bound.empty() ? kMaxAhead : 1); });
            if (stopping) { return; }
        }
        bool ok;
        {
            std::lock_guard<std::recursive_mutex> lck(mtx);
            ok = fe && sdrpp_cuda_frontend_submit(fe, SDRPP_FMT_CF32, _in->readBuf, count) >= 0;
        }
        if (ok) { ok = sdrpp_cuda_frontend_wait_input(fe) >= 0; }
        _in->flush();
        if (!ok) {
            flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error());
            continue;
        }
        { std::lock_guard<std::mutex> f(flowMtx); submitted++; }
        flowCv.notify_all();
    }
}

void IQFrontEnd::deliverLoop() {
    while (true) {
        {
            std::unique_lock<std::mutex> f(flowMtx);
            flowCv.wait(f, [&] { return stopping || submitted > delivered; });
            if (stopping) { return; }
        }
        if (sdrpp_cuda_frontend_wait(fe) < 0) {    // blocks on the block's completion event, no lock held
            flog::error("[IQFrontEnd] {0}", sdrpp_cuda_last_error());
        } else {
            deliverBlock();
        }
        {
            // the spectrum thread may lag one block: rows of block d are still intact while block d+1 is delivered
            std::unique_lock<std::mutex> f(flowMtx);
            flowCv.wait(f, [&] { return stopping || rowsDone >= rowsPosted - 1; });
            delivered++;
        }
        flowCv.notify_all();
    }
}

void IQFrontEnd::deliverBlock() {
    std::unique_lock<std::recursive_mutex> lck(mtx);
    // spectrum rows go to the spectrum thread
    const float* rows = nullptr;
    const int nrows = sdrpp_cuda_fft_rows(fe, &rows);
    // VFO blocks into each RxVFO::out (Splitter::run + RxVFO::run of the reference, splitter.h:46-60, rx_vfo.h:102-114)
    outItems.clear();
    for (auto& [name, vfo] : vfos) {
        const sdrpp_cf32* iq = nullptr;
        const int n = sdrpp_cuda_vfo_output(fe, vfo->vfoId, &iq, nullptr);
        if (n > 0 && iq) { outItems.push_back(OutItem{ vfo, iq, n }); }
    }
    outTaps = bound;
    outRaw = 0;
    if (!outTaps.empty()) {
        outRaw = sdrpp_cuda_frontend_read_iq(fe, (sdrpp_cf32*)outTaps[0]->writeBuf, STREAM_BUFFER_SIZE);
        for (size_t i = 1; i < outTaps.size(); i++) { memcpy(outTaps[i]->writeBuf, outTaps[0]->writeBuf, sizeof(dsp::complex_t) * (size_t)std::max(outRaw, 0)); }
    }
    // never hold the control mutex while blocked on a consumer; swapMtx keeps removeVFO / unbindIQStream out meanwhile
    std::lock_guard<std::mutex> sw(swapMtx);
    lck.unlock();
    {
        std::lock_guard<std::mutex> f(flowMtx);
        if (nrows > 0) { rowsQueue.emplace_back(rows, nrows); rowsPosted++; }
        shardGen++;
        shardsLeft = (int)helpers.size();
    }
    flowCv.notify_all();
    deliverShard(0);
    for (auto* s : outTaps) { if (outRaw > 0) { s->swap(outRaw); } }
    std::unique_lock<std::mutex> f(flowMtx);
    flowCv.wait(f, [&] { return shardsLeft == 0 || stopping; });
}

// copy + swap of every (1 + helpers)-th VFO block: 512 VFOs are 512 condition-variable hand-overs per IQ block
void IQFrontEnd::deliverShard(int k) {
    const size_t stride = helpers.size() + 1;
    for (size_t i = (size_t)k; i < outItems.size(); i += stride) {
        const OutItem& it = outItems[i];
        memcpy(it.vfo->out.writeBuf, it.iq, sizeof(dsp::complex_t) * (size_t)it.n);
        it.vfo->out.swap(it.n);
    }
}

void IQFrontEnd::helperLoop(int k) {
    long long seen = 0;
    while (true) {
        {
            std::unique_lock<std::mutex> f(flowMtx);
            flowCv.wait(f, [&] { return stopping || shardGen > seen; });
            if (shardGen == seen) { return; }   // stopping, nothing posted
            seen = shardGen;
        }
        deliverShard(k);
        { std::lock_guard<std::mutex> f(flowMtx); shardsLeft--; }
        flowCv.notify_all();
    }
}

// acquire/release are always called as a pair, once per line (iq_frontend.cpp:239-248); acquire may return NULL (the
// row is skipped). The reference runs this on the FFT sink's own thread too.
void IQFrontEnd::spectrumLoop() {
    while (true) {
        const float* rows; int nrows;
        {
            std::unique_lock<std::mutex> f(flowMtx);
            flowCv.wait(f, [&] { return stopping || rowsPosted > rowsDone; });
            if (rowsQueue.empty()) { return; }
            rows = rowsQueue.front().first; nrows = rowsQueue.front().second;
            rowsQueue.pop_front();
        }
        for (int r = 0; r < nrows; r++) {
            float* dst = _acquireFFTBuffer ? _acquireFFTBuffer(_fftCtx) : nullptr;
            if (dst) { memcpy(dst, rows + (size_t)r * _fftSize, sizeof(float) * (size_t)_fftSize); }
            if (_releaseFFTBuffer) { _releaseFFTBuffer(_fftCtx); }
        }
        { std::lock_guard<std::mutex> f(flowMtx); rowsDone++; }
        flowCv.notify_all();
    }
}
