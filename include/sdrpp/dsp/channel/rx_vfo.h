// dsp::channel::RxVFO -- host-side mirror of the reference's per-VFO digital down-converter
// (core/src/dsp/channel/rx_vfo.h:6-135: FrequencyXlator -> RationalResampler -> channel FIR).
//
// The arithmetic runs in the CUDA library. Two modes:
//   * attached: created by IQFrontEnd::addVFO; the VFO is one member of the front end's batched
//     channelizer, IQFrontEnd's worker fills `out` every block, run()/process() are not used;
//   * standalone: constructed directly, like any dsp::Processor; process(count, in, out) pushes the
//     block through a private one-VFO front end (H2D, kernels, D2H inside the call).
// Same public names, argument meaning and error behaviour as the reference class. Replaces core/src/dsp/channel/rx_vfo.h
// in the source overlay (tools/make_overlay.py); dsp::Processor / dsp::block / dsp::complex_t are the reference's own.
#pragma once
#include <cstring>
#include <mutex>
#include "../processor.h"
// the reference's rx_vfo.h brings these two in and code downstream relies on that (e.g. dsp/demod/broadcast_fm.h uses
// dsp::channel::FrequencyXlator without including it)
#include "frequency_xlator.h"
#include "../multirate/rational_resampler.h"
#include <sdrpp_cuda.h>

class IQFrontEnd;

namespace dsp::channel {
    class RxVFO : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        RxVFO() {}
        RxVFO(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            init(in, inSamplerate, outSamplerate, bandwidth, offset);
        }
        ~RxVFO() {
            if (!base_type::_block_init) { return; }
            base_type::stop();
            if (ownsFe && fe) { sdrpp_cuda_frontend_destroy(fe); }
        }

        void init(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            _inSamplerate = inSamplerate; _outSamplerate = outSamplerate; _bandwidth = bandwidth; _offset = offset;
            sdrpp_cuda_frontend_cfg cfg{};
            cfg.sample_rate = inSamplerate; cfg.decim_ratio = 1; cfg.max_block = STREAM_BUFFER_SIZE;
            fe = sdrpp_cuda_frontend_create(&cfg);
            ownsFe = true;
            vfoId = fe ? sdrpp_cuda_vfo_create(fe, outSamplerate, bandwidth, offset, SDRPP_DEMOD_NONE) : -1;
            base_type::init(in);
        }

        void setInSamplerate(double inSamplerate) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            _inSamplerate = inSamplerate;
            if (ownsFe && fe) { sdrpp_cuda_frontend_set_sample_rate(fe, inSamplerate); }
            base_type::tempStart();
        }
        void setOutSamplerate(double outSamplerate, double bandwidth) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            _outSamplerate = outSamplerate; _bandwidth = bandwidth;
            withEngine([&] { sdrpp_cuda_vfo_set_out_samplerate(fe, vfoId, outSamplerate, bandwidth); });
            base_type::tempStart();
        }
        void setBandwidth(double bandwidth) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            _bandwidth = bandwidth;
            withEngine([&] { sdrpp_cuda_vfo_set_bandwidth(fe, vfoId, bandwidth); });
        }
        void setOffset(double offset) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            _offset = offset;
            withEngine([&] { sdrpp_cuda_vfo_set_offset(fe, vfoId, offset); });
        }
        void reset() {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            withEngine([&] { sdrpp_cuda_vfo_reset(fe, vfoId); });
            base_type::tempStart();
        }

        // One block through the VFO; returns the output count. in/out may alias (rx_vfo.h:89-100 passes out,out).
        inline int process(int count, const complex_t* in, complex_t* out) {
            if (!fe || vfoId < 0 || !ownsFe) { return -1; }
            if (sdrpp_cuda_frontend_submit(fe, SDRPP_FMT_CF32, in, count) < 0) { return -1; }
            if (sdrpp_cuda_frontend_wait(fe) < 0) { return -1; }
            const sdrpp_cf32* iq = nullptr;
            const int n = sdrpp_cuda_vfo_output(fe, vfoId, &iq, nullptr);
            if (n > 0) { memcpy(out, iq, sizeof(complex_t) * (size_t)n); }
            return n;
        }

        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            int outCount = process(count, base_type::_in->readBuf, out.writeBuf);
            base_type::_in->flush();
            if (outCount < 0) { return -1; }
            if (outCount) {
                if (!out.swap(outCount)) { return -1; }
            }
            return outCount;
        }

    protected:
        friend class ::IQFrontEnd;
        // attached mode: share the front end's engine; engineMtx serialises control calls with its worker
        void attach(sdrpp_cuda_frontend* shared, int id, std::recursive_mutex* mtx, double inSr, double outSr, double bw, double off) {
            fe = shared; vfoId = id; ownsFe = false; engineMtx = mtx;
            _inSamplerate = inSr; _outSamplerate = outSr; _bandwidth = bw; _offset = off;
            base_type::_in = nullptr;
            base_type::registerOutput(&out);
            base_type::_block_init = true;
        }
        // an attached VFO has no worker of its own: IQFrontEnd's worker feeds `out`
        void doStart() override { if (ownsFe) { base_type::doStart(); } }
        void doStop() override { if (ownsFe) { base_type::doStop(); } }
        // attached mode: the front end re-planned this VFO for a new input rate (IQFrontEnd::setSampleRate / setDecimation)
        void noteInSamplerate(double inSr) { _inSamplerate = inSr; }
        template <class F> void withEngine(F f) {
            if (!fe || vfoId < 0) { return; }
            if (engineMtx) { std::lock_guard<std::recursive_mutex> l(*engineMtx); f(); } else { f(); }
        }

        sdrpp_cuda_frontend* fe = nullptr;
        int vfoId = -1;
        bool ownsFe = false;
        int pendingOut = 0;
        std::recursive_mutex* engineMtx = nullptr;
        double _inSamplerate = 0, _outSamplerate = 0, _bandwidth = 0, _offset = 0;
    };
}
