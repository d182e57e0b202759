// dsp::compression::SampleStreamDecompressor: wire packet (u16 compression | u16 PCMType | f32 scaler | payload) ->
// complex samples (reference: core/src/dsp/compression/sample_stream_decompressor.h:6-53). process() keeps the
// reference's signature and return value (samples written, 0 for an unknown sample type); the conversion runs on the
// GPU through sdrpp_cuda_pcm_decompress. A source that feeds the front end directly can skip this block and hand the
// packet to sdrpp_cuda_frontend_submit_pcm instead.
#pragma once
#include "../processor.h"
#include <sdrpp_cuda.h>
#include "pcm_type.h"

namespace dsp::compression {
    class SampleStreamDecompressor : public Processor<uint8_t, complex_t> {
        using base_type = Processor<uint8_t, complex_t>;
    public:
        SampleStreamDecompressor() {}
        SampleStreamDecompressor(stream<uint8_t>* in) { base_type::init(in); }

        // count = packet size in bytes. Errors of the device path (bad pointer, CUDA failure) end the block like a
        // stopped stream does: nothing is written and 0 is returned.
        inline int process(int count, const uint8_t* in, complex_t* out) {
            if (count < 8) { return 0; }
            const int n = sdrpp_cuda_pcm_decompress(in, count, reinterpret_cast<sdrpp_cf32*>(out));
            return n > 0 ? n : 0;
        }

        int run() {
            const int bytes = base_type::_in->read();
            if (bytes < 0) { return -1; }
            const int produced = process(bytes, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (produced > 0 && !base_type::out.swap(produced)) { return -1; }
            return produced;
        }
    };
}
