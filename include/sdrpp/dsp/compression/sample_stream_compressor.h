// dsp::compression::SampleStreamCompressor: complex samples -> wire packet with a per-block scaler (reference:
// core/src/dsp/compression/sample_stream_compressor.h:6-80). The static process() keeps the reference's signature and
// return value (packet size in bytes); block maximum and saturating int8/int16 packing run on the GPU through
// sdrpp_cuda_pcm_compress, bit-identical to the reference's generic-VOLK result.
#pragma once
#include "../processor.h"
#include <sdrpp_cuda.h>
#include "pcm_type.h"

namespace dsp::compression {
    class SampleStreamCompressor : public Processor<complex_t, uint8_t> {
        using base_type = Processor<complex_t, uint8_t>;
    public:
        SampleStreamCompressor() {}
        SampleStreamCompressor(stream<complex_t>* in, PCMType pcmType) { init(in, pcmType); }

        void init(stream<complex_t>* in, PCMType pcmType) {
            _pcmType = pcmType;
            base_type::init(in);
        }

        void setPCMType(PCMType pcmType) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            _pcmType = pcmType;
            base_type::tempStart();
        }

        // out must hold 8 + count * sizeof(complex_t) bytes (the float32 case), like the reference's stream buffer.
        inline static int process(int count, PCMType pcmType, const complex_t* in, uint8_t* out) {
            const int bytes = sdrpp_cuda_pcm_compress((int)pcmType, reinterpret_cast<const sdrpp_cf32*>(in), count, out);
            return bytes > 0 ? bytes : 0;
        }

        int run() {
            const int count = base_type::_in->read();
            if (count < 0) { return -1; }
            const int bytes = process(count, _pcmType, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (bytes > 0 && !base_type::out.swap(bytes)) { return -1; }
            return bytes;
        }

    protected:
        PCMType _pcmType = PCM_TYPE_I16;
    };
}
