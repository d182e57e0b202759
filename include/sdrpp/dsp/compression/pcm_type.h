// Sample types of the SDR++ server wire packet (reference: core/src/dsp/compression/pcm_type.h:4-8; the values go
// over the wire in the packet header, so the order is part of the protocol).
#pragma once

namespace dsp::compression {
    enum PCMType { PCM_TYPE_I8 = 0, PCM_TYPE_I16 = 1, PCM_TYPE_F32 = 2 };
}
