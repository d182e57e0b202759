// dsp::window::windowType / createWindow (reference: core/src/dsp/window/window.h:28-64). The enum order
// is persisted as an int in the config (core.cpp:134), so it is part of the interface. The table
// itself comes from the library's host-side design code (bit-identical to the reference).
#pragma once
#include "../../../sdrpp_cuda.h"

namespace dsp::window {
    enum windowType {
        RECTANGULAR,
        HAMMING,
        HANN,
        BLACKMAN,
        NUTTALL,
        BLACKMAN_HARRIS4,
        BLACKMAN_HARRIS7
    };

    // buffer must hold size+1 floats (the centred form touches buffer[size] when size is odd)
    inline void createWindow(windowType type, float* buffer, int size, bool centered) {
        sdrpp_cuda_design_window((int)type, buffer, size, centered ? 1 : 0);
    }
}
