// TEST INFRASTRUCTURE: the oracle's generic-VOLK shim plus the VOLK names that only off-path code of the SDR++ modules
// uses, so that unmodified module sources pass `g++ -fsyntax-only` against the drop-in overlay
// (tests/test_module_compile.py). Real VOLK is not installed in this image.
#pragma once
#include "../../../oracle/shim/volk/volk.h"
static inline void volk_64f_convert_32f(float* out, const double* in, unsigned int n) { for (unsigned int i = 0; i < n; i++) out[i] = (float)in[i]; }
