"""sdrpp_b200: B200-native SDR++ signal-path hot loop.

The product is sdrpp_b200/libsdrpp_cuda.so (hand-written sm_100a kernels behind the C ABI in
include/sdrpp_cuda.h) plus the C++ header mirror of the reference's dsp::/sigpath:: interface in
include/sdrpp/. This package only holds the build recipe and a ctypes binding for tests/bench.
"""
from . import cuda  # noqa: F401
