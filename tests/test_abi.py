"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/sdrpp_cuda.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdrpp_cuda.h")).read()
    return sorted(set(re.findall(r"SDRPP_API[^;(]*?\b(sdrpp_cuda_\w+)\s*\(", text)))


def test_header_symbols_exported(cuda_lib):
    lib = ctypes.CDLL(cuda_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdrpp_cuda.h but not exported"
    assert sorted(cuda_lib.SYMBOLS) == names, "sdrpp_b200.cuda.SYMBOLS out of sync with the header"


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "sdrpp_cuda.h")).read()
    for cite in ("iq_frontend.cpp", "rx_vfo.h", "window.h", "rational_resampler.h", "quadrature.h", "rtl_sdr_source"):
        assert cite in text


def test_no_cpu_fallback(cuda_lib):
    if cuda_lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cuda_lib.SdrppCudaError):
        cuda_lib.convert(cuda_lib.FMT_U8_RTL, np.zeros(16, np.uint8))
    with pytest.raises(cuda_lib.SdrppCudaError):
        cuda_lib.Frontend(2.4e6)
    with pytest.raises(cuda_lib.SdrppCudaError):
        cuda_lib.init(0)
    assert "CUDA" in cuda_lib.last_error() or "fallback" in cuda_lib.last_error()


def test_product_never_imports_oracle():
    """The product path (sdrpp_b200/, include/) must not reference the oracle."""
    bad = []
    for base in ("sdrpp_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle|liboracle|oracle/_ref|pyoracle", t, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
