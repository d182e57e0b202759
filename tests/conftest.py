import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Achieved parity residuals, printed after the run (also under -q) and written to gpurun_out/parity_residuals.txt:
# tests call parity_report("name", gpu_vs_ideal=..., ref_walk=...) so the numbers behind every gate are on record.
_RESIDUALS = []


def parity_report(name, **vals):
    _RESIDUALS.append((name, vals))


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if not _RESIDUALS:
        return
    lines = []
    for name, vals in _RESIDUALS:
        lines.append(name + ": " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))
    terminalreporter.section("parity residuals (achieved, beside each gate)")
    for ln in lines:
        terminalreporter.write_line(ln)
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_residuals.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
    except OSError:
        pass


def _ensure_port():
    path = os.path.join(ROOT, "oracle", "liboracle_port.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    return path


def _ensure_lib():
    from sdrpp_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (checker only)."""
    _ensure_port()
    from oracle import pyoracle
    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own dsp/ headers compiled in oracle/_ref (skips when absent)."""
    from oracle import pyoracle
    if not pyoracle.have_ref():
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def report():
    return parity_report


@pytest.fixture(scope="session")
def ref64():
    from oracle import pyoracle
    if not pyoracle.have_ref("f64"):
        pytest.skip("oracle/_ref f64 flavour not built")
    return pyoracle.Ref("f64")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library (built if needed). No GPU required to load it."""
    _ensure_lib()
    from sdrpp_b200 import cuda
    cuda.lib()
    return cuda


@pytest.fixture(scope="session")
def gpu(cuda_lib):
    if cuda_lib.device_count() <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box (no CPU fallback exists)")
    cuda_lib.init(0)
    return cuda_lib
