"""GPU parity at the BASELINE.json workloads themselves, whole configuration at once: the exact
stream format, block size, front-end decimation, spectrum size/window and VFO set that bench.py --config N times,
checked against the oracle by tests/parity_workload.py (ideal-NCO chain <= 1e-5 on spot VFOs of every class and tile
position, demod front ends stage-isolated, per-block counts exact, spectrum rows <= 0.01 dB)."""
import json

import pytest

from sdrpp_b200 import workloads
from tests import parity_workload

pytestmark = pytest.mark.gpu

# blocks per config: enough for two spectrum frames where that is affordable (cfg4: one 1M-point frame of the x4-decimated stream)
CASES = [(2, 40, 0), (3, 22, 0), (4, 14, 0), (5, 4, 0), (5, 4, 1)]


@pytest.mark.parametrize("cfg,nblocks,mode", CASES)
def test_whole_config_against_oracle(gpu, report, cfg, nblocks, mode):
    """cfg 5 mode 0 is exactly the benchmarked configuration: 512 alternating NFM/AM VFOs with demod, 614,400-sample cf32
    blocks, saturated 1M-point Blackman-Harris-4 spectrum, tensor-core stage 1; mode 1 = FP32 stage 1 on the same input."""
    w = workloads.config(cfg)
    res = parity_workload.check_workload(gpu, w, nblocks=nblocks, stage1_mode=mode)
    report(f"cfg{cfg} whole config mode{mode}", **{k: v for k, v in res.items() if k not in ("workload", "failures", "adjudicated")},
           adjudicated=json.dumps(res["adjudicated"]))
    assert res["ok"], res["failures"]
    assert res["vfos_checked"] >= min(w.nvfo, 32)
    if cfg in (4, 5) and mode == 0:
        assert res["stage1_tensor_launches"] >= nblocks - 1, "the tensor-core stage 1 did not run"
