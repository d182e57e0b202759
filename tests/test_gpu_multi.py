"""Two ranks on two GPUs through the library's own communicator (sdrpp_cuda_comm_*, NCCL over NVLink): the VFO set
sharded, every raw block broadcast inside sdrpp_cuda_frontend_submit*, and each rank's spot VFOs checked against the
oracle (bench.py's parity_check leg). Needs two CUDA devices: skipped on a one-GPU box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [2, 5])
def test_sharded_outputs_match_oracle_on_every_rank(gpu, cfg):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + cfg), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "20", "--warmup", "3",
           "--config", str(cfg), "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2
    par = line["parity_check"]
    assert par["ok"], par
    assert par["ranks"] == 2 and par["vfos_checked"] >= min(8, line["config"]["vfos"])
    assert par["worst_iq_rel_rms_vs_ideal_nco_oracle"] <= 1e-5
    assert line["comm"]["nranks"] == 2 and line["comm"]["broadcasts"] > 20
