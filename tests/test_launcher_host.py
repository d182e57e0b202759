"""CPU: the command list behind the CUDA-graph replay (sdrpp_b200/csrc/launcher.h) as a host-only unit: planning makes no
CUDA call, descriptor records, argument packing, and the byte-identity of repeated sequences that the graph cache keys on."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launcher_plans_without_cuda_calls(tmp_path):
    exe = str(tmp_path / "launcher_test")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "tests", "cpp", "launcher_test.cpp"),
                           "-o", exe, "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "launcher_test: ok" in r.stdout, r.stdout + r.stderr
