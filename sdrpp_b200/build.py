"""Build recipe for the CUDA library (sdrpp_b200/libsdrpp_cuda.so), sm_100a only.

    python -m sdrpp_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is built in-tree (git-ignored) so it travels to the GPU
box with the repo snapshot. No torch dependency: the library is plain CUDA runtime behind a C ABI.
"""
import hashlib
import os
import struct
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsdrpp_cuda.so")
BLOB = os.path.join(HERE, "data", "decim_plans.bin")

SOURCES = ["design.cpp", "preproc.cu", "fft.cu", "channelizer.cu", "channelizer_tc.cu", "comm.cu", "engine.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "--expt-relaxed-constexpr",
    # device code may contract to FMA (fp32 tolerance applies); conversions use explicit _rn intrinsics
    "-I", os.path.join(ROOT, "include"),
] + os.environ.get("SDRPP_EXTRA_NVCC", "").split()  # experiment knob, e.g. "-DSDRPP_S1_R=4 -DSDRPP_S1_W=12"


def _gen_blob_inc():
    """decim_plans.bin -> csrc/decim_plans_blob.inc (comma-separated u32 words, generated file)."""
    out = os.path.join(CSRC, "decim_plans_blob.inc")
    data = open(BLOB, "rb").read()
    assert len(data) % 4 == 0
    words = struct.unpack("<%dI" % (len(data) // 4), data)
    text = ",\n".join(", ".join("0x%08xu" % w for w in words[i:i + 8]) for i in range(0, len(words), 8)) + "\n"
    if not os.path.exists(out) or open(out).read() != text:
        with open(out, "w") as f:
            f.write(text)
    return out


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    _gen_blob_inc()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sdrpp_cuda.h")]
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if src.endswith(".cu"):
            cmd += ["-Xptxas", "-v"] if verbose else []
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for obj, log in results:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", LIB, *[o for o, _ in results], "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
