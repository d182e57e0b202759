#!/usr/bin/env python3
"""Builds the drop-in source overlay of the reference's core/src.

    python tools/make_overlay.py [--ref /root/reference] [--out build/overlay]

The overlay is a directory tree that mirrors <ref>/core/src with every file a SYMLINK into the reference tree -- no
reference file is copied -- except for the few files of the hot path, which point at this repository's replacements:

    dsp/stream.h                                   -> include/sdrpp/dsp/stream.h            (pinned double buffer)
    dsp/channel/rx_vfo.h                           -> include/sdrpp/dsp/channel/rx_vfo.h    (RxVFO over the C ABI)
    dsp/compression/sample_stream_compressor.h     -> include/sdrpp/dsp/compression/...     (SDR++ server packets)
    dsp/compression/sample_stream_decompressor.h   -> include/sdrpp/dsp/compression/...
    signal_path/iq_frontend.h / .cpp               -> include/sdrpp/signal_path/iq_frontend.h, sdrpp_b200/host/iq_frontend.cpp

Everything else -- dsp/block.h, dsp/processor.h, dsp/types.h, every demodulator, the GUI, VFOManager, the module API --
is the reference's own file. Passing -I<overlay> INSTEAD of -I<ref>/core/src makes both <dsp/...> includes and the
reference's relative includes ("../processor.h", "../dsp/channel/rx_vfo.h") resolve inside the overlay (GCC resolves a
quoted include relative to the directory the including file was found in, i.e. the symlink's directory), so exactly one
definition of every header is seen and modules (radio, recorder, scanner, file_source ...) compile unchanged.
A maintainer applies the same substitution in core/CMakeLists.txt by overwriting those six files."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# overlay path (relative to core/src) -> replacement (relative to this repository)
REPLACED = {
    "dsp/stream.h": "include/sdrpp/dsp/stream.h",
    "dsp/channel/rx_vfo.h": "include/sdrpp/dsp/channel/rx_vfo.h",
    "dsp/compression/sample_stream_compressor.h": "include/sdrpp/dsp/compression/sample_stream_compressor.h",
    "dsp/compression/sample_stream_decompressor.h": "include/sdrpp/dsp/compression/sample_stream_decompressor.h",
    "signal_path/iq_frontend.h": "include/sdrpp/signal_path/iq_frontend.h",
    "signal_path/iq_frontend.cpp": "sdrpp_b200/host/iq_frontend.cpp",
}


def build(ref="/root/reference", out=None, quiet=True):
    src = os.path.join(ref, "core", "src")
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found")
    out = out or os.path.join(ROOT, "build", "overlay")
    os.makedirs(out, exist_ok=True)

    def link(target, path):
        if os.path.islink(path):
            if os.readlink(path) == target:
                return
            os.unlink(path)
        elif os.path.exists(path):
            raise RuntimeError(f"{path} exists and is not a symlink")
        os.symlink(target, path)

    # Every DIRECTORY is a real directory and every FILE one symlink: a symlinked directory would send a relative
    # include that climbs out of it ("../stream.h" from dsp/demod/) to the reference's parent directory, past the
    # replaced files (the kernel resolves "dir/.." through the link).
    for dirpath, dirnames, filenames in os.walk(src):
        rel = os.path.relpath(dirpath, src)
        rel = "" if rel == "." else rel
        odir = os.path.join(out, rel)
        if os.path.islink(odir):
            os.unlink(odir)
        os.makedirs(odir, exist_ok=True)
        for name in filenames:
            r = os.path.join(rel, name) if rel else name
            if r not in REPLACED:
                link(os.path.join(dirpath, name), os.path.join(odir, name))
    for rel, repl in REPLACED.items():
        link(os.path.join(ROOT, repl), os.path.join(out, rel))
    if not quiet:
        print(out)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    try:
        build(a.ref, a.out, quiet=False)
    except Exception as e:  # noqa: BLE001
        sys.exit(str(e))


# ---------------------------------------------------------------------------------------------------------------------
# Compile recipes on top of the overlay (used by tests/test_module_compile.py, tests/test_cpp_mirror.py, build())
# ---------------------------------------------------------------------------------------------------------------------
def fmt_include():
    """fmt is not installed in this image; torch bundles a header-only copy."""
    import importlib.util
    spec = importlib.util.find_spec("torch")
    return os.path.join(os.path.dirname(spec.origin), "include")


def module_flags(overlay):
    """Include path for compiling UNMODIFIED reference sources (modules, core/src/*.cpp) against the overlay. Shims stand
    in for libraries that are not installed here and are never linked: volk, fftw3, GL, FLAC, lame (tests/shims, oracle/shim)."""
    return ["-std=c++17", "-DFMT_HEADER_ONLY", "-I" + overlay, "-I" + os.path.join(overlay, "imgui"), "-I" + os.path.join(ROOT, "include"),
            "-I" + os.path.join(ROOT, "tests", "shims"), "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + fmt_include()]


DEMO = os.path.join(ROOT, "tests", "cpp", "mirror_demo")


def _demo_digest():
    import hashlib
    h = hashlib.sha256()
    paths = [os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp"), os.path.join(ROOT, "sdrpp_b200", "host", "iq_frontend.cpp"),
             os.path.join(ROOT, "include", "sdrpp_cuda.h")]
    for base in ("include/sdrpp", "include/sdrpp_headless"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            paths += [os.path.join(dp, f) for f in fs]
    for p in sorted(paths):
        h.update(os.path.relpath(p, ROOT).encode())   # relative: the repository lives elsewhere on the GPU box
        h.update(open(p, "rb").read())
    return h.hexdigest()


def demo_is_current():
    stamp = DEMO + ".stamp"
    return os.path.exists(DEMO) and os.path.exists(stamp) and open(stamp).read() == _demo_digest()


def build_demo(ref="/root/reference"):
    """tests/cpp/mirror_demo: a GUI-less program that drives the path like an SDR++ module, compiled against the overlay
    -- the reference's own dsp/block.h, processor.h, types.h, utils/flog.cpp, threading.cpp -- with the replaced
    IQFrontEnd / RxVFO / stream on top, and linked against libsdrpp_cuda.so only. Built here (the reference tree is not
    on the GPU box); the binary travels with the snapshot and a content stamp says whether it is current."""
    import subprocess
    if demo_is_current():
        return DEMO
    ov = build(ref)
    cmd = ["g++", "-O2", "-DSDRPP_HEADLESS", "-I" + os.path.join(ROOT, "include", "sdrpp_headless")] + module_flags(ov) + [
        os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp"), os.path.join(ov, "signal_path", "iq_frontend.cpp"),
        os.path.join(ov, "utils", "flog.cpp"), os.path.join(ov, "utils", "threading.cpp"), os.path.join(ov, "utils", "stack_trace.cpp"),
        "-o", DEMO, "-L" + os.path.join(ROOT, "sdrpp_b200"), "-lsdrpp_cuda", "-Wl,-rpath,$ORIGIN/../../sdrpp_b200", "-lpthread", "-ldl"]
    subprocess.check_call(cmd)
    with open(DEMO + ".stamp", "w") as f:
        f.write(_demo_digest())
    return DEMO
