#!/usr/bin/env python3
"""Extract the PowerDecimator plan table from the reference into a data blob.

The cascaded-decimator FIR coefficients (core/src/dsp/multirate/decim/plans.h:126-140 and
decim/taps/fir_*.h) are numeric data that cannot be regenerated: the script that produced them is
not part of the reference tree (plans.h:17-22). Results must match the reference, so the numbers
themselves are needed. This tool reads them THROUGH the compiled reference (oracle/_ref, built by
`make -C oracle ref`, which compiles the reference headers where they lie) -- the values are the
compiler's own double->float roundings of the literals, and the declared lengths are honoured
(fir_4_2 declares 12 taps although its initialiser lists 13 values).

Output: sdrpp_b200/data/decim_plans.bin (little-endian)
    u32 magic 'SPDP' (0x50445053), u32 version=1, u32 n_firs, u32 n_plans, u32 pool_len
    firs [n_firs]  : u32 len, u32 pool_off
    plans[n_plans] : u32 ratio, u32 n_stages, 4 x (u32 decimation, u32 fir_idx)
    f32 pool[pool_len]
Run here only (needs /root/reference); the blob is committed.
"""
import ctypes, os, struct, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsdrpp_ref.so"))
    lib.ref_decim_plan.restype = ctypes.c_int
    firs, fir_index, plans = [], {}, []
    for k in range(1, 14):
        ratio = 1 << k
        dec = (ctypes.c_int * 4)(); cnt = (ctypes.c_int * 4)(); ptr = (ctypes.POINTER(ctypes.c_float) * 4)()
        n = lib.ref_decim_plan(ratio, dec, cnt, ptr)
        assert 1 <= n <= 4, (ratio, n)
        stages = []
        for i in range(n):
            taps = np.ctypeslib.as_array(ptr[i], shape=(cnt[i],)).astype(np.float32).copy()
            key = (taps.tobytes(), cnt[i])
            if key not in fir_index:
                fir_index[key] = len(firs)
                firs.append(taps)
            stages.append((dec[i], fir_index[key]))
        prod = 1
        for d, _ in stages: prod *= d
        assert prod == ratio, (ratio, stages)
        plans.append((ratio, stages))
    pool = np.concatenate(firs)
    out = bytearray()
    out += struct.pack("<5I", 0x50445053, 1, len(firs), len(plans), len(pool))
    off = 0
    for t in firs:
        out += struct.pack("<2I", len(t), off); off += len(t)
    for ratio, stages in plans:
        out += struct.pack("<2I", ratio, len(stages))
        for i in range(4):
            d, f = stages[i] if i < len(stages) else (0, 0)
            out += struct.pack("<2I", d, f)
    out += pool.astype("<f4").tobytes()
    path = os.path.join(ROOT, "sdrpp_b200", "data", "decim_plans.bin")
    with open(path, "wb") as f: f.write(out)
    print(f"wrote {path}: {len(firs)} firs, {len(plans)} plans, {len(pool)} taps, {len(out)} bytes")
    for ratio, stages in plans:
        print(ratio, [(d, len(firs[f])) for d, f in stages])

if __name__ == "__main__":
    sys.exit(main())
