"""Run-to-run determinism probe: the same ragged-block stream through the tensor-core and the FP32 stage 1, several
times each; every run of a mode must reproduce the first one bit for bit. Prints where a run differs.
  python tools/flake_probe.py [iterations]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import pyoracle as po          # noqa: E402  (probe only; nothing here is a product path)
from sdrpp_b200 import cuda as gpu         # noqa: E402
from sdrpp_b200 import synth               # noqa: E402


def run(sr, vfos, blocks, mode, mb):
    out = [[] for _ in vfos]
    with gpu.Frontend(sr, max_block=mb) as fe:
        fe.set_stage1_mode(mode)
        ids = [fe.add_vfo(*v) for v in vfos]
        for b in blocks:
            fe.process(po.FMT_CF32, b)
            for i, vid in enumerate(ids):
                out[i].append(fe.vfo_output(vid)[0].copy())
    return out


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    gpu.init(0)
    inSR, outSR, bw, blk = 61.44e6, 48e3, 12.5e3, 307200
    sizes = [blk, blk // 3 + 1, 7, blk - 5, 513, blk]
    offs = synth.vfo_grid(37, inSR)
    x = synth.baseband(sum(sizes), inSR, 22, carriers=[(float(o), "fm") for o in offs[::4]], noise_dbfs=-50.0).astype(np.complex64)
    blocks, p = [], 0
    for s in sizes:
        blocks.append(x[p:p + s]); p += s
    vfos = [(outSR, bw, float(o), po.DEMOD_NONE) for o in offs]
    first = {}
    for it in range(iters):
        for mode in (0, 1):
            o = run(inSR, vfos, blocks, mode, blk)
            if mode not in first:
                first[mode] = o
                continue
            for v in range(len(vfos)):
                for b in range(len(blocks)):
                    a, c = first[mode][v][b], o[v][b]
                    if len(a) != len(c) or not np.array_equal(a.view(np.uint32), c.view(np.uint32)):
                        idx = np.nonzero(a != c)[0] if len(a) == len(c) else []
                        print(f"iter {it} mode {mode} vfo {v} block {b}: {len(idx)} of {len(a)} outputs differ, first at {idx[:6]}, "
                              f"max |d| {np.max(np.abs(a - c)) if len(a) == len(c) else -1:.3e}")
    print("done")


if __name__ == "__main__":
    main()
