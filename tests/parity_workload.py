"""TEST INFRASTRUCTURE: whole-workload parity check of the CUDA path against the oracle (SURVEY 8d gates).

Used by tests/test_gpu_workloads.py (every BASELINE config at its full shape) and by bench.py OUTSIDE its timed
region ("parity_check" in the JSON line; on every rank at N > 1, each rank checking VFOs of its own shard). Nothing here
is on the product path: the front end under test is driven through the C ABI exactly as the benchmark drives it, the
oracle (oracle/pyoracle.py, plain-C port; ideal-NCO flavour) is only the checker.

Per spot-checked VFO: per-block output counts exact; complex output <= 1e-5 relative RMS against the ideal-NCO oracle
chain on the same raw input (conversion -> [PowerDecimator] -> RxVFO); demod front end stage-isolated (oracle demod on
the GPU's own VFO output) <= 1e-5. A noise-only channel sits 60-80 dB under the full-band signal whose fp32 rounding
every implementation -- the reference's included -- folds into it; where such a channel misses 1e-5 it is adjudicated
against the fp64-accumulated flavour of the same chain: the GPU's distance to that truth must not exceed twice the fp32
oracle's own (both numbers are reported). Spectrum rows: <= 0.01 dB on bins within 100 dB of the row's peak against the
fp64 DFT of the same windowed frame, or -- for N >= 128K, SURVEY C.10 -- no worse than twice the oracle's own fp32 FFT."""
import numpy as np

from oracle import pyoracle as po

TOL = 1e-5


def default_spots(w, count=32):
    """VFOs to check: the ones that carry a tone plus first / last members of several 16-VFO tiles of each class."""
    n = w.nvfo
    if n <= count:
        return list(range(n))
    ncls = len({(v[0], v[1], v[3]) for v in w.vfos})
    spots = list(w.tone_vfos())
    tile = 16 * ncls                     # VFO ids covered by one tile of each class (classes alternate)
    t = 0
    while len(set(spots)) < count and t * tile < n:
        for k in range(ncls):
            for cand in (t * tile + k, min(n - 1, t * tile + tile - ncls + k)):
                if cand < n:
                    spots.append(cand)
        t += 3
    k = 7
    while len(set(spots)) < count:       # top up with VFOs spread over the set
        spots.append(k % n)
        k += 37
    out = sorted(set(spots))
    return out[:max(count, len(w.tone_vfos()))]


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64 if np.isrealobj(a) else np.complex128)
    b = np.asarray(b, dtype=np.float64 if np.isrealobj(b) else np.complex128)
    den = np.sqrt(np.mean(np.abs(b) ** 2))
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / (den if den > 0 else 1.0))


def check_workload(cuda, w, nblocks=4, spots=None, vfo_ids=None, with_fft=True, stage1_mode=0, submit=None, raw_blocks=None,
                   rows_to_check=2, frontend_setup=None, spot_count=32):
    """Runs `nblocks` blocks of workload `w` through a fresh front end and checks spot VFOs and spectrum rows.

    vfo_ids: the subset of w.vfos this front end owns (a rank's shard; default all). spots: indices INTO w.vfos to
    check (must be owned). submit(fe, fmt, raw_block) -> None: how a block reaches the front end (default: host
    submit + wait; bench.py passes its broadcast path). Returns a JSON-serialisable dict with "ok"."""
    port = po.Port()
    port64 = None
    owned = list(range(w.nvfo)) if vfo_ids is None else list(vfo_ids)
    if spots is None:
        spots = [i for i in default_spots(w, spot_count) if i in set(owned)]
        if vfo_ids is not None and len(spots) < min(spot_count, len(owned)):
            extra = owned[:: max(1, len(owned) // max(1, spot_count - len(spots)))]
            spots = sorted(set(spots + extra))[:max(spot_count, len(spots))]
    raw = w.make_blocks(nblocks) if raw_blocks is None else raw_blocks
    fe = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size if with_fft else 0, fft_rate=w.fft_rate,
                       fft_window=w.fft_window, max_block=w.block)
    res = {"workload": w.describe(), "blocks": int(nblocks), "vfos_owned": len(owned), "vfos_checked": len(spots), "gate": TOL}
    try:
        if frontend_setup is not None:
            frontend_setup(fe)
        fe.set_stage1_mode(stage1_mode)
        ids = {i: fe.add_vfo(*w.vfos[i]) for i in owned}
        got = {i: [] for i in spots}
        rows = []
        for b in range(nblocks):
            if submit is None:
                fe.process(w.fmt, raw[b])
            else:
                submit(fe, w.fmt, raw[b])
                fe.wait()
            for i in spots:
                got[i].append(fe.vfo_output(ids[i]))
            if with_fft:
                r = fe.fft_rows()
                if len(r):
                    rows.append(r)
        res["stage1_tensor_launches"] = int(fe.stage1_tensor_launches)
    finally:
        fe.close()

    # the oracle's input stream at the effective rate
    xs = []
    pd = port.powerdecim(w.decim) if w.decim > 1 else None
    for b in range(nblocks):
        x = raw[b] if w.fmt == po.FMT_CF32 else port.convert(w.fmt, raw[b])
        xs.append(pd.process(x) if pd is not None else np.ascontiguousarray(x))

    worst_iq, worst_dm, worst_walk, adjudicated, failures = 0.0, 0.0, 0.0, [], []
    counts_ok = True
    for i in spots:
        osr, bw, off, dm = w.vfos[i]
        o = port.rxvfo(w.eff_sr, osr, bw, off, ideal_nco=True)
        o32 = port.rxvfo(w.eff_sr, osr, bw, off) if i == spots[0] else None    # the reference rotator's walk, one VFO
        ys = [o.process(x) for x in xs]
        if [len(y) for y in ys] != [len(a) for a, _ in got[i]]:
            counts_ok = False
            failures.append(f"vfo {i}: counts {[len(a) for a, _ in got[i]]} != {[len(y) for y in ys]}")
            continue
        g_iq, r_iq = np.concatenate([a for a, _ in got[i]]), np.concatenate(ys)
        err = _rel(g_iq, r_iq)
        if o32 is not None:
            worst_walk = _rel(np.concatenate([o32.process(x) for x in xs]), r_iq)
        if err > TOL:
            # adjudicate against the fp64-accumulated chain (same ideal NCO): GPU must be within 2x the fp32 oracle's own scatter
            if port64 is None:
                port64 = po.Port("f64")
            t = port64.rxvfo(w.eff_sr, osr, bw, off, ideal_nco=True)
            truth = np.concatenate([t.process(x) for x in xs])
            e_gpu, e_orc = _rel(g_iq, truth), _rel(r_iq, truth)
            adjudicated.append({"vfo": int(i), "gpu_vs_fp32_oracle": err, "gpu_vs_f64_truth": e_gpu, "fp32_oracle_vs_f64_truth": e_orc})
            if e_gpu > max(TOL, 2.0 * e_orc):
                failures.append(f"vfo {i}: iq {err:.3e}; vs f64 truth {e_gpu:.3e}, fp32 oracle's own {e_orc:.3e}")
        else:
            worst_iq = max(worst_iq, err)
        if dm:
            d = port.demod(dm, bw, osr, ideal_nco=True)
            iso = np.concatenate([d.process(a) for a, _ in got[i]])
            gd = np.concatenate([q for _, q in got[i]])
            s = 1 if dm == po.DEMOD_QUAD else 0
            e = _rel(gd[s:], iso[s:])
            worst_dm = max(worst_dm, e)
            if e > TOL:
                failures.append(f"vfo {i}: demod {dm} stage-isolated {e:.3e}")
    res.update({"counts_exact": counts_ok, "worst_iq_rel_rms": worst_iq, "worst_demod_stage_isolated": worst_dm,
                "ref_f32_rotator_vs_ideal": worst_walk, "adjudicated": adjudicated})

    if with_fft:
        stream = np.concatenate(xs)
        N = w.fft_size
        grow = np.concatenate(rows) if rows else np.zeros((0, N), np.float32)
        nfr = len(stream) // N
        res["rows"] = int(grow.shape[0])
        if grow.shape[0] != nfr:
            failures.append(f"spectrum rows {grow.shape[0]} != {nfr}")
        win = port.window(w.fft_window, N)
        worst_db, worst_ratio = 0.0, 0.0
        for f in sorted(set([0, nfr - 1]))[:rows_to_check] if nfr > 0 and grow.shape[0] == nfr else []:
            row32, _, row64 = port.spectrum(N, stream[f * N:(f + 1) * N], win)
            mask = row64 >= row64.max() - 100.0
            dg = float(np.abs(grow[f].astype(np.float64) - row64)[mask].max())
            dr = float(np.abs(row32.astype(np.float64) - row64)[mask].max())
            worst_db = max(worst_db, dg)
            worst_ratio = max(worst_ratio, dg / max(dr, 1e-12))
            strict = float(np.abs(grow[f].astype(np.float64) - row64)[row64 >= row64.max() - 60.0].max())
            if strict > 0.01:
                failures.append(f"row {f}: {strict:.4f} dB within 60 dB of the peak")
            if dg > 0.01 and not (N >= (1 << 17) and dg <= 2.0 * dr):
                failures.append(f"row {f}: {dg:.4f} dB within 100 dB of the peak (oracle's own fp32 FFT: {dr:.4f} dB)")
        res.update({"worst_row_db_within_100dB": worst_db, "row_err_over_oracle_fp32_fft_err": worst_ratio})
    res["failures"] = failures
    res["ok"] = not failures
    return res
