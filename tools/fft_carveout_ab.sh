# A/B of the spectrum kernels' shared-memory carve-out preference (SDRPP_FFT_CARVEOUT), one box
for c in max default; do
  export SDRPP_FFT_CARVEOUT=$c
  echo "== carveout $c"
  python tools/fft_bench.py 20 18 20
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ --csv --log-file gpurun_out/ab6_$c.csv python tools/fft_bench.py 20 1 6 24 > /dev/null 2>&1
  timeout 100 python bench.py --steps 600 --warmup 20 --no-cpu-baseline --no-cpp --no-parity > gpurun_out/ab6_bench_$c.json 2> gpurun_out/ab6_bench_$c.err
  python -c "
import json;d=json.load(open('gpurun_out/ab6_bench_$c.json'));print('bench', round(d['value']), round(d['e2e']['value']), d['kernel_ms_per_step'], round(d['spectrum_batched']['msps']))"
done
export SDRPP_FFT_CARVEOUT=50
echo "== carveout 50"; python tools/fft_bench.py 20 18 20
python - <<'PY'
import csv, collections
for tag in ("ab6_max","ab6_default"):
    d=collections.defaultdict(list)
    rows=[r for r in csv.reader(open(f"gpurun_out/{tag}.csv")) if len(r)>10]
    h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
    for r in rows[1:]:
        d[(r[ki][:20], r[mi])].append(float(r[vi].replace(",","")))
    for k,v in d.items(): print(tag,k,"n=%d"%len(v),"min %.1f med %.1f"%(min(v),sorted(v)[len(v)//2]))
PY
