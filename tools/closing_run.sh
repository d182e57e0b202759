# closing run of the round (the E16 leg needs sdrpp_b200/libsdrpp_cuda_e16.so: SDRPP_EXTRA_NVCC=-DSDRPP_FFT1024_E16 python -m sdrpp_b200.build, copied aside): whole GPU suite, the default bench line, the E16 spectrum variant, ncu launch list, ncu full of the FFT pair
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/t40.log 2>&1; tail -2 gpurun_out/t40.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2x_bench_default_20_steps.json 2> gpurun_out/r2x_default.err; python -c "
import json;d=json.load(open('gpurun_out/r2x_bench_default_20_steps.json'));print('bench', round(d['value']), round(d['e2e']['value']), d.get('parity_check',{}).get('ok'), d.get('roofline_spectrum'))"
export SDRPP_CUDA_LIB=$PWD/sdrpp_b200/libsdrpp_cuda_e16.so
python tools/fft_bench.py 20 18 20
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ --csv --log-file gpurun_out/ab5_e16.csv python tools/fft_bench.py 20 1 6 24 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:fft_ --csv --log-file gpurun_out/ab5w_e16.csv python tools/fft_bench.py 20 1 6 1 > /dev/null 2>&1
unset SDRPP_CUDA_LIB
python - <<'PY'
import csv, collections
for tag in ("ab5_e16","ab5w_e16"):
    d=collections.defaultdict(list)
    rows=[r for r in csv.reader(open(f"gpurun_out/{tag}.csv")) if len(r)>10]
    h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
    for r in rows[1:]:
        d[(r[ki][:20], r[mi])].append(float(r[vi].replace(",","")))
    for k,v in d.items(): print(tag,k,"n=%d"%len(v),"min %.1f med %.1f"%(min(v),sorted(v)[len(v)//2]))
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 8 --warmup 3 --graph-warmup 0 --no-cpu-baseline --no-cpp --no-parity > gpurun_out/r2x_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 90 ncu --set full --import-source on --clock-control none -k regex:fft_ -c 4 -o gpurun_out/r2x_fft python tools/fft_bench.py 20 1 2 24 > gpurun_out/r2x_ncu2.log 2>&1; echo "full rc=$?"
