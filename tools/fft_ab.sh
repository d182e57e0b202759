# A/B of two builds of the library on one box: spectrum tests on the new build, then batched timings and ncu per-launch
# durations of the FFT pair for both. "old" is a build of the commit to compare against, kept beside the new one:
#   git stash; python -m sdrpp_b200.build; cp sdrpp_b200/libsdrpp_cuda.so sdrpp_b200/libsdrpp_cuda_old.so; git stash pop; python -m sdrpp_b200.build
# (sdrpp_b200/cuda.py loads SDRPP_CUDA_LIB when it is set).  gpurun -- 'bash tools/fft_ab.sh > gpurun_out/ab.log 2>&1'
timeout 300 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_display.py -m gpu -x -q 2>&1 | tail -3
for lib in old new; do
  if [ $lib = old ]; then export SDRPP_CUDA_LIB=$PWD/sdrpp_b200/libsdrpp_cuda_old.so; else unset SDRPP_CUDA_LIB; fi
  echo "== $lib"
  python tools/fft_bench.py 20 18 20
  python tools/fft_bench.py 20 64 10
  python tools/fft_bench.py 18 64 20
  python tools/fft_bench.py 16 256 20
  python tools/fft_bench.py 16 1024 10
  ncu --metrics gpu__time_duration.sum,smsp__cycles_active.avg,smsp__inst_executed.sum --clock-control none -k regex:fft_ --csv --log-file gpurun_out/ab4_$lib.csv python tools/fft_bench.py 20 1 6 24 > /dev/null 2>&1
  ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:fft_ --csv --log-file gpurun_out/ab4w_$lib.csv python tools/fft_bench.py 20 1 6 1 > /dev/null 2>&1
done
python - <<'PY'
import csv, collections
for tag in ("ab4_old","ab4_new","ab4w_old","ab4w_new"):
    d=collections.defaultdict(list)
    rows=[r for r in csv.reader(open(f"gpurun_out/{tag}.csv")) if len(r)>10]
    h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
    for r in rows[1:]:
        d[(r[ki][:20], r[mi])].append(float(r[vi].replace(",","")))
    for k,v in d.items(): print(tag,k,"n=%d"%len(v),"min %.1f med %.1f"%(min(v),sorted(v)[len(v)//2]))
PY
