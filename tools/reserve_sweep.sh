for r in 8 40 74 100; do
  SDRPP_RESERVE_SMS=$r python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$r bench.py --gpus 2 --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/sw_n2_$r.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw_n2_$r.json'));print('N=2 reserve',$r,round(d['value']),round(d['e2e']['value']),d['kernel_ms_per_step'])"
done
for r in 16 32 48; do
  SDRPP_RESERVE_SMS=$r python bench.py --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/sw_n1_$r.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw_n1_$r.json'));print('N=1 reserve',$r,round(d['value']),round(d['e2e']['value']),d['kernel_ms_per_step'])"
done
