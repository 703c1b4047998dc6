# 8-GPU weak-scaling check of both bench workloads (run with: gpurun --gpus 8 -- 'bash tools/bench_scale8.sh')
mkdir -p gpurun_out/scale
for wl in robomove_m20 template_m100; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 8 --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2> gpurun_out/scale/err_$wl.log | tail -1 > gpurun_out/scale/n8_$wl.json
  python -c "
import json;d=json.load(open('gpurun_out/scale/n8_$wl.json'));print('$wl', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/scale/err_$wl.log
done
