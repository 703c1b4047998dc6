#!/bin/bash
# 8-GPU evidence run (one box, `gpurun --gpus 8`): strong scaling of fixed global batches and the named multi-GPU
# configurations of BASELINE.json.  One JSON line per run under gpurun_out/scale8/.
set -u
O=gpurun_out/scale8; mkdir -p $O
run() {  # name nproc args...
  local name=$1 n=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --no-cpu-baseline "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:34s} N={d['n_gpus']} {d['scaling']:6s} {d['value']:.4g} p-steps/s  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.4g}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
# strong scaling: the global minibatch is fixed and split by particle index
for n in 2 4 8; do run strong_template_m100_n$n $n --scaling strong --no-extra --steps 20 --warmup 5; done
for n in 2 8; do run strong_spring_b32_n$n $n --workload spring_template_b32 --scaling strong --no-extra --steps 30 --warmup 5; done
run strong_sweep_1m_m100_n8 8 --workload sweep_1m_m100 --scaling strong --no-extra --steps 2 --warmup 1
# weak scaling of the named configurations at 8 GPUs (N = 1, 2, 4, 8 of the default line is the driver's SCALE run)
run weak_default_n8 8 --steps 10 --warmup 3
run weak_sarcos_m100_n8 8 --workload sarcos_m100 --no-extra --steps 5 --warmup 3
run weak_voliro_m20_n8 8 --workload voliro_m20 --no-extra --steps 10 --warmup 3
