# crossover between cooperative and tensor-core kernels at M=100
for spec in "template_m100 64 2" "template_m100 64 8" "template_m100 128 2" "template_m100 128 8" "template_m100 256 2" "template_m100 256 8" "sarcos_m100 512 2" "sarcos_m100 512 8" "sarcos_m100 2048 0"; do
  set -- $spec
  timeout 200 python bench.py --workload $1 --batch $2 --flags $3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:30], 'flags=$3 N=',d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],3), 'psteps/s %.3g'%d['value'], {k:round(v,2) for k,v in d['roofline']['kernel_ms_avg'].items()})"
done
