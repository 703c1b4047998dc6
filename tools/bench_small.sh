# step time at the reference's own minibatch sizes (latency-bound regime), per kernel path
for spec in "robomove_m20 32 0" "robomove_m20 32 1" "template_m100 32 0" "template_m100 32 2" "sarcos_m100 5 0" "sarcos_m100 5 2"; do
  set -- $spec
  python bench.py --workload $1 --batch $2 --flags $3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:40], 'flags=$3 B=',d['config']['batch_per_gpu'], 'N=',d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'psteps/s %.3g'%d['value']); print('   ', {k:round(v,3) for k,v in d['roofline']['kernel_ms_avg'].items()})"
done
