#!/bin/bash
# Throughput of one bench workload against the per-GPU batch (waves of resident CTAs): bash tools/batch_sweep.sh <workload> b1 b2 ...
W=$1; shift
for b in "$@"; do
  python bench.py --workload $W --batch $b --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['global_batch'], '%.4g' % d['value'], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()}, round(d['l2_working_set_mib']['workspace']))"
done
