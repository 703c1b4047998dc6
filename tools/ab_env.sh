#!/bin/bash
# A/B of an environment setting on one box:  tools/ab_env.sh VAR "v1 v2 ..." [bench args]
var=$1; vals=$2; shift 2
for rep in 1 2; do for v in $vals; do
  env $var=$v python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['ms_per_step'],3), {k:round(x,2) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done; done
