#!/bin/bash
# as ab_libs.sh, three alternations of 30 steps, with the per-kernel times
for rep in 1 2 3; do for l in ab_libs/*.so; do
  CBFSSM_B200_LIB=$PWD/$l python bench.py --no-extra --no-cpu-baseline --steps 30 --warmup 5 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done; done
