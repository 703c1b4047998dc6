#!/bin/bash
# A/B of kernel builds on one box: every ab_libs/*.so runs the headline bench (alternating, twice).
for rep in 1 2; do for l in ab_libs/*.so; do
  CBFSSM_B200_LIB=$PWD/$l python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', round(d['ms_per_step'],3))"
done; done
