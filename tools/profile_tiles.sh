#!/bin/bash
# ncu --set full of the tensor-path message reverse kernel with one-tile and three-tile CTAs (batch 2272: whole
# waves of both).  Outputs under gpurun_out/tiles/.
set -u
O=gpurun_out/tiles; mkdir -p $O
for t in 1 3; do
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --batch 2272"
  CBFSSM_B200_TC_TILES=$t $CMD > $O/plain_$t.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$t.log; exit 1; }
  CBFSSM_B200_TC_TILES=$t ncu --clock-control none --set full --import-source on -k regex:bm_reverse_tc --launch-skip 4 -c 1 -f \
     -o $O/bm_reverse_tc_tiles$t $CMD > $O/ncu_$t.log 2>&1
done
ls -la $O
