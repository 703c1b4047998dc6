#!/bin/bash
# Headline workload with one-tile and three-tile CTAs of the tensor path, at batches that fill whole waves of each.
for b in 1515 2272 1136; do for t in 1 3; do
  CBFSSM_B200_TC_TILES=$t python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 --batch $b 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch',$b,'tiles',$t,'ms',round(d['ms_per_step'],3),'rate %.4g'%d['value'])"
done; done
