#!/bin/bash
# Round-2 evidence run (one B200): ncu launch list of the default bench command and --set full captures of the
# dominant kernels at the bench batch sizes.  Outputs under gpurun_out/r02prof/.
set -u
O=gpurun_out/r02prof; mkdir -p $O
NCU="ncu --clock-control none"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
$NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $O/launches_default.csv $CMD > $O/ncu_launches.log 2>&1
for k in bm_reverse_tc fw_reverse_tc tc_outer bm_forward_tc; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o $O/full_m100_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/full_m100_$k.log 2>&1
done
for k in bm_reverse_fast fw_reverse_fast bm_forward_fast; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o $O/full_m20_$k \
     python bench.py --workload robomove_m20 --steps 1 --warmup 3 --no-cpu-baseline > $O/full_m20_$k.log 2>&1
done
ls -la $O | head -30
