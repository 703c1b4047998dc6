#!/bin/bash
# Round-end evidence run (one B200): full GPU test suite, the two bench workloads, ncu launch lists and
# --set full captures of the dominant kernels at the bench batch sizes. Outputs under gpurun_out/final/.
set -u
O=gpurun_out/final4; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python bench.py > $O/bench_m20.json 2> $O/bench_m20.err; tail -c 600 $O/bench_m20.json
python bench.py --workload template_m100 --no-cpu-baseline > $O/bench_m100.json 2> $O/bench_m100.err; tail -c 600 $O/bench_m100.json
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_m20.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_m20.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_m100.csv \
   python bench.py --workload template_m100 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_m100.log 2>&1
for k in bm_reverse_fast fw_reverse_fast; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o $O/full_m20_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/full_m20_$k.log 2>&1
done
for k in bm_reverse_tc fw_reverse_tc tc_outer; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o $O/full_m100_$k \
     python bench.py --workload template_m100 --steps 1 --warmup 3 --no-cpu-baseline > $O/full_m100_$k.log 2>&1
done
ls -la $O
