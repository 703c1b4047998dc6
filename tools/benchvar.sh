# usage: bash tools/benchvar.sh [extra bench args]; compares libcbf_v*.so variants against the default library
for lib in "" $(ls cbf_ssm_b200/libcbf_v*.so 2>/dev/null); do
  echo "== lib=$lib"
  CBFSSM_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step']); print(d['roofline']['kernel_ms_avg'])"
done
