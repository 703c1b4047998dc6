# quick check of the tensor path: parity tests + the M=100 bench line for the default library and any libcbf_v*.so
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for lib in "" $(ls cbf_ssm_b200/libcbf_v*.so 2>/dev/null); do
echo "== lib=$lib"
CBFSSM_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --workload template_m100 --steps 5 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['batch_per_gpu'], d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['kernel_ms_avg'])"
done
