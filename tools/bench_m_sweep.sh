# cooperative vs tensor path for M without a register instantiation (libcbf_v1.so = tensor path from M = 16)
for M in 8 12 16; do for lib in "" cbf_ssm_b200/libcbf_v1.so; do
CBFSSM_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --workload template_m100 --M $M --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('M', $M, 'lib', '$lib' or 'default', 'ms', round(d['ms_per_step'],2), 'psteps/s %.3g' % d['value'], len(d['roofline']['kernel_ms_avg']))"
done; done
