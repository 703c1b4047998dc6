# every bench workload with and without NaN-poisoned workspaces: the losses must be finite and identical
for w in robomove_m20 template_m100 sarcos_m100 voliro_m20 sweep_d8_m100 spring_template_b32; do
  for p in "" 1; do
    CBFSSM_B200_POISON_WS=$p timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w', 'poison=$p', 'loss', repr(d['e2e']['loss']), 'ms', round(d['ms_per_step'],2))"
  done
done
