#!/bin/bash
# Same-box A/B of the step.  Usage: bash tools/step_ab.sh OUT.log "ENV=.. ENV=.." "ENV=.." ...   (each argument one variant;
# "-" = defaults).  Two interleaved repetitions, headline only; per-layer lines of the refiner's heavy layers appended.
out=$1; shift
: > $out
n=0
for v in "$@"; do n=$((n+1)); done
for rep in 1 2; do
  i=0
  for v in "$@"; do
    i=$((i+1)); tag="V${i}r${rep}"
    if [ "$v" = "-" ]; then envs="MQ_NOOP=1"; else envs="$v"; fi
    env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --layer-table gpurun_out/layers_ab_${tag}.md 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$tag [$v]', 'ms_per_step', round(d['ms_per_step'],2), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['clocks']['reasons'])
" >> $out
  done
done
for f in gpurun_out/layers_ab_V*r2.md; do echo "== $f" >> $out; grep "ref.up0.conv1\|ref.mid.conv1\|ref.down1.conv2\|ref.up1.conv1\|ref.up2.conv1\|ref.down0.conv2\|ref.pre.conv2\|^total" $f >> $out; done
