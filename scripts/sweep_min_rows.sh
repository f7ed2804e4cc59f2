# K1 dispatch threshold sweep: stage time of the additive synthesis per config and SGB_SYNTH_MIN_ROWS
for c in 1 2 3; do for m in 0 128 224 320 448 100000; do
  SGB_SYNTH_MIN_ROWS=$m timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg$c min_rows $m synth %.2f total %.2f value %.0f' % (d['stage_ms']['synth'], d['ms_per_step'], d['value']))"
done; done
