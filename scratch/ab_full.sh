#!/bin/bash
# bench.py headline + extra configs for the default build and every scratch/variants/*.so: clusterer ms per config
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-14s zed720 step %.4f sf %.4f cl %.4f |' % ('$1', d['ms_per_step'], d['roofline']['kernel_ms'], d['clusterer']['ms']), ' | '.join('%s cl %.3f sf %.3f' % (e['config'].split('_')[0], e['cl_ms'], e['sf_ms']) for e in d['extra']['configs']), '| strong %.3f ms' % d['strong_scaling_config3']['ms_per_batch'])"; }
for rep in 1 2; do
python bench.py --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | show default
for so in scratch/variants/*.so; do MOD_B200_LIB=$so python bench.py --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | show $(basename $so .so); done
done
