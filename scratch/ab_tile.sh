#!/bin/bash
# same-run A/B: list-driven tile kernel vs raster-scan kernel (bench headline config + the other configs)
run() { local label=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-e2e --steps 200 --strong-passes 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ex=' | '.join('%s cl %.3f sf %.3f' % (c['config'].split('_')[0], c['cl_ms'], c['sf_ms']) for c in d['extra']['configs'])
print('%-10s zed720 step %.4f sf %.4f cl %.4f objs %d | %s | strong %.3f ms' % ('$label', d['ms_per_step'], d['roofline']['kernel_ms'], d['clusterer']['ms'], d['objects_per_step'], ex, d['strong_scaling_config3']['ms_per_batch']))"
}
for rep in 1 2; do
run lists MOD_B200_CL_TILE=lists
run scan MOD_B200_CL_TILE=scan
done
