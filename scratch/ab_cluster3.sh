#!/bin/bash
run() { local label=$1; shift
  env "$@" python bench.py --no-extra --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s pairs/s %8.0f  step %.4f ms  sf %.4f  cl %.4f  objs %d' % ('$label', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clusterer']['ms'], d['objects_per_step']))"
}
for pf in 0 600 1184 2368 4736 0; do
run "tile z prefetch $pf" MOD_B200_CL_PF=$pf
done
