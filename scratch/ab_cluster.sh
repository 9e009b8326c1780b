#!/bin/bash
# same-run A/B of the clusterer's tile-kernel variants and sub-batch pipeline depths (bench headline config)
run() { # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-extra --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s pairs/s %8.0f  step %.4f ms  sf %.4f  cl %.4f  objs %d' % ('$label', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clusterer']['ms'], d['objects_per_step']))"
}
for rep in 1 2; do
run "bytes sub1" MOD_B200_CL_TILE=bytes MOD_B200_CL_SUBBATCHES=1
run "bits sub1" MOD_B200_CL_SUBBATCHES=1
run "bits sub2" MOD_B200_CL_SUBBATCHES=2
run "bits sub4" MOD_B200_CL_SUBBATCHES=4
run "bits sub7" MOD_B200_CL_SUBBATCHES=7
run "dense sub4" MOD_B200_CL_TILE=dense MOD_B200_CL_SUBBATCHES=4
done
