#!/bin/bash
run() { local label=$1; shift
  env "$@" python bench.py --no-extra --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s pairs/s %8.0f  step %.4f ms  sf %.4f  cl %.4f  objs %d' % ('$label', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clusterer']['ms'], d['objects_per_step']))"
}
run "bytes sub1" MOD_B200_CL_TILE=bytes MOD_B200_CL_SUBBATCHES=1
for t in 1 2 4; do for k in 2 4; do
run "bytes sub$k tail$t" MOD_B200_CL_TILE=bytes MOD_B200_CL_SUBBATCHES=$k MOD_B200_CL_TAIL_CTAS=$t
done; done
run "bytes sub1" MOD_B200_CL_TILE=bytes MOD_B200_CL_SUBBATCHES=1
