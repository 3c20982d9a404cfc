#!/bin/bash
# timing experiments: knock out one component of the forward kernel at a time (results invalid)
for d in 0 1 2 4 8 16 3 7 31; do
  DCN_FWD_DBG=$d python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dbg=$d', {k:round(v['avg_ms'],3) for k,v in d['kernels'].items() if 'umma' in k})"
done
