#!/bin/bash
# sweep forward-kernel knobs on the GPU box; prints per-kernel avg ms
for st in 2 3 4; do
  DCN_FWD_STAGES=$st python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('stages=$st', {k:round(v['avg_ms'],3) for k,v in d['kernels'].items() if 'umma' in k or 'nhwc' in k})"
done
