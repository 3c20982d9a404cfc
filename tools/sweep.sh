#!/bin/bash
# usage: tools/sweep.sh out.log "--workload cfg2" "--workload cfg3" ...   (kernel-time table per workload)
out=$1; shift
mkdir -p gpurun_out
: > "$out"
for w in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $w 2>gpurun_out/sweep.err > gpurun_out/sweep.json
  python - "$w" >> "$out" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/sweep.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 2),
          {k: round(v["avg_ms"], 3) for k, v in d["kernels"].items() if v["avg_ms"] > 0.02})
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sweep.err").read()[-400:])
PY
done
cat "$out"
