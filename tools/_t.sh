timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_oracle.py -x -q -m gpu 2>&1 | tail -2
for k in 0 1; do
if [ $k = 1 ]; then export DCN_FWD_NO_SPLIT88=1; fi
for w in "cfg3 torch fp32" "c4 torch fp32"; do
set -- $w
python bench.py --workload $1 --variant $2 --operand $3 --no-e2e --no-cpu-baseline --no-detector-dp --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/t.json
python - <<P
import json
d=json.load(open("gpurun_out/t.json"))
print("nosplit=$k $w", round(d["ms_per_step"],3), {k:round(v["avg_ms"],3) for k,v in d["kernels"].items() if "umma" in k})
P
done; done
