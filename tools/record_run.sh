#!/bin/bash
# Record run for profiles/: default bench, reference arm, ncu launch list and one full capture of the
# two dominant kernels (each ncu pass only after the same command exited 0 without ncu).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bwd_data_kernel|umma_gemm_kernel|gout_tiles|nchw_to_nhwc|nhwc_to_nchw" -c 6 \
    -o gpurun_out/prof_final2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_final2.log 2>&1
echo done
