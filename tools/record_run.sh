#!/bin/bash
# Record run for profiles/ (round 2): default bench (with detector_dp), reference arm, ncu launch list and full captures
# of the dominant kernels of every BASELINE config (each ncu pass only after the same command exited 0 without ncu).
# Run on the GPU box through gpurun; tools/refresh_profiles.py r2 then copies the judged formats into profiles/.
set -x
mkdir -p gpurun_out /tmp/dcn_prof
O=gpurun_out
R=/tmp/dcn_prof      # the .ncu-rep files stay on the box (gpurun brings back at most 64 MiB): they are summarised there
NCU="ncu --set full --clock-control none --import-source on"
python bench.py > $O/bench_final.json 2> $O/bench_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_reference.json 2> $O/bench_final_reference.err
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detector-dp > $O/plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detector-dp > $O/ncu_launches_final.log 2>&1
$NCU -k regex:"bwd_data_kernel|umma_gemm_kernel|gout_tiles|nchw_to_nhwc|nhwc_to_nchw" -c 6 \
    -o $R/prof_final2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detector-dp > $O/ncu_final2.log 2>&1
# whole layer (offset conv on the engine)
python bench.py --scope layer --no-cpu-baseline --no-detector-dp --steps 10 > $O/bench_layer.json 2> $O/bench_layer.err
$NCU -k regex:"conv_kernel" -c 3 -o $R/prof_layer_conv -f \
    python bench.py --scope layer --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-detector-dp > $O/ncu_layer.log 2>&1
# configs[2]
for v in jittor torch; do
  python bench.py --workload cfg3 --variant $v --steps 20 > $O/bench_cfg3_$v.json 2> $O/bench_cfg3_$v.err
  $NCU -k regex:"bwd_data_kernel|umma_gemm_kernel" -c 3 -o $R/prof_cfg3_$v -f \
      python bench.py --workload cfg3 --variant $v --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_cfg3_$v.log 2>&1
done
# configs[3]
for v in jittor torch; do for op in bf16 fp32; do
  python bench.py --workload stack --variant $v --operand $op --steps 10 > $O/bench_stack_${v}_$op.json 2> $O/bench_stack_${v}_$op.err
done; done
$NCU -k regex:"bwd_data_kernel|umma_gemm_kernel" -c 26 -o $R/prof_stack_jittor_bf16 -f \
    python bench.py --workload stack --variant jittor --operand bf16 --steps 1 --warmup 0 > $O/ncu_stack.log 2>&1
# configs[4] / configs[0]
python bench.py --workload detector --steps 20 --warmup 5 > $O/det_1gpu_final.json 2> $O/det_final.err
python bench.py --workload detector --global-batch 16 --steps 50 --warmup 10 > $O/det_b16_graph.json 2>> $O/det_final.err
# single layers of the stack and of the detector
for w in c3 c4 c5 det2 det5; do for v in jittor torch; do
  python bench.py --workload $w --variant $v --no-cpu-baseline --no-e2e --steps 10 > $O/layer_${w}_$v.json 2>/dev/null
done; done
# summaries (text) of everything above, in the judged formats, next to the bench lines
mkdir -p $O/profiles_r2
python tools/refresh_profiles.py r2 $O $R $O/profiles_r2 > $O/refresh.log 2>&1
echo done
