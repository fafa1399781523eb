#!/bin/bash
# Round-2 ncu evidence, run on the GPU box:  gpurun -- 'bash profiles/capture_r2.sh'
# (numbers printed by anything running under ncu are never bench values)
set -x
mkdir -p gpurun_out
K='regex:conv3x3_tc_kernel|wgrad_tc_kernel|head_gather_kernel|head_scatter_kernel'
# 1. launch list of ONE eager training iteration of the default workload (c2)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_train_step_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-graph --roof-steps 1 \
    > gpurun_out/r2_launchlist_bench.json 2> gpurun_out/r2_launchlist_bench.err
# 2. ncu --set full of every tensor-core kernel form at the top block's shapes
ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" \
    -o gpurun_out/r2_kernels_full_256 -f python profiles/kernels_microbench_r2.py --no-warm \
    > gpurun_out/r2_kernels_256_under_ncu.log 2>&1
ncu -i gpurun_out/r2_kernels_full_256.ncu-rep --page raw --csv | python profiles/summarize_ncu.py \
    > gpurun_out/r2_kernels_ncu_full_256.csv
ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" \
    -o gpurun_out/r2_kernels_full_512 -f python profiles/kernels_microbench_r2.py --no-warm --sub --S 512 --B 2 \
    > gpurun_out/r2_kernels_512_under_ncu.log 2>&1
ncu -i gpurun_out/r2_kernels_full_512.ncu-rep --page raw --csv | python profiles/summarize_ncu.py \
    > gpurun_out/r2_kernels_ncu_full_512.csv
# 3. the same kernels timed with CUDA events, no profiler
python profiles/kernels_microbench_r2.py --reps 5 > gpurun_out/r2_kernels_256_events.log 2>&1
python profiles/kernels_microbench_r2.py --reps 5 --sub --S 512 --B 2 > gpurun_out/r2_kernels_512_events.log 2>&1
rm -f gpurun_out/r2_kernels_full_256.ncu-rep gpurun_out/r2_kernels_full_512.ncu-rep
ls -la gpurun_out | tail -12
