#!/bin/bash
# Runs under gpurun (1 GPU). Produces the launch list and full captures of the hot kernels in gpurun_out/.
set -x
K='gemm_tcgen05|attention_kernel|router_kernel|plan_kernel|embed_kernel|ln2_permute|combine_kernel|head_kernel|cast_bf16'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --evals 2 --layers 12 > gpurun_out/launches.stdout 2>&1
# full capture: 3 layers are enough (same kernels), second evaluation (skip the first eval's launches of each kernel)
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 28 -c 26 -o gpurun_out/prof_step -f \
    python scripts/profile_step.py --evals 2 --layers 3 > gpurun_out/prof_step.stdout 2>&1
ls -la gpurun_out/
