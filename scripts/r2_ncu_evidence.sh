#!/bin/bash
# ncu evidence of round 2 (one B200 under gpurun): launch list of bench.py, and per-kernel DRAM / L2 / tensor-path
# metrics of one denoiser evaluation at B=256 (tcgen05-aware tensor metric: sm__ops_path_tensor_op_utchmma_*).
set -u
K='gemm_tcgen05|attention|router_kernel|plan_kernel|embed_kernel|ln2_permute|combine_kernel|head_kernel'
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tmem.sum,sm__cycles_elapsed.avg.per_second,launch__registers_per_thread,launch__grid_size'
ncu --metrics "$M" --clock-control none -k regex:"$K" -s 90 -c 90 --csv --log-file gpurun_out/r2h_ncu_kernel_metrics.csv \
    python scripts/profile_step.py --evals 2 --layers 12 > gpurun_out/r2h_ncu_metrics.stdout 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1800 --csv --log-file gpurun_out/r2h_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2h_ncu_bench.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05_2cta_kernel|combine_kernel|attention_tma_kernel' -s 20 -c 8 -o gpurun_out/r2h_prof_full -f \
    python scripts/profile_step.py --evals 2 --layers 3 > gpurun_out/r2h_prof_full.stdout 2>&1
ls -la gpurun_out/ | grep r2h
