#!/bin/bash
# Round-2 validation on one B200 (run through gpurun): GPU tests, smoke, sanitizer passes, both bench arms, ncu metric names.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2g_pytest.log 2>&1; tail -4 gpurun_out/r2g_pytest.log
grep -E "worst sampled-entry|full depth|engine vs|FiLM-ResNet|rg[0-9]+ " gpurun_out/r2g_pytest.log | cut -c1-260 > gpurun_out/r2g_parity_lines.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|utc|tmem" | cut -c1-160 > gpurun_out/r2g_ncu_tensor_metrics.txt; wc -l gpurun_out/r2g_ncu_tensor_metrics.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm_matches_fp32_reference and cta2 and 448-1024-1024" > gpurun_out/r2g_racecheck_gemm.log 2>&1; tail -5 gpurun_out/r2g_racecheck_gemm.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "fused_expert_mlp_kernel_is_bit_identical and tiny" > gpurun_out/r2g_racecheck_mlp_fused.log 2>&1; tail -5 gpurun_out/r2g_racecheck_mlp_fused.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py tests/test_resnet.py -m gpu -q -x -k "tile_widths and 300-1024-512 or s64" > gpurun_out/r2g_memcheck_new.log 2>&1; tail -5 gpurun_out/r2g_memcheck_new.log
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2g_bench_ref.json
