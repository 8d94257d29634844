timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_models_bf16_gpu.py tests/test_stage_abi_gpu.py tests/test_bench_shape_parity_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4_m.txt 2>&1; head -12 gpurun_out/r2_profile_ops_cfg4_m.txt; grep -A11 "halo family" gpurun_out/r2_profile_ops_cfg4_m.txt
