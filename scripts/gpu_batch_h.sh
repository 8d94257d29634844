timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_models_bf16_gpu.py tests/test_stage_abi_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/run_stem_e12_once.py 400 2>&1 | tail -2
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4_h.txt 2>&1; cat gpurun_out/r2_profile_ops_cfg4_h.txt | head -14; grep -A12 "halo family" gpurun_out/r2_profile_ops_cfg4_h.txt
