timeout 400 python -m pytest tests/test_stage_abi_gpu.py tests/test_models_bf16_gpu.py tests/test_tc_gpu.py tests/test_trajectories_gpu.py tests/test_feeder_gpu.py tests/test_streaming_gpu.py -m gpu -q 2>&1 | tail -12
python scripts/run_stem_e12_once.py 64 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_e.err
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4.txt 2>&1; echo prof rc=$?
