timeout 300 python -m pytest tests/test_stage_abi_gpu.py tests/test_models_bf16_gpu.py tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -8
python scripts/run_stem_e12_once.py 64 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_stem_e12 -s 3 -c 1 -o gpurun_out/r2_stem_e12_v2 python scripts/run_stem_e12_once.py 16 > gpurun_out/ncu_stem.log 2>&1; echo ncu_stem rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_d.err
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4.txt 2>&1; echo prof rc=$?
