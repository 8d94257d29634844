timeout 240 python -m pytest tests/test_vit_fused_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_a.err
timeout 600 python -m pytest tests/test_multigpu_equality_gpu.py tests/test_bench_shape_parity_gpu.py tests/test_feeder_gpu.py tests/test_trajectories_gpu.py tests/test_models_bf16_gpu.py tests/test_streaming_gpu.py -m gpu -q 2>&1 | tail -15
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_cfg4.csv python scripts/ncu_traffic.py run cfg4 > gpurun_out/ncu4.log 2>&1; echo ncu4 rc=$?
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_cfg2.csv python scripts/ncu_traffic.py run cfg2 > gpurun_out/ncu2.log 2>&1; echo ncu2 rc=$?
