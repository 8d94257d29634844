timeout 300 python -m pytest tests/test_accumulate_tiles_gpu.py tests/test_tc_gpu.py tests/test_accumulate_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_b.err
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4.txt 2>&1; echo prof rc=$?
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:'k_tc_|k_chunk|k_band|k_counts_norm|k_window' --csv --log-file gpurun_out/r2_traffic_cfg4.csv python scripts/ncu_traffic.py run cfg4 > gpurun_out/ncu4.log 2>&1; echo ncu4 rc=$?
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_chunk|k_band|k_window' --csv --log-file gpurun_out/r2_traffic_cfg2.csv python scripts/ncu_traffic.py run cfg2 > gpurun_out/ncu2.log 2>&1; echo ncu2 rc=$?
