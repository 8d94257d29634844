M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_tc_|k_convlstm|k_chunk|k_band|k_counts_norm|k_window' --csv --log-file gpurun_out/r2_traffic_cfg4.csv python scripts/ncu_traffic.py run cfg4 > gpurun_out/ncu4.log 2>&1; echo ncu4 rc=$?
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_chunk|k_band|k_window' --csv --log-file gpurun_out/r2_traffic_cfg2.csv python scripts/ncu_traffic.py run cfg2 > gpurun_out/ncu2.log 2>&1; echo ncu2 rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu_bench rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.err; echo bench rc=$?
