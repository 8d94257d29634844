timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tc_stem_e12 -s 3 -c 1 -o gpurun_out/r2_stem_e12_final python scripts/run_stem_e12_once.py 64 > gpurun_out/ncu_stem.log 2>&1; echo ncu_stem rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_convlstm_scan -s 2 -c 1 -o gpurun_out/r2_convlstm_scan python scripts/exp_scan_timeline.py > gpurun_out/ncu_scan.log 2>&1; echo ncu_scan rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'^k_tc_conv3x3_halo$' -s 15 -c 1 -o gpurun_out/r2_halo_e22 python scripts/ncu_traffic.py run cfg4 > gpurun_out/ncu_e22.log 2>&1; echo ncu_e22 rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu_bench rc=$?; wc -l gpurun_out/r2_launches_bench_cfg4.csv
ls -la gpurun_out/*.ncu-rep | tail -4
