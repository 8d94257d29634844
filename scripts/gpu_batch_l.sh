M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_tc_|k_convlstm|k_chunk|k_band|k_counts_norm|k_window' --csv --log-file gpurun_out/r2_traffic_cfg4.csv python scripts/ncu_traffic.py run cfg4 > gpurun_out/ncu4.log 2>&1; echo ncu4 rc=$?
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_chunk|k_band|k_window' --csv --log-file gpurun_out/r2_traffic_cfg2.csv python scripts/ncu_traffic.py run cfg2 > gpurun_out/ncu2.log 2>&1; echo ncu2 rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_l.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_l.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print(d['roofline']['frac'], d['roofline']['achieved'])
PY
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4.txt 2>&1; echo prof rc=$?
