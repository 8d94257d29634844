timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python scripts/run_stem_e12_once.py 64 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_f.err
timeout 300 python scripts/profile_ops.py trajectories > gpurun_out/r2_profile_ops_cfg4.txt 2>&1; echo prof rc=$?
