timeout 600 python -m pytest tests/test_accumulate_tiles_gpu.py tests/test_trajectories_gpu.py tests/test_multigpu_equality_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_k.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_k.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print(d['roofline']['frac'], d['roofline']['achieved'], d['rooflines_other'] if 'rooflines_other' in d else '')
PY
