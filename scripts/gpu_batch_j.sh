timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_j.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_j.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print(d['roofline']['frac'], d['roofline']['achieved'])
PY
