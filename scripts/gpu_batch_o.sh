timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_o.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_o.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print(d['roofline']['frac'], d['roofline']['achieved'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
