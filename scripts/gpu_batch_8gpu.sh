for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err; echo bench$n rc=$?; tail -2 gpurun_out/r2_bench_${n}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${n}gpu.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
PY
done
nvidia-smi topo -m 2>&1 | head -14; lscpu | grep -i "numa\|socket\|model name" | head
