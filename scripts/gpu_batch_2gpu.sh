timeout 600 python -m pytest tests/test_multigpu_equality_gpu.py -m gpu -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo bench2 rc=$?; tail -3 gpurun_out/r2_bench_2gpu.err
