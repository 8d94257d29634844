"""GPU, >= 2 devices: shard = single-GPU result on hardware (SURVEY section 4 layer 6). Two ranks (torchrun, NCCL)
evaluate a fixed set of trajectories through evfly_b200.sharding.evaluate_trajectories -- rank r takes trajectories
r::G with private recurrent state, one all_gather of the velocity commands -- and rank 0 compares the gathered result
bit for bit with its own evaluation of ALL trajectories (learner/evaluation_tools.py:62-66 runs them one by one on one
device). Also run on one GPU, where the sharded path degenerates to the plain loop."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    if world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "equality", "--gpus", "1"]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--workload", "equality", "--gpus", str(world)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_single_gpu_sharded_path_equals_plain_loop(cuda_lib):
    d = _run(1)
    assert d["equal_bit_for_bit"] and d["finite"] and d["n_trajectories"] == 8


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box (gpurun --gpus 2)")
def test_two_ranks_equal_one_gpu_bit_for_bit(cuda_lib):
    d = _run(2)
    assert d["n_gpus"] == 2 and d["equal_bit_for_bit"] and d["finite"], d
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(d, open(os.path.join(ROOT, "gpurun_out", "equality_2gpu.json"), "w"))
