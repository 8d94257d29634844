"""Per-stage CUDA-event timing of the pipeline (exploration; bench.py is the contract bench)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import _lib, ops  # noqa: E402
from evfly_b200.events import to_device  # noqa: E402
from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model  # noqa: E402
from evfly_b200.synthetic import synthetic_stream  # noqa: E402
from oracle.synth_ckpt import shapes_of, synth_state_dict  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.set_grad_enabled(False)
    m = build_deployed_model("cpu")
    m.load_state_dict(synth_state_dict(shapes_of(m), 31))
    m = m.cuda()
    import evfly_b200
    evfly_b200.set_precision(m, os.environ.get("EVFLY_PRECISION", "fp32"))
    pipe = PerceptionPipeline(m, sensor_hw=(260, 346))
    rec, edges = synthetic_stream(0, T, 100_000, 260, 346)
    d, de = to_device(rec), torch.from_numpy(edges).cuda()
    res = {"T": T}
    res["L1+L2_ms"] = ev_time(lambda: pipe.frames_from_windows(d, de))
    frames, _, _ = pipe.frames_from_windows(d, de)
    res["forward_ms"] = ev_time(lambda: pipe.forward(frames.clone(), carry_state=False))
    u = m.origunet
    res["unet_ms"] = ev_time(lambda: u([frames.clone(), None, None]))
    depth = torch.rand((T, 1, 260, 346), device="cuda")
    dv = torch.full((T, 1), 4.0, device="cuda")
    res["vitlstm_ms"] = ev_time(lambda: m.vitfly_vitlstm([depth, dv, None, None]))
    t0 = time.perf_counter()
    for _ in range(3):
        pipe.forward(frames.clone(), carry_state=False)
    torch.cuda.synchronize()
    res["forward_wall_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
