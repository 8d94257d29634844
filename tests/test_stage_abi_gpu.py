"""GPU: the stage-level C ABI driven from PLAIN C (tests/native/stage_driver.c: no Python, no torch in the process) gives
bit for bit what the Python drop-in module gives -- the packed weights travel through a file, the C program calls
evfly_unet_forward twice (fresh state, then the carried state) the way the C++ side of evfly_ros would."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

import evfly_b200
from evfly_b200.pipeline import build_deployed_model
from oracle.synth_ckpt import shapes_of, synth_state_dict, synthetic_frames

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "native", "stage_driver")


def test_c_program_drives_unet_forward(cuda_lib, tmp_path):
    assert os.path.exists(DRIVER), "tests/native/stage_driver is built by __graft_entry__.build()"
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), "bf16")
        N, n_traj, H, W = 4, 2, 260, 346
        frames = synthetic_frames(21, N)
        stage, keep = m.origunet.packed()["stage"]
        blob = tmp_path / "unet.bin"
        with open(blob, "wb") as f:
            f.write(struct.pack("<4if", N, n_traj, H, W, float(m.origunet.evs_min_cutoff)))
            for t in keep + [frames]:
                raw = t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes()
                f.write(struct.pack("<q", len(raw)))
                f.write(raw)
        out = tmp_path / "out.bin"
        r = subprocess.run([DRIVER, str(blob), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "kernel launches" in r.stdout
        got = np.fromfile(out, dtype=np.float32)
        # the same two forwards through the Python module
        dv = torch.full((N, 1), 4.0, device="cuda")
        _, (d1, _, (hu, _)) = m.origunet.forward([frames.clone().cuda(), dv, None], n_traj=n_traj)
        fr2 = frames.clone().cuda()
        fr2[fr2.abs() < m.origunet.evs_min_cutoff] = 0.0            # the C program reuses the frames the first call cut off in place
        _, (d2, _, _) = m.origunet.forward([fr2, dv, [hu, None]], n_traj=n_traj)
        want = torch.cat([d1.flatten(), hu[0][0].flatten(), hu[0][1].flatten(), d2.flatten()]).cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(got, want)
