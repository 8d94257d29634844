"""GPU: the drop-in nn.Modules (fp32 exact path, through the C ABI) against the oracle run live
on the CPU and against the golden vectors produced by the reference's own modules.
Tolerance: rtol 1e-5 (north_star fp32 path) with atol 1e-5 for re-association noise."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as M
from oracle.synth_ckpt import synth_state_dict, synthetic_depth, synthetic_frames
from tests.test_models_cpu import build

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield

RTOL, ATOL = 1e-5, 1e-5


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "models_golden.npz"))


@pytest.fixture(scope="module")
def manifest(golden_dir):
    return json.load(open(os.path.join(golden_dir, "state_dict_manifest.json")))


def close(got, want, rtol=RTOL, atol=ATOL):
    want = want.numpy() if isinstance(want, torch.Tensor) else want
    np.testing.assert_allclose(got.detach().cpu().numpy(), want, rtol=rtol, atol=atol)


def load(name, manifest, seed):
    m = build(name)
    m.load_state_dict(synth_state_dict(manifest[name], seed), strict=True)
    return m.cuda().eval().float()


DEPTH = synthetic_depth(1, 3)
DESVEL = torch.tensor([[4.0], [5.5], [3.0]])
QUAT = torch.tensor([[1.0, 0, 0, 0], [0.9, 0.1, -0.2, 0.3], [0.7, 0.0, 0.7, 0.1]])


@pytest.mark.parametrize("name,seed", [("LSTMNetVIT", 11), ("LSTMNet", 14), ("UNetConvLSTMNet", 15)])
def test_stateful_vitfly_models_vs_golden(cuda_lib, G, manifest, name, seed):
    m = load(name, manifest, seed)
    vel, h = m([DEPTH.clone().cuda(), DESVEL.cuda(), QUAT.clone().cuda()])
    close(vel, G[f"{name}_vel"]); close(h[0], G[f"{name}_h"]); close(h[1], G[f"{name}_c"])
    vel2, h2 = m([DEPTH.flip(0).clone(), DESVEL, None, h])            # CPU inputs are moved, state carried
    close(vel2, G[f"{name}_vel2"]); close(h2[0], G[f"{name}_h2"])


@pytest.mark.parametrize("name,seed", [("ViT", 12), ("ConvNet", 13)])
def test_stateless_vitfly_models_vs_golden(cuda_lib, G, manifest, name, seed):
    m = load(name, manifest, seed)
    vel, h = m([DEPTH.clone().cuda(), DESVEL.cuda(), QUAT.clone().cuda()])
    assert h is None
    close(vel, G[f"{name}_vel"])
    close(m([DEPTH.flip(0).clone().cuda(), DESVEL.cuda(), None])[0], G[f"{name}_vel2"])


def test_lstmnetvit_resize_and_sequence_semantics(cuda_lib, G, manifest):
    m = load("LSTMNetVIT", manifest, 11)
    X = [torch.from_numpy(G["LSTMNetVIT_big_in"]).cuda(), DESVEL[:2].cuda(), None]
    close(m(X)[0], G["LSTMNetVIT_big_vel"])
    assert X[0].shape[-2:] == (60, 90) and X[2].shape == (2, 4)       # refine_inputs mutates the list like the reference
    # one call over N frames == N calls carrying the state
    vel, h = m([DEPTH.clone().cuda(), DESVEL.cuda(), QUAT.cuda()])
    st, outs = None, []
    for t in range(3):
        v, st = m([DEPTH[t:t + 1].clone().cuda(), DESVEL[t:t + 1].cuda(), QUAT[t:t + 1].cuda(), st])
        outs.append(v)
    close(torch.cat(outs), vel.cpu()); close(st[0], h[0].cpu()); close(st[1], h[1].cpu())


def test_mix_transformer_stage_vs_oracle(cuda_lib, manifest):
    m = load("LSTMNetVIT", manifest, 11)
    sd = synth_state_dict(manifest["LSTMNetVIT"], 11)
    s1 = M.mix_transformer_stage(sd, "encoder_blocks.0", DEPTH, **M.STAGE1)
    s2 = M.mix_transformer_stage(sd, "encoder_blocks.1", s1, **M.STAGE2)
    g1 = m.encoder_blocks[0](DEPTH.cuda())
    assert g1.shape == (3, 32, 15, 23)
    close(g1, s1); close(m.encoder_blocks[1](g1), s2)


def test_origunet_deployed_vs_golden_and_oracle(cuda_lib, G, manifest):
    m = load("OrigUNet_deployed", manifest, 21)
    frames = synthetic_frames(3, 2)
    fin = frames.clone().cuda()
    X = [fin, None, None]
    vel, (yi, yu, (hu, hv)) = m(X)
    assert X[2] == (None, None) and hv is None and yi.shape == (2, 1, 260, 346) and yu.shape == (2, 1, 68, 148)
    close(vel, G["UNetD_vel"]); close(yu, G["UNetD_upconv"]); close(yi[..., ::4, ::4], G["UNetD_interp_sub"])
    close(hu[0][0][:, ::16], G["UNetD_h"]); close(hu[0][1][:, ::16], G["UNetD_c"])
    s = yi.double().sum().item()
    assert abs(s - G["UNetD_interp_sum"][0]) <= 1e-5 * G["UNetD_interp_sum"][1]
    _, (_, yu2, _) = m([frames.flip(0).clone().cuda(), None, [hu, None]])
    close(yu2, G["UNetD_upconv2"])
    # live oracle on the same weights, full-resolution depth map
    sd = synth_state_dict(manifest["OrigUNet_deployed"], 21)
    _, (oi, ou, _) = M.orig_unet(sd, frames.clone(), None, **M.DEPLOYED_UNET_CFG)
    close(yi, oi); close(yu, ou)


def test_origunet_default_ctor_and_velpred11(cuda_lib, G, manifest):
    frames = synthetic_frames(3, 2)
    m = load("OrigUNet_default", manifest, 22)
    vel, (yi, yu, _) = m([frames[:1].clone().cuda(), None, None])
    close(vel, G["UNet0_vel"]); close(yu, G["UNet0_upconv"]); close(yi[..., ::4, ::4], G["UNet0_interp_sub"])
    m = load("OrigUNet_velpred11", manifest, 23)
    fin = frames.clone().cuda()
    vel, (_, yu, (hu, hv)) = m([fin, None, None])
    close(yu, G["UNetV_upconv"]); close(vel, G["UNetV_vel"])
    assert hu is None and hv is None
    ref = frames.clone(); ref[ref.abs() < 0.3] = 0.0
    assert torch.equal(fin.cpu(), ref)                               # form_input zeroed the caller's tensor in place


def test_full_model_deployed(cuda_lib, G, manifest):
    m = load("OrigUNet_w_VITFLY_ViTLSTM", manifest, 31)
    frames = synthetic_frames(3, 2)
    dv = torch.tensor([[4.0], [4.0]]).cuda()
    vel, (dep, yu, ((hu, hvp), hv)) = m([frames.clone().cuda(), dv, [None, None], None])
    close(vel, G["Full_vel"]); close(dep[..., ::4, ::4], G["Full_depth_sub"]); close(yu, G["Full_upconv"])
    close(hv[0], G["Full_hv"]); close(hv[1], G["Full_cv"])
    vel2, (dep2, _, _) = m([frames.flip(0).clone().cuda(), dv, [hu, None], hv])
    close(vel2, G["Full_vel2"]); close(dep2[..., ::4, ::4], G["Full_depth2_sub"])


def test_full_model_sequence_of_16_vs_oracle(cuda_lib, manifest):
    # a longer sequence: the recurrences (ConvLSTM over 16 steps, LSTM over 16 steps) accumulate
    # rounding differences, so this is the realistic fp32 tolerance check
    m = load("OrigUNet_w_VITFLY_ViTLSTM", manifest, 31)
    sd = synth_state_dict(manifest["OrigUNet_w_VITFLY_ViTLSTM"], 31)
    frames = synthetic_frames(9, 16)
    dv = torch.full((16, 1), 4.0)
    vel, (dep, _, _) = m([frames.clone().cuda(), dv.cuda(), [None, None], None])
    ovel, (odep, _, _) = M.orig_unet_w_vitlstm(sd, frames.clone(), dv, None, None, **M.DEPLOYED_UNET_CFG)
    close(dep, odep, rtol=1e-5, atol=2e-5); close(vel, ovel, rtol=1e-5, atol=2e-5)
