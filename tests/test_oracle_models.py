"""CPU: the functional model oracle (oracle/model_oracle.py) against the golden vectors that
the reference's own nn.Modules produced (tests/golden/make_golden_models.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as M
from oracle.synth_ckpt import synth_state_dict, synthetic_depth, synthetic_frames



@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield

RTOL, ATOL = 1e-5, 2e-6      # fp32 re-association noise between two CPU op orders


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "models_golden.npz"))


@pytest.fixture(scope="module")
def manifest(golden_dir):
    return json.load(open(os.path.join(golden_dir, "state_dict_manifest.json")))


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


ENC = dict(num_layers=2, kernel_sizes=[5, 3], kernel_strides=[2, 2], out_channels=[8, 32], activations=["relu", "relu"],
           pool_type="max", invert_pool_inputs=True, pool_kernels=[2, 2], pool_strides=[2, 2], conv_function="conv2d")
FC = dict(num_layers=4, layer_sizes=[1024, 128, 16, 1], activations=["leaky_relu"] * 3 + ["tanh"], dropout_p=0.1)
DEPTH = synthetic_depth(1, 3)
DESVEL = torch.tensor([[4.0], [5.5], [3.0]])
QUAT = torch.tensor([[1.0, 0, 0, 0], [0.9, 0.1, -0.2, 0.3], [0.7, 0.0, 0.7, 0.1]])


def test_param_counts_match_reference_docstrings(manifest):
    # vitfly_models.py:36,75,114,155,191 (trainable parameters; buffers excluded)
    def count(name):
        skip = ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")
        return sum(int(np.prod(s)) for k, s in manifest[name].items() if not k.endswith(skip))
    assert count("ConvNet") == 235269 and count("LSTMNet") == 2949937
    assert count("UNetConvLSTMNet") == 2955822 and count("ViT") == 3101199 and count("LSTMNetVIT") == 3563663
    assert count("OrigUNet_default") == 7759809


@pytest.mark.parametrize("name,seed,fn,stateful", [
    ("LSTMNetVIT", 11, M.lstmnet_vit, True), ("LSTMNet", 14, M.lstmnet, True),
    ("UNetConvLSTMNet", 15, M.unet_convlstm_net, True), ("ViT", 12, M.vit, False), ("ConvNet", 13, M.convnet, False)])
def test_vitfly_models(G, manifest, name, seed, fn, stateful):
    sd = synth_state_dict(manifest[name], seed)
    if stateful:
        vel, hc = fn(sd, DEPTH.clone(), DESVEL, QUAT)
        close(vel, G[f"{name}_vel"]); close(hc[0], G[f"{name}_h"]); close(hc[1], G[f"{name}_c"])
        vel2, hc2 = fn(sd, DEPTH.flip(0).clone(), DESVEL, None, hc)
        close(vel2, G[f"{name}_vel2"]); close(hc2[0], G[f"{name}_h2"])
    else:
        close(fn(sd, DEPTH.clone(), DESVEL, QUAT), G[f"{name}_vel"])
        close(fn(sd, DEPTH.flip(0).clone(), DESVEL, None), G[f"{name}_vel2"])


def test_lstmnetvit_resizes_large_depth(G, manifest):
    sd = synth_state_dict(manifest["LSTMNetVIT"], 11)
    vel, _ = M.lstmnet_vit(sd, torch.from_numpy(G["LSTMNetVIT_big_in"]), DESVEL[:2])
    close(vel, G["LSTMNetVIT_big_vel"])


def test_sequence_equals_stepwise(manifest):
    # the batch dimension is time for the LSTM
    sd = synth_state_dict(manifest["LSTMNetVIT"], 11)
    vel, hc = M.lstmnet_vit(sd, DEPTH.clone(), DESVEL, QUAT)
    st, outs = None, []
    for t in range(3):
        v, st = M.lstmnet_vit(sd, DEPTH[t:t + 1].clone(), DESVEL[t:t + 1], QUAT[t:t + 1], st)
        outs.append(v)
    close(torch.cat(outs), vel.numpy(), atol=1e-5); close(st[0], hc[0].numpy(), atol=1e-5)


def test_origunet_deployed(G, manifest):
    sd = synth_state_dict(manifest["OrigUNet_deployed"], 21)
    frames = synthetic_frames(3, 2)
    vel, (yi, yu, (hu, _)) = M.orig_unet(sd, frames.clone(), None, **M.DEPLOYED_UNET_CFG)
    close(vel, G["UNetD_vel"]); close(yu, G["UNetD_upconv"], atol=1e-5); close(yi[..., ::4, ::4], G["UNetD_interp_sub"], atol=1e-5)
    close(hu[0][0][:, ::16], G["UNetD_h"], atol=1e-5); close(hu[0][1][:, ::16], G["UNetD_c"], atol=1e-5)
    assert abs(yi.double().sum().item() - G["UNetD_interp_sum"][0]) <= 1e-5 * G["UNetD_interp_sum"][1]
    _, (_, yu2, _) = M.orig_unet(sd, frames.flip(0).clone(), (hu, None), **M.DEPLOYED_UNET_CFG)
    close(yu2, G["UNetD_upconv2"], atol=1e-5)


def test_origunet_default_ctor_and_velpred11(G, manifest):
    frames = synthetic_frames(3, 2)
    sd = synth_state_dict(manifest["OrigUNet_default"], 22)
    vel, (yi, yu, _) = M.orig_unet(sd, frames[:1].clone(), None)
    close(vel, G["UNet0_vel"]); close(yu, G["UNet0_upconv"], atol=1e-5); close(yi[..., ::4, ::4], G["UNet0_interp_sub"], atol=1e-5)
    sd = synth_state_dict(manifest["OrigUNet_velpred11"], 23)
    vel, (_, yu, _) = M.orig_unet(sd, frames.clone(), None, form_bev=1, cutoff=0.3, skip_type="none", velpred=11, enc_params=ENC, fc_params=FC)
    close(yu, G["UNetV_upconv"], atol=1e-5); close(vel, G["UNetV_vel"], atol=1e-5)


def test_full_model_deployed(G, manifest):
    sd = synth_state_dict(manifest["OrigUNet_w_VITFLY_ViTLSTM"], 31)
    frames = synthetic_frames(3, 2)
    dv = torch.tensor([[4.0], [4.0]])
    vel, (dep, yu, ((hu, _), hv)) = M.orig_unet_w_vitlstm(sd, frames.clone(), dv, None, None, **M.DEPLOYED_UNET_CFG)
    close(vel, G["Full_vel"], atol=1e-5); close(dep[..., ::4, ::4], G["Full_depth_sub"], atol=1e-5); close(yu, G["Full_upconv"], atol=1e-5)
    close(hv[0], G["Full_hv"], atol=1e-5); close(hv[1], G["Full_cv"], atol=1e-5)
    vel2, (dep2, _, _) = M.orig_unet_w_vitlstm(sd, frames.flip(0).clone(), dv, hu, hv, **M.DEPLOYED_UNET_CFG)
    close(vel2, G["Full_vel2"], atol=1e-5); close(dep2[..., ::4, ::4], G["Full_depth2_sub"], atol=1e-5)
    assert 0.05 < float(np.clip(G["Full_depth_sub"] * 2, 0, 1).std())      # the ViT input is not saturated
