"""Generate tests/golden/models_golden.npz + state_dict_manifest.json by RUNNING THE REFERENCE's
own nn.Modules (imported read-only from /root/reference/learner) on seeded synthetic
checkpoints (oracle/synth_ckpt.py) and inputs. Authoring container only.

    python tests/golden/make_golden_models.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/learner")
import learner_models as RL  # noqa: E402
import vitfly_models as RV  # noqa: E402
from oracle.synth_ckpt import shapes_of, synth_state_dict, synthetic_depth, synthetic_frames  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)
quiet = lambda *a, **k: None

ENC = dict(num_layers=2, kernel_sizes=[5, 3], kernel_strides=[2, 2], out_channels=[8, 32], activations=["relu", "relu"],
           pool_type="max", invert_pool_inputs=True, pool_kernels=[2, 2], pool_strides=[2, 2], conv_function="conv2d")
FC = dict(num_layers=4, layer_sizes=[1024, 128, 16, 1], activations=["leaky_relu"] * 3 + ["tanh"], dropout_p=0.1)

g, manifest = {}, {}


def load(model, seed, name):
    shapes = shapes_of(model)
    manifest[name] = {k: list(v) for k, v in shapes.items()}
    model.load_state_dict(synth_state_dict(shapes, seed), strict=True)
    return model.eval().float()


def sub(a, step=4):
    return a[..., ::step, ::step].contiguous().numpy()


# ---- vitfly_models ---------------------------------------------------------------------------
depth = synthetic_depth(1, 3)
desvel = torch.tensor([[4.0], [5.5], [3.0]])
quat = torch.tensor([[1.0, 0, 0, 0], [0.9, 0.1, -0.2, 0.3], [0.7, 0.0, 0.7, 0.1]])
for name, cls, seed in (("LSTMNetVIT", RV.LSTMNetVIT, 11), ("ViT", RV.ViT, 12), ("ConvNet", RV.ConvNet, 13),
                        ("LSTMNet", RV.LSTMNet, 14), ("UNetConvLSTMNet", RV.UNetConvLSTMNet, 15)):
    m = load(cls(), seed, name)
    out, h = m([depth.clone(), desvel, quat.clone()])
    g[f"{name}_vel"] = out.numpy()
    if h is not None:
        g[f"{name}_h"], g[f"{name}_c"] = h[0].numpy(), h[1].numpy()
        out2, h2 = m([depth.flip(0).clone(), desvel, None, h])        # state carried, quat defaulted
        g[f"{name}_vel2"], g[f"{name}_h2"] = out2.numpy(), h2[0].numpy()
    else:
        out2, _ = m([depth.flip(0).clone(), desvel, None])
        g[f"{name}_vel2"] = out2.numpy()
# LSTMNetVIT fed a 260x346 depth map (refine_inputs resizes, vitfly_models.py:28-29)
m = load(RV.LSTMNetVIT(), 11, "LSTMNetVIT")
big = torch.nn.functional.interpolate(synthetic_depth(2, 2), size=(260, 346), mode="bicubic").clamp(0, 1)
g["LSTMNetVIT_big_in"] = big.numpy()
g["LSTMNetVIT_big_vel"] = m([big.clone(), desvel[:2], None])[0].numpy()

# ---- OrigUNet --------------------------------------------------------------------------------
frames = synthetic_frames(3, 2)
# deployed config (configs/eval_config_real.txt:39-47)
m = load(RL.OrigUNet(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346], logger=quiet,
                     velpred=0, enc_params=ENC, fc_params=FC, form_BEV=2, evs_min_cutoff=1e-3, skip_type="interp"), 21, "OrigUNet_deployed")
vel, (yi, yu, (hu, _)) = m([frames.clone(), None, None])
g["UNetD_vel"], g["UNetD_interp_sub"], g["UNetD_upconv"] = vel.numpy(), sub(yi), yu.numpy()
g["UNetD_h"], g["UNetD_c"] = hu[0][0].numpy()[:, ::16], hu[0][1].numpy()[:, ::16]
g["UNetD_interp_sum"] = np.array([yi.double().sum().item(), yi.double().abs().sum().item()])
vel, (yi2, yu2, _) = m([frames.flip(0).clone(), None, [hu, None]])                 # state carried
g["UNetD_upconv2"] = yu2.numpy()
# constructor defaults: 2-channel input, crop skips, no recurrence
m = load(RL.OrigUNet(num_recurrent=[0, 0], logger=quiet), 22, "OrigUNet_default")
vel, (yi, yu, _) = m([frames[:1].clone(), None, None])
g["UNet0_vel"], g["UNet0_interp_sub"], g["UNet0_upconv"] = vel.numpy(), sub(yi), yu.numpy()
# joint model (configs/eval_config_sim_joint.txt: velpred = 11), bev=1, no skips
m = load(RL.OrigUNet(num_in_channels=2, num_recurrent=[0, 0], input_shape=[1, 1, 260, 346], logger=quiet, velpred=11,
                     enc_params=ENC, fc_params=FC, form_BEV=1, evs_min_cutoff=0.3, skip_type="none"), 23, "OrigUNet_velpred11")
vel, (yi, yu, _) = m([frames.clone(), None, None])
g["UNetV_vel"], g["UNetV_upconv"] = vel.numpy(), yu.numpy()

# ---- OrigUNet_w_VITFLY_ViTLSTM, deployed config -------------------------------------------------
m = load(RL.OrigUNet_w_VITFLY_ViTLSTM(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346],
                                      logger=quiet, velpred=0, enc_params=ENC, fc_params=FC, form_BEV=2, evs_min_cutoff=1e-3,
                                      skip_type="interp", is_deployment=False), 31, "OrigUNet_w_VITFLY_ViTLSTM")
dv = torch.tensor([[4.0], [4.0]])
vel, (dep, yu, ((hu, _), hv)) = m([frames.clone(), dv, [None, None], None])
g["Full_vel"], g["Full_depth_sub"], g["Full_upconv"] = vel.numpy(), sub(dep), yu.numpy()
g["Full_hv"], g["Full_cv"] = hv[0].numpy(), hv[1].numpy()
vel2, (dep2, _, _) = m([frames.flip(0).clone(), dv, [hu, None], hv])
g["Full_vel2"], g["Full_depth2_sub"] = vel2.numpy(), sub(dep2)

np.savez_compressed(os.path.join(OUT, "models_golden.npz"), **g)
json.dump(manifest, open(os.path.join(OUT, "state_dict_manifest.json"), "w"), indent=0)
print({k: (v.shape, float(np.abs(v).mean())) for k, v in g.items()})
print("bytes:", os.path.getsize(os.path.join(OUT, "models_golden.npz")))
