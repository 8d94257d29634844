"""Measure the bf16 path's error against the fp32 oracle (exploration)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evfly_b200
from oracle import model_oracle as M
from oracle.synth_ckpt import synth_state_dict, synthetic_frames
from tests.test_models_cpu import build
torch.set_grad_enabled(False)
man = json.load(open("tests/golden/state_dict_manifest.json"))
def stats(got, want):
    g, w = got.float().cpu().double().numpy(), want.double().numpy()
    e = np.abs(g - w); s = np.abs(w).mean()
    return dict(rel_l2=float(np.linalg.norm(g - w) / np.linalg.norm(w)), mean_over_scale=float(e.mean() / s), max_over_scale=float(e.max() / s), scale=float(s))
name, seed = "OrigUNet_w_VITFLY_ViTLSTM", 31
m = build(name); sd = synth_state_dict(man[name], seed); m.load_state_dict(sd)
m = evfly_b200.set_precision(m.cuda().eval(), "bf16")
for T in (2, 8):
    frames = synthetic_frames(9, T); dv = torch.full((T, 1), 4.0)
    vel, (dep, yu, ((hu, _), hv)) = m([frames.clone().cuda(), dv.cuda(), [None, None], None])
    ovel, (odep, oyu, ((ohu, _), ohv)) = M.orig_unet_w_vitlstm(sd, frames.clone(), dv, None, None, **M.DEPLOYED_UNET_CFG)
    print(T, json.dumps({"depth": stats(dep, odep), "upconv": stats(yu, oyu), "vel": stats(vel, ovel), "h_unet": stats(hu[0][0], ohu[0][0])}))
