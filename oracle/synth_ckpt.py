"""Seeded synthetic checkpoints (there are no pretrained weights in the reference tree and no
network: README.md:131 points at an external tarball). TEST INFRASTRUCTURE.

synth_state_dict(shapes, seed) fills a {key: shape} table deterministically, key by key, with
He-scaled weights so activations stay O(1) through the 23 un-normalised UNet convs, and with
spectral-norm u/v vectors converged by power iteration (otherwise sigma ~ 1e-3 and outputs blow
up). The same function feeds the reference modules (golden generation) and
the drop-in modules (tests, bench), so both load bit-identical tensors.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(key: str, seed: int):
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    sd = {}
    for key in shapes:
        shape = tuple(shapes[key])
        r = _rng(key, seed)
        leaf = key.split(".")[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if leaf == "running_var":
            a = r.uniform(0.5, 1.5, shape)
        elif leaf == "running_mean":
            a = 0.1 * r.standard_normal(shape)
        elif leaf.startswith("weight_ih") or leaf.startswith("weight_hh") or leaf.startswith("bias_ih") or leaf.startswith("bias_hh"):
            hidden = shape[0] // 4
            a = r.uniform(-1, 1, shape) / np.sqrt(hidden)
        elif leaf in ("weight_u", "weight_v"):
            continue  # filled with weight_orig below
        elif leaf in ("weight", "weight_orig"):
            if len(shape) == 1:                      # LayerNorm / BatchNorm gain
                a = 1.0 + 0.1 * r.standard_normal(shape)
            else:
                is_convT = "upconv" in key and len(shape) == 4
                fan_in = shape[0] if is_convT else int(np.prod(shape[1:]))
                a = r.standard_normal(shape) * np.sqrt(2.0 / fan_in)
                if "unet_out" in key:            # keep the depth map inside clip(2*d, 0, 1)'s range
                    a *= 0.15
        elif leaf == "bias":
            a = 0.05 * r.standard_normal(shape)
        else:
            raise KeyError(f"synth_state_dict: no rule for {key} {shape}")
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    # spectral-norm vectors
    for key in shapes:
        if key.endswith(".weight_orig"):
            base = key[: -len("weight_orig")]
            w = sd[key].double().reshape(sd[key].shape[0], -1).numpy()
            r = _rng(base + "uv", seed)
            u = r.standard_normal(w.shape[0])
            for _ in range(50):
                v = w.T @ u
                v /= np.linalg.norm(v) + 1e-12
                u = w @ v
                u /= np.linalg.norm(u) + 1e-12
            sd[base + "weight_u"] = torch.from_numpy(u.astype(np.float32))
            sd[base + "weight_v"] = torch.from_numpy(v.astype(np.float32))
    return {k: sd[k] for k in shapes}


def shapes_of(module) -> dict:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def synthetic_frames(seed: int, n: int, H: int = 260, W: int = 346, active: float = 0.35) -> torch.Tensor:
    """Normalised event frames in [-1,1] like run.py:253 produces: multiples of 1/k with a
    fraction `active` of non-zero pixels (> 3 %)."""
    r = np.random.default_rng([seed, 77])
    cnt = r.integers(-6, 7, (n, 1, H, W)) * (r.random((n, 1, H, W)) < active)
    return torch.from_numpy(np.clip(cnt / 4.0, -1, 1).astype(np.float32))


def synthetic_depth(seed: int, n: int, H: int = 60, W: int = 90) -> torch.Tensor:
    r = np.random.default_rng([seed, 78])
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.empty((n, 1, H, W), dtype=np.float32)
    for i in range(n):
        cx, cy, s = r.uniform(0, W), r.uniform(0, H), r.uniform(8, 30)
        out[i, 0] = np.clip(0.5 + 0.5 * np.sin(xx / s + i) * np.cos(yy / s) - 0.6 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s)), 0, 1)
    return torch.from_numpy(out)
