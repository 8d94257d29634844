"""Drop-in for the event-frame normalisation inside evfly's learner/dataloading.py (:508-533) -- the step
between accumulation and the model on the offline path ("next" row N1). File IO stays with the reference.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .events import _device


def normalize_event_frames(ev, rescale_evs: float = -1.0, evs_min_cutoff=None) -> torch.Tensor:
    """One trajectory of event frames [T,H,W] -> normalised frames (float32 CUDA tensor), exactly as
    dataloading.py does per trajectory:
      rescale_evs > 0   : clamp(ev / rescale_evs, -1, 1)                                   (:511-512)
      rescale_evs == -1 : per frame s = quantile(|ev|, 0.97); clamp(ev / s, -1, 1)         (:515-524)
      evs_min_cutoff    : |ev| < cutoff -> 0                                               (:531-533)
    """
    lib = _lib.load()
    dev = _device()
    x = torch.as_tensor(ev).to(device=dev, dtype=torch.float32).contiguous()
    T = x.shape[0]
    elems = x[0].numel()
    cutoff = 0.0 if evs_min_cutoff is None else float(evs_min_cutoff)
    out = torch.empty_like(x)
    if rescale_evs > 0.0:
        ops.map4d(x.view(T, -1), out.view(T, -1), div=float(rescale_evs), lo=-1.0, hi=1.0)
        if cutoff > 0:
            _lib.check(lib.evfly_min_cutoff_f32(_lib.ptr(out), out.numel(), cutoff, _lib.stream_ptr()), "evfly_min_cutoff_f32")
    elif rescale_evs == -1.0:
        _lib.check(lib.evfly_quantile_scale_clip(_lib.ptr(x), T, elems, 0.97, -1.0, 1.0, cutoff, _lib.ptr(out), None,
                                                 _lib.stream_ptr()), "evfly_quantile_scale_clip")
    else:
        out.copy_(x)
        if cutoff > 0:
            _lib.check(lib.evfly_min_cutoff_f32(_lib.ptr(out), out.numel(), cutoff, _lib.stream_ptr()), "evfly_min_cutoff_f32")
    return out


def resize_trajectory(x, resize_input) -> torch.Tensor:
    """dataloading.py:401-416: F.interpolate(traj.unsqueeze(1), size=resize_input, mode='bilinear',
    align_corners=False).squeeze() for one trajectory of images / depths / event frames [T,H,W] -> [T,h,w]
    (float32 CUDA tensor). Same sampling positions and lerp order as torch (evfly_resize_bilinear_f32)."""
    dev = _device()
    x = torch.as_tensor(x).to(device=dev, dtype=torch.float32).contiguous()
    T, H, W = x.shape
    h, w = resize_input
    if (H, W) == (h, w):
        return x
    return ops.resize_bilinear(x.view(T, 1, H, W), (h, w), align_corners=False).view(T, h, w)


# ---------------------------------------------------------------------------------------------------------------
# Row N4: the dataset's event-frame wire format -> device buffers (learner/dataloading.py:158-171, 348, 401-416,
# 501-533; written by utils/to_events.py:441-456). A dataset directory holds one folder per trajectory with a
# data.csv (21 columns for the png datasets, column 1 = timestamp, column 2 = desired velocity: :362-365) and ONE
# `evs_frames*.npy` for all trajectories: a numpy OBJECT array with one float array [T_k, H, W] per trajectory, or --
# the single-trajectory case of to_events.py:446-448 -- a plain float array [1, T, H, W]. The h5 variant
# (utils/to_h5.py:16-45: groups {data, ims, depths, evs, ...}) needs h5py, which this image does not have; when it
# is importable the same reader takes the `evs` / `data` datasets of every group.
# ---------------------------------------------------------------------------------------------------------------
def read_event_frames(path):
    """evs_frames*.npy -> list of float32 numpy arrays [T_k, H, W], one per trajectory (host side, no copy when the
    file already stores float32)."""
    import numpy as np
    arr = np.load(path, allow_pickle=True)
    if arr.dtype == object:
        return [np.ascontiguousarray(a, dtype=np.float32) for a in arr]
    if arr.ndim == 4:                    # np.asarray(alltrajs_frames) of a single trajectory (or equal-length ones)
        return [np.ascontiguousarray(a, dtype=np.float32) for a in arr]
    if arr.ndim == 3:
        return [np.ascontiguousarray(arr, dtype=np.float32)]
    raise ValueError(f"{path}: expected an object array of [T,H,W] arrays or a float array [n,T,H,W], got {arr.dtype} {arr.shape}")


def read_trajectory_meta(traj_folder):
    """data.csv of one trajectory folder -> float64 [T, ncol] exactly as dataloading.py:196-223 parses it (header line
    skipped; for the 21-column png datasets, lines with another column count are dropped)."""
    import os
    import numpy as np
    rows = []
    with open(os.path.join(traj_folder, "data.csv")) as f:
        lines = f.readlines()
    for line in lines[1:]:
        cols = line.strip().split(",")
        if len(cols) != 21:
            continue
        rows.append([float(x) for x in cols])
    return np.array(rows, dtype=np.float64)


class EventFrameDataset:
    """The event frames of a dataset on the device, normalised like the reference dataloader returns them:
    per trajectory resize (dataloading.py:412-414) -> rescale / per-frame 97th-percentile scale + clamp (:508-524) ->
    min cutoff (:531-533), through the kernels of normalize_event_frames / resize_trajectory (bit-exact against the
    reference on its own golden outputs). Frames travel host -> device through ONE reusable pinned staging buffer.

        ds = EventFrameDataset(data_dir, events="evs_frames.npy", rescale_evs=-1.0, evs_min_cutoff=1e-3, resize_input=(260, 346))
        ds.frames[k]      float32 CUDA [T_k, h, w]      ds.desvel[k]  float32 CUDA [T_k]      ds.lengths[k]
        model([ds.frames[k].unsqueeze(1), ds.desvel[k].view(-1, 1), [None, None], None])       # learner.py:992,1071
    """

    def __init__(self, data_dir, events="evs_frames.npy", rescale_evs=-1.0, evs_min_cutoff=None, resize_input=None, traj_ids=None, device=None):
        import glob
        import os
        import numpy as np
        dev = _device(device)
        if not events.endswith(".npy"):
            events = events + ".npy"             # the reference's np.load adds nothing; its configs name the file without suffix... (:164)
        path = os.path.join(data_dir, events)
        if not os.path.exists(path) and os.path.exists(path[:-4]):
            path = path[:-4]
        host = read_event_frames(path)
        folders = sorted(p for p in glob.glob(os.path.join(data_dir, "*")) if os.path.isdir(p))
        ids = list(range(len(host))) if traj_ids is None else list(traj_ids)
        self.folders = [folders[i] for i in ids] if len(folders) >= len(host) else []
        self.frames, self.desvel, self.lengths = [], [], []
        cap = max(int(np.prod(host[i].shape)) for i in ids) if ids else 0
        stage = torch.empty((cap,), dtype=torch.float32).pin_memory() if cap else None
        for i in ids:
            fr = host[i]
            n = fr.size
            stage[:n].copy_(torch.from_numpy(fr).reshape(-1))
            d = stage[:n].to(dev, non_blocking=True).view(fr.shape)
            if resize_input is not None and tuple(d.shape[-2:]) != tuple(resize_input):
                d = resize_trajectory(d, resize_input)
            d = normalize_event_frames(d, rescale_evs=rescale_evs, evs_min_cutoff=evs_min_cutoff)
            torch.cuda.current_stream().synchronize()            # the staging buffer is reused by the next trajectory
            self.frames.append(d)
            self.lengths.append(int(fr.shape[0]))
            if self.folders:
                meta = read_trajectory_meta(self.folders[len(self.frames) - 1])
                self.desvel.append(torch.from_numpy(meta[:, 2].astype(np.float32)).to(dev))
