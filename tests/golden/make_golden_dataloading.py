"""Golden vectors for the dataset side (rows N1 / N4): the REFERENCE dataloader run here on a tiny synthetic dataset
in its own folder format (trajectory folders with <ts>_im.png / <ts>_depth.png + data.csv, evs_frames.npy object
array; learner/dataloading.py:30-560). h5py is not installed and not needed for the folder format, so an empty stub
module satisfies the import. Stored: the raw event frames and what the reference returns for them under the three
normalisation modes (:508-533), with and without the resize_input step (:401-416).

    python tests/golden/make_golden_dataloading.py     (needs /root/reference; writes dataloading_golden.npz)
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np

sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/learner")
import dataloading as RD  # noqa: E402  (the reference module)


def build_dataset(root, rng, n_traj, T, H, W):
    evs = np.empty(n_traj, dtype=object)
    for k in range(n_traj):
        d = os.path.join(root, f"{k}")
        os.makedirs(d)
        rows = []
        for t in range(T):
            ts = f"{0.1 * (t + 1):.3f}"
            cv2.imwrite(os.path.join(d, f"{ts}_im.png"), rng.integers(0, 256, (H, W), dtype=np.uint8))
            cv2.imwrite(os.path.join(d, f"{ts}_depth.png"), rng.integers(0, 256, (H, W), dtype=np.uint8))
            row = np.zeros(21)
            row[1], row[2] = float(ts), 4.0 + k
            rows.append(row)
        np.savetxt(os.path.join(d, "data.csv"), np.array(rows), delimiter=",", header=",".join(f"c{i}" for i in range(21)), comments="")
        # event frames: multiples of 0.2 (count differences), sparse, a few hot pixels, one all-zero frame
        n = rng.poisson(0.4, (T, H, W)) - rng.poisson(0.4, (T, H, W))
        n[rng.random((T, H, W)) < 0.002] += 40
        if k == 0:
            n[0] = 0
        evs[k] = (0.2 * n).astype(np.float32)
    np.save(os.path.join(root, "evs_frames.npy"), evs, allow_pickle=True)
    return evs


def main():
    rng = np.random.default_rng(20260118)
    g = {}
    for tag, (H, W, resize) in {"native": (60, 90, None), "resized": (120, 180, (60, 90))}.items():
        with tempfile.TemporaryDirectory() as tmp:
            root = os.path.join(tmp, "ds")
            os.makedirs(root)
            evs = build_dataset(root, rng, n_traj=2, T=3, H=H, W=W)
            for k in range(2):
                g[f"{tag}_in{k}"] = evs[k]
            for mode, kw in {"q97": dict(rescale_evs=-1.0), "q97_cut": dict(rescale_evs=-1.0, evs_min_cutoff=0.05),
                             "div2_cut": dict(rescale_evs=2.0, evs_min_cutoff=0.05), "raw": dict(rescale_evs=0.0)}.items():
                train, _, is_png = RD.dataloader(root, val_split=0.0, seed=-2, do_transform=False, events="evs_frames", use_h5=False,
                                                 resize_input=resize, logger=lambda *a: None, **kw)
                tevs = train[4]
                for k in range(2):
                    g[f"{tag}_{mode}_out{k}"] = tevs[k].numpy()
                if mode == "raw":
                    g[f"{tag}_desvel"] = train[3].numpy()
                    g[f"{tag}_lengths"] = np.asarray(train[2])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataloading_golden.npz"), **g)
    print({k: (v.shape, str(v.dtype)) for k, v in g.items()})


if __name__ == "__main__":
    main()
