"""GPU (row N4): the event-frame dataset reader against the REFERENCE dataloader's own outputs (tests/golden/
dataloading_golden.npz, produced by running learner/dataloading.py on the synthetic folder dataset that
make_golden_dataloading.py builds): the same folder layout is rebuilt here (data.csv + evs_frames.npy object array),
read by EventFrameDataset, and the normalised frames must equal the reference's bit for bit in every mode."""
import os

import numpy as np
import pytest
import torch

from evfly_b200.dataloading import EventFrameDataset, read_event_frames, read_trajectory_meta

pytestmark = pytest.mark.gpu


def _build(root, G, tag):
    evs = np.empty(2, dtype=object)
    for k in range(2):
        d = os.path.join(root, f"{k}")
        os.makedirs(d)
        T = G[f"{tag}_in{k}"].shape[0]
        rows = np.zeros((T, 21))
        rows[:, 1] = [float(f"{0.1 * (t + 1):.3f}") for t in range(T)]
        rows[:, 2] = 4.0 + k
        np.savetxt(os.path.join(d, "data.csv"), rows, delimiter=",", header=",".join(f"c{i}" for i in range(21)), comments="")
        evs[k] = G[f"{tag}_in{k}"]
    np.save(os.path.join(root, "evs_frames.npy"), evs, allow_pickle=True)


@pytest.mark.parametrize("tag,resize", [("native", None), ("resized", (60, 90))])
def test_reader_equals_reference_dataloader(cuda_lib, golden_dir, tmp_path, tag, resize):
    G = np.load(os.path.join(golden_dir, "dataloading_golden.npz"))
    root = str(tmp_path / "ds")
    os.makedirs(root)
    _build(root, G, tag)
    for mode, kw in {"q97": dict(rescale_evs=-1.0), "q97_cut": dict(rescale_evs=-1.0, evs_min_cutoff=0.05),
                     "div2_cut": dict(rescale_evs=2.0, evs_min_cutoff=0.05), "raw": dict(rescale_evs=0.0)}.items():
        ds = EventFrameDataset(root, events="evs_frames", resize_input=resize, **kw)
        assert ds.lengths == [int(x) for x in G[f"{tag}_lengths"]]
        for k in range(2):
            got, want = ds.frames[k].cpu().numpy(), G[f"{tag}_{mode}_out{k}"]
            assert got.shape == want.shape
            assert np.array_equal(got, want, equal_nan=True), (tag, mode, k, np.nanmax(np.abs(got - want)))
        assert np.array_equal(torch.cat(ds.desvel).cpu().numpy(), G[f"{tag}_desvel"].reshape(-1))
