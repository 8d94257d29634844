"""CPU (row N4): host-side parsing of the dataset wire format -- the object-array and the single-trajectory float-array
forms of evs_frames.npy (utils/to_events.py:441-456) and data.csv (learner/dataloading.py:196-223)."""
import os

import numpy as np

from evfly_b200.dataloading import read_event_frames, read_trajectory_meta


def test_object_array_and_single_trajectory_forms(tmp_path):
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((3, 6, 8)).astype(np.float32), rng.standard_normal((5, 6, 8)).astype(np.float64)
    obj = np.empty(2, dtype=object)
    obj[0], obj[1] = a, b
    np.save(tmp_path / "evs_frames.npy", obj, allow_pickle=True)
    got = read_event_frames(tmp_path / "evs_frames.npy")
    assert len(got) == 2 and got[0].dtype == np.float32 and np.array_equal(got[0], a) and np.array_equal(got[1], b.astype(np.float32))
    np.save(tmp_path / "single.npy", np.asarray([a]))                 # to_events.py:446-448
    got = read_event_frames(tmp_path / "single.npy")
    assert len(got) == 1 and np.array_equal(got[0], a)


def test_data_csv_rows_with_the_wrong_column_count_are_dropped(tmp_path):
    d = tmp_path / "0"
    os.makedirs(d)
    rows = np.arange(63, dtype=np.float64).reshape(3, 21)
    lines = [",".join(f"c{i}" for i in range(21))] + [",".join(repr(float(x)) for x in r) for r in rows]
    lines.insert(2, "1.0,2.0,3.0")                                    # a truncated line (dataloading.py:216-219 skips it)
    (d / "data.csv").write_text("\n".join(lines) + "\n")
    meta = read_trajectory_meta(str(d))
    assert meta.shape == (3, 21) and np.array_equal(meta, rows)
