"""Generate tests/golden/events_golden.npz by RUNNING THE REFERENCE (utils/ev_utils.py from
/root/reference, imported read-only; matplotlib is stubbed because only the plotting helpers
need it). Run in the authoring container only; the GPU box uses the committed .npz.

    python tests/golden/make_golden_events.py
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
sys.path.insert(0, os.path.join(REF, "utils"))
import ev_utils as ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "events_golden.npz")


def rows_uniform(seed, n, H, W, pol_values=(0.0, 1.0), t_hi=33e6):
    rng = np.random.default_rng(seed)
    rows = np.empty((n, 4))
    rows[:, 0] = np.sort(rng.uniform(0, t_hi, n))
    rows[:, 1] = rng.integers(0, W, n)
    rows[:, 2] = rng.integers(0, H, n)
    rows[:, 3] = rng.choice(pol_values, n)
    return rows


g = {}

# case A: all_events=True, integer coordinates, p in {0,1} (data_gather/depth_and_events_script.py:179)
H, W = 24, 32
rows = rows_uniform(1, 3000, H, W)
g["A_rows"], g["A_frame"] = rows, ref.form_eventframe(rows.copy(), H, W, all_events=True)

# case B: edge coordinates -- fractional, negative, exactly W / H (right edge is included by
# np.histogram2d), beyond, NaN; polarities 0, 1, -1, 2, 0.5
rows = rows_uniform(2, 400, H, W)
rows[:40, 1] = np.array([0, 0.5, 31, 31.999, 32, 32.0001, -0.0, -1e-9, -1, 33] * 4)
rows[40:80, 2] = np.array([0, 0.25, 23, 23.5, 24, 24.5, -0.5, 1e-12, 25, np.nan] * 4)
rows[80:120, 3] = np.array([0, 1, -1, 2, 0.5, -0.5, 0, 1, 1, 0] * 4)
g["B_rows"] = rows
g["B_frame_all"] = ref.form_eventframe(rows.copy(), H, W, all_events=True)
g["B_frame_timed"], _ = ref.form_eventframe(rows.copy(), H, W, times0=0.0, times1=[0.02], all_events=False)

# case C: timed mode with p in {-1,+1}, thresholds != 0.2
rows = rows_uniform(3, 5000, H, W, pol_values=(-1.0, 1.0))
t1 = [0.0215]
g["C_rows"] = rows
g["C_frame"], g["C_times1"] = ref.form_eventframe(rows.copy(), H, W, times0=0.004, times1=t1, pos_thresh=0.3, neg_thresh=0.15)
g["C_args"] = np.array([0.004, t1[0], 0.3, 0.15])

# case D: N mode (first N events after times0); returns times1 = (t_last + 1) / 1e9
g["D_frame"], g["D_times1"] = ref.form_eventframe(rows.copy(), H, W, times0=0.01, N=777)
g["D_args"] = np.array([0.01, 777])
# N larger than what is left
g["D2_frame"], g["D2_times1"] = ref.form_eventframe(rows.copy(), H, W, times0=0.03, N=100000)

# case E: empty inputs
g["E_all"] = ref.form_eventframe(np.zeros((0, 4)), H, W, all_events=True)
e_frame, e_t = ref.form_eventframe(np.zeros((0, 4)), H, W, times0=0.5, times1=[1.0])
g["E_timed"], g["E_timed_t"] = e_frame, np.array(e_t)

# case F: a hot pixel with hundreds of events (counts >> 1) and a frame with one polarity only
rows = rows_uniform(4, 2000, H, W)
rows[:900, 1:3] = (5, 7)
rows[900:, 3] = 1.0
g["F_rows"], g["F_frame"] = rows, ref.form_eventframe(rows.copy(), H, W, all_events=True)

# case G: BASELINE config 1 shape (100k events, 260x346) -- only a digest of the frame is stored
H1, W1 = 260, 346
rng = np.random.default_rng(0)
n = 100_000
rows = np.empty((n, 4))
rows[:, 0] = np.sort(rng.integers(0, 33_333_333, n))
rows[:, 1] = rng.integers(0, W1, n)
rows[:, 2] = rng.integers(0, H1, n)
rows[:, 3] = rng.integers(0, 2, n)
fr = ref.form_eventframe(rows.copy(), H1, W1, all_events=True)
g["G_seed_n_H_W"] = np.array([0, n, H1, W1])
g["G_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(fr).tobytes()).digest(), dtype=np.uint8)
g["G_sum_abs"] = np.array([np.abs(fr).sum(), fr.sum(), fr.max(), fr.min()])

np.savez_compressed(OUT, **g)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(g), "arrays")
