"""GPU: edge cases of the L1/L2 entry points -- empty inputs, single events, ragged sizes, windows without
events, extra columns / tensor inputs of form_eventframe, and argument validation through the ABI."""
import numpy as np
import pytest
import torch

from evfly_b200 import _lib
from evfly_b200.events import L1, make_records, to_device
from evfly_b200.ev_utils import form_eventframe, form_voxelgrid
from evfly_b200.synthetic import records_to_rows, synthetic_window
from oracle import ev_oracle as O

pytestmark = pytest.mark.gpu


def test_empty_and_single_event_inputs(cuda_lib):
    H, W = 7, 9
    empty = torch.empty((0, 16), dtype=torch.uint8, device="cuda")
    assert not L1.accumulate_counts(empty, H, W).any()
    c, v = L1.voxelize_window(empty, H, W, 5, 0, 100, algo=1)
    assert not c.any() and not v.any()
    c, v = L1.voxelize_window(empty, H, W, 5, 0, 100, algo=0)
    assert not c.any() and not v.any()
    edges = torch.tensor([0, 10, 20, 30], dtype=torch.int64, device="cuda")
    c, v = L1.accumulate_windows(empty, edges, H, W, 3)
    assert c.shape == (3, 2, H, W) and not c.any() and not v.any()           # outputs are zero-filled by the call
    one = make_records([8], [6], [15], [1])
    c, v = L1.accumulate_windows(to_device(one), edges, H, W, 3)
    assert c.sum().item() == 1 and c[1, 1, 6, 8].item() == 1
    assert abs(v[1].sum().item() - 1.0) < 1e-6 and v[1, :, 6, 8].sum().item() == pytest.approx(1.0)
    # event exactly on the window boundaries: t0 inclusive, t1 exclusive (ev_utils.py:128)
    two = make_records([1, 2], [1, 1], [10, 20], [1, 1])
    c, _ = L1.voxelize_window(to_device(two), H, W, 2, 10, 20)
    assert c[1, 1, 1].item() == 1 and c[1, 1, 2].item() == 0


def test_ragged_sizes_and_unroll_tails(cuda_lib):
    for n in (1, 3, 255, 256, 257, 1023, 1025, 4097):
        for H, W in ((1, 1), (3, 5), (260, 346)):
            rec = synthetic_window(n, n, H, W)
            assert np.array_equal(L1.accumulate_counts(to_device(rec), H, W).cpu().numpy(), O.event_counts(rec, H, W))


def test_form_eventframe_input_variants(cuda_lib):
    H, W = 24, 32
    rec = synthetic_window(1, 500, H, W)
    rows = records_to_rows(rec)
    want = O.form_eventframe(rows, H, W, all_events=True)
    assert np.array_equal(form_eventframe(torch.from_numpy(rows), H, W, all_events=True), want)          # torch CPU tensor
    assert np.array_equal(form_eventframe(torch.from_numpy(rows).cuda(), H, W, all_events=True), want)   # already on the device
    assert np.array_equal(form_eventframe(rows.astype(np.float32), H, W, all_events=True), want)          # float32 rows
    wide = np.concatenate([rows[:, :3], np.full((rows.shape[0], 2), 9.0), rows[:, 3:]], axis=1)           # extra columns: p is the LAST one
    assert np.array_equal(form_eventframe(wide, H, W, all_events=True), want)
    with pytest.raises(ValueError):
        form_eventframe(rows[:, :3], H, W, all_events=True)


def test_voxel_bins_follow_time(cuda_lib):
    # a monotone ramp of events: bin b collects the events around b/(B-1) of the window
    H, W, B, n = 2, 2, 5, 4001
    t = np.linspace(0, 4000, n).astype(np.int64)
    rec = make_records(np.zeros(n, int), np.zeros(n, int), t, np.ones(n, int))
    c, v = form_voxelgrid(to_device(rec), H, W, 0, 4001, B)
    v = v[:, 0, 0].cpu().numpy()
    assert c[1, 0, 0].item() == n and abs(v.sum() - n) < 1e-2
    assert v[0] < v[1] and abs(v[1] - v[2]) < 2 and abs(v[2] - v[3]) < 2 and v[4] < v[3]     # triangle kernels: ends get half the mass


def test_abi_argument_validation_on_device(cuda_lib):
    lib = cuda_lib
    x = torch.zeros((4, 16), dtype=torch.uint8, device="cuda")
    c = torch.zeros((2, 4, 4), dtype=torch.int32, device="cuda")
    assert lib.evfly_accumulate_counts(x.data_ptr(), -1, 4, 4, c.data_ptr(), None) == -1
    assert lib.evfly_voxelize_window(x.data_ptr(), 4, 4, 4, 0, 0, 10, c.data_ptr(), None, None, 0, None) == -1      # B = 0
    assert lib.evfly_voxelize_window(x.data_ptr(), 4, 4, 4, 5, 0, 10, c.data_ptr(), None, None, 1, None) == -1      # staged needs a workspace
    assert lib.evfly_decode_crop(None, None, 1, 4, 4, 2, 2, 0.2, c.data_ptr(), None) == -1
    with pytest.raises(_lib.EvflyError):
        _lib.ptr(torch.zeros(3))                     # host tensors never reach the ABI
    with pytest.raises(_lib.EvflyError):
        L1.accumulate_counts(torch.zeros((4, 16), dtype=torch.uint8), 4, 4)
