"""Warm per-entry-point timing of one bench step (CUDA events around every C-ABI call, aggregated by name).
Unlike an ncu launch list the caches are warm and the kernels overlap as in production; event overhead is
a few microseconds per call. Usage: python scripts/profile_ops.py [trajectories|pipeline]"""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from evfly_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "trajectories"
torch.cuda.set_device(0)
wl = bench.WORKLOADS[name](0, torch.device("cuda", 0))
wl.OVERLAP = False      # serial: every entry point timed on its own
from evfly_b200 import tc
tc.USE_STAGE_ABI = False   # per-operator calls from Python (the stage-level call enqueues the same kernels from C++)
for i in range(3):
    wl.step(i)
torch.cuda.synchronize()
lib = _lib.load()
records = collections.defaultdict(list)
shapes = collections.defaultdict(list)
halo_shapes = collections.defaultdict(list)
order = []
orig = {}
for fn in _lib.SIGNATURES:
    if fn in ("evfly_last_error", "evfly_abi_version", "evfly_launch_count") or fn.endswith("_bytes"):
        continue
    f = getattr(lib, fn)
    orig[fn] = f
    def make(fn, f):
        def wrapped(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rc = f(*a); e1.record()
            records[fn].append((e0, e1))
            order.append((fn, e0, e1))
            if fn == "evfly_tc_conv_bf16":       # per-shape breakdown of the GEMM family
                st = a[0]._obj if hasattr(a[0], "_obj") else a[0].contents
                shapes[(st.M_rows, st.Cin, st.n_rows, st.taps, st.convt, st.Hp, st.Wp)].append((e0, e1))
            if fn in ("evfly_tc_conv3x3_halo_bf16", "evfly_tc_conv3x3_halo_compact_bf16", "evfly_tc_conv3x3_halo_pool_bf16", "evfly_tc_conv3x3_halo_pool_rows_bf16", "evfly_tc_conv3x3_halo_out1_bf16"):
                off = {"evfly_tc_conv3x3_halo_bf16": 4, "evfly_tc_conv3x3_halo_compact_bf16": 4, "evfly_tc_conv3x3_halo_pool_bf16": 5, "evfly_tc_conv3x3_halo_pool_rows_bf16": 5, "evfly_tc_conv3x3_halo_out1_bf16": 6}[fn]
                N, Hp, Wp, vh, vw, Cin, Cout = a[off:off + 7]      # (N, Hp, Wp, vh, vw, Cin, Cout) follow the pointers
                halo_shapes[(fn.replace("evfly_tc_conv3x3_", ""), N, vh, vw, Cin, Cout)].append((e0, e1))
            return rc
        return wrapped
    setattr(lib, fn, make(fn, f))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); wl.step(3); b.record()
torch.cuda.synchronize()
total = a.elapsed_time(b)
rows = sorted(((sum(x.elapsed_time(y) for x, y in v), len(v), k) for k, v in records.items()), reverse=True)
acc = sum(r[0] for r in rows)
print(json.dumps({"workload": name, "step_ms_instrumented": total, "sum_of_calls_ms": acc}))
for ms, n, k in rows[:24]:
    print(f"{ms:8.3f} ms {n:5d} calls {100 * ms / total:5.1f}%  {k}")

print("evfly_tc_conv_bf16 by shape (M_rows, Cin, Cout_rows, taps, convt, Hp, Wp):")
for k, v in sorted(shapes.items(), key=lambda kv: -sum(x.elapsed_time(y) for x, y in kv[1])):
    ms = sum(x.elapsed_time(y) for x, y in v)
    M, Cin, n_rows, taps = k[0], k[1], k[2], k[3]
    tf = 2.0 * M * Cin * taps * n_rows * len(v) / ms / 1e9
    print(f"{ms:8.3f} ms {len(v):3d} calls {tf:7.1f} TFLOP/s  {k}")

print("halo family by shape (entry, N, valid_h, valid_w, Cin, Cout):")
for k, v in sorted(halo_shapes.items(), key=lambda kv: -sum(x.elapsed_time(y) for x, y in kv[1])):
    ms = sum(x.elapsed_time(y) for x, y in v)
    _, N, vh, vw, Cin, Cout = k
    tf = 2.0 * N * (vh - 2) * (vw - 2) * Cin * 9 * Cout * len(v) / ms / 1e9
    print(f"{ms:8.3f} ms {len(v):3d} calls {tf:7.1f} TFLOP/s  {k}")

if "--each" in sys.argv:
    print("every call in launch order (ms):")
    for fn, e0, e1 in order:
        print(f"{e0.elapsed_time(e1):8.3f}  {fn}")
