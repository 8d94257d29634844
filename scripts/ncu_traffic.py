"""DRAM traffic per launch of the kernels bench.py's rooflines are about, from one ncu pass.

On the GPU box (under gpurun; one GPU, never a multi-rank command):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'k_tc_|k_convlstm|k_chunk|k_band|k_counts_norm|k_window' \
        --csv --log-file gpurun_out/r2_traffic_cfg4.csv python scripts/ncu_traffic.py run cfg4
    ncu ... --log-file gpurun_out/r2_traffic_cfg2.csv python scripts/ncu_traffic.py run cfg2
Here:
    python scripts/ncu_traffic.py parse gpurun_out/r2_traffic_cfg4.csv gpurun_out/r2_traffic_cfg2.csv
writes profiles/r2_dram_traffic.json (read by bench.py: roofline.traffic) and a per-kernel table next to it.
`run` executes 2 warm steps and 1 measured step of the bench workload without side-stream overlap; `parse` keeps the
last third of the launch list (= the measured step)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TC = re.compile(r"k_tc_conv|k_tc_stem|k_convlstm_scan")
ACC = re.compile(r"k_chunk_plan|k_chunk_sort|k_band_accumulate|k_window_ranges")
NORM = re.compile(r"k_counts_normalise")


def run(which):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if which == "cfg4":
        wl = bench.TrajectoryEvalWorkload(0, dev)
        wl.OVERLAP = False
        for i in range(3):
            wl.step(i)
    else:
        wl = bench.AccumulateWorkload(0, dev)
        for i in range(3):
            wl.step(i)
        for i in range(3):
            wl.L1.accumulate_windows(wl.d_a[i & 1], wl.edges_a, wl.H, wl.W, wl.B, counts=wl.counts_a, voxel=wl.voxel_a, algo="tiles")
    torch.cuda.synchronize()


def rows_of(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    launches = {}
    for r in csv.DictReader(lines[start:]):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
        d[r["Metric Name"]] = v * scale
    return [launches[k] for k in sorted(launches)]


def summarise(rows, pattern):
    sel = [r for r in rows if pattern.search(r["name"])]
    if not sel:
        return None
    rd = sum(r.get("dram__bytes_read.sum", 0) for r in sel)
    wr = sum(r.get("dram__bytes_write.sum", 0) for r in sel)
    t = sum(r.get("gpu__time_duration.sum", 0) for r in sel)
    return {"launches": len(sel), "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_launch": (rd + wr) / len(sel),
            "dram_bytes_total": rd + wr, "kernel_seconds_under_ncu": t}


def parse(paths):
    out, table = {}, []
    for p in paths:
        rows = rows_of(p)
        if "cfg4" in p:
            rows = rows[len(rows) * 2 // 3:]
            out["tc_conv_family"] = summarise(rows, TC)
            acc = summarise(rows, re.compile(ACC.pattern + "|" + NORM.pattern))
            out["accumulate_cfg4"] = acc
            tag = "cfg4 step (16 trajectories x 100 windows)"
        else:
            half = len(rows) // 2
            out["accumulate_cfg2b"] = summarise(rows[half * 2 // 3:half], ACC)
            out["accumulate_cfg2a"] = summarise(rows[half + (len(rows) - half) * 2 // 3:], ACC)
            tag = "cfg2"
        per = {}
        for r in rows:
            k = re.sub(r"\(.*", "", r["name"])
            a = per.setdefault(k, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += r.get("gpu__time_duration.sum", 0)
            a[2] += r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
        table.append(tag)
        tot = sum(a[1] for a in per.values())
        for k, a in sorted(per.items(), key=lambda kv: -kv[1][1]):
            table.append(f"  {a[1] * 1e3:9.3f} ms {100 * a[1] / tot:5.1f}%  {a[0]:5d} launches  {a[2] / 1e6:10.1f} MB dram  {k}")
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "r2_launches_by_kernel.txt"), "w").write("\n".join(table) + "\n")
    print(json.dumps(out, indent=1))
    print("\n".join(table))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        parse(sys.argv[2:])
