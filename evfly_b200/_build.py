"""In-tree build of libevfly_b200.so with nvcc for sm_100a (B200).

`python -m evfly_b200._build [--force] [-v]` or `__graft_entry__.build()`; nvcc cross-compiles without a
GPU. Every .cu under csrc/ is compiled to its own object (in parallel, rebuilt only when it or a header
changed) and the objects are linked into the shared library. The .so is git-ignored but travels to the GPU
box with the gpurun snapshot, together with build_info.json: what was compiled, when, from which source
hash and whether this call compiled or reused it -- bench.py re-hashes csrc/ and reports whether the
library it loaded matches the sources next to it.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libevfly_b200.so")
INFO_PATH = os.path.join(PKG_DIR, "build_info.json")
HEADER = os.path.join(ROOT, "include", "evfly_b200.h")
# test-only library: the hardware-assumption probe of tests/test_tc_gpu.py is not part of the product ABI
PROBE_SRC = os.path.join(ROOT, "tests", "native", "tc_probe.cu")
PROBE_LIB = os.path.join(ROOT, "tests", "native", "libevfly_tc_probe.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = [*ARCH_FLAGS, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [HEADER]


def source_hash() -> str:
    h = hashlib.sha256()
    for p in sources() + _headers():
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def read_info() -> dict | None:
    try:
        return json.load(open(INFO_PATH))
    except Exception:
        return None


def needs_build() -> bool:
    info = read_info()
    return not os.path.exists(LIB_PATH) or info is None or info.get("source_hash") != source_hash()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libevfly_b200.so cannot be built")
    return nvcc


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    if not force and not needs_build():
        info = read_info() or {}
        info["last_call"] = "reused (sources unchanged since the recorded compile)"
        json.dump(info, open(INFO_PATH, "w"), indent=1)
        _build_probe(nvcc, force=False)
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(p) for p in _headers())
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_mtime)
        jobs.append((src, obj, stale))
    t0 = time.time()

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return ""
        cmd = [nvcc, *COMMON_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        return _run(cmd)

    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        sys.stderr.write("".join(logs))
    tmp = LIB_PATH + ".tmp"
    _run([nvcc, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-soname", "-Xlinker", "libevfly_b200.so", "-o", tmp, *[j[1] for j in jobs]])
    os.replace(tmp, LIB_PATH)
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    info = {"source_hash": source_hash(), "compiled_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
            "nvcc": ver, "flags": " ".join(COMMON_FLAGS), "objects_compiled": sum(1 for j in jobs if j[2]), "objects_total": len(jobs),
            "seconds": round(time.time() - t0, 1), "last_call": "compiled"}
    json.dump(info, open(INFO_PATH, "w"), indent=1)
    _build_probe(nvcc, force=True)
    return LIB_PATH


DRIVER_SRC = os.path.join(ROOT, "tests", "native", "stage_driver.c")
DRIVER_BIN = os.path.join(ROOT, "tests", "native", "stage_driver")


def _build_driver(nvcc: str, force: bool) -> None:
    """The plain-C program that drives the stage-level ABI (tests/test_stage_abi_gpu.py)."""
    if not os.path.exists(DRIVER_SRC):
        return
    if not force and os.path.exists(DRIVER_BIN) and os.path.getmtime(DRIVER_BIN) >= max(os.path.getmtime(DRIVER_SRC), os.path.getmtime(HEADER)):
        return
    _run([nvcc, "-x", "c", DRIVER_SRC, "-o", DRIVER_BIN, "-L", PKG_DIR, "-l:libevfly_b200.so", "-lcudart",
          "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../evfly_b200"])


def _build_probe(nvcc: str, force: bool) -> None:
    _build_driver(nvcc, force)
    if not os.path.exists(PROBE_SRC):
        return
    if not force and os.path.exists(PROBE_LIB) and os.path.getmtime(PROBE_LIB) >= os.path.getmtime(PROBE_SRC):
        return
    _run([nvcc, *COMMON_FLAGS, "-shared", "-I", CSRC, "-o", PROBE_LIB, PROBE_SRC, "-L", PKG_DIR, "-l:libevfly_b200.so",
          "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../evfly_b200"])


def build_state() -> dict:
    """What bench.py reports: the record of the compile that produced the loaded library and whether the
    sources next to it still hash to what was compiled."""
    info = read_info() or {}
    return {"compiled_at": info.get("compiled_at"), "nvcc": info.get("nvcc"), "source_hash": info.get("source_hash"),
            "matches_sources": info.get("source_hash") == source_hash(), "last_build_call": info.get("last_call")}


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
