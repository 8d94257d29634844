"""Build recipe of the C oracle (gcc only). TEST INFRASTRUCTURE.

`python -m oracle.build` -> oracle/libev_oracle.so (git-ignored, travels with gpurun).
oracle/_ref/ (the reference compiled from its own sources) does not exist for this project:
the only native reference code on the path, the two ROS nodes, needs roscpp and the
prophesee/dv message headers, which the image does not have -- see DESIGN.md.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ev_oracle.c")
LIB = os.path.join(HERE, "libev_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", LIB + ".tmp", SRC, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed: " + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
