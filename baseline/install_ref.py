"""Make the reference's own hot-path sources importable on the GPU box: baseline/_ref/ (git-ignored, NOT gpurun-ignored,
so it travels with the snapshot; never part of the repo's history). The reference is plain Python without a
setup.py, so "installing" it is placing the five files of the path where
baseline/reference_arm.py imports them from. Called by __graft_entry__.build() when /root/reference is present.

    python baseline/install_ref.py
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["learner/learner_models.py", "learner/vitfly_models.py", "learner/ViTsubmodules.py",
         "learner/ConvLSTM_pytorch/convlstm.py", "utils/ev_utils.py"]


def install(ref_root="/root/reference") -> bool:
    if not os.path.isdir(ref_root):
        return os.path.isdir(DST)
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("installed" if install() else "reference tree not found", DST)
