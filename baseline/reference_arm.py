"""The reference's OWN code path for the hot path, on the host CPU: utils/ev_utils.py::form_eventframe per window
(np.histogram2d), run.py:250-253's percentile scaling, and learner_models.OrigUNet_w_VITFLY_ViTLSTM.forward in the
deployed configuration -- imported unmodified from baseline/_ref/ (see install_ref.py). Used by bench.py's
`--impl reference` arm and `cpu_baseline` leg only. No code of evfly_b200 runs here."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

ENC = dict(num_layers=2, kernel_sizes=[5, 3], kernel_strides=[2, 2], out_channels=[8, 32], activations=["relu", "relu"],
           pool_type="max", invert_pool_inputs=True, pool_kernels=[2, 2], pool_strides=[2, 2], conv_function="conv2d")
FC = dict(num_layers=4, layer_sizes=[1024, 128, 16, 1], activations=["leaky_relu"] * 3 + ["tanh"], dropout_p=0.1)


def available() -> bool:
    return os.path.exists(os.path.join(REF, "learner", "learner_models.py")) and os.path.exists(os.path.join(REF, "utils", "ev_utils.py"))


class ReferenceArm:
    def __init__(self, state_dict):
        import contextlib
        import io
        import torch
        for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):      # plotting helpers only
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
        saved = list(sys.path)
        sys.path.insert(0, os.path.join(REF, "learner"))
        sys.path.insert(0, os.path.join(REF, "utils"))
        try:
            for m in ("learner_models", "vitfly_models", "ViTsubmodules", "ev_utils", "ConvLSTM_pytorch", "ConvLSTM_pytorch.convlstm"):
                sys.modules.pop(m, None)
            import ev_utils
            import learner_models
        finally:
            sys.path[:] = saved
        self.torch, self.ev_utils = torch, ev_utils
        with contextlib.redirect_stdout(io.StringIO()):
            # the shipped configuration: learner/configs/eval_config_real.txt:39-47
            m = learner_models.OrigUNet_w_VITFLY_ViTLSTM(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346],
                                                         logger=lambda *a, **k: None, velpred=0, enc_params=ENC, fc_params=FC, form_BEV=2,
                                                         evs_min_cutoff=1e-3, skip_type="interp", is_deployment=False)
        m.load_state_dict(state_dict, strict=True)
        self.model = m.eval().float()
        torch.set_num_threads(os.cpu_count())
        self.cores = torch.get_num_threads()

    def trajectory(self, rows_per_window, H, W, desvel=4.0):
        """rows_per_window: list of float64 [n,4] = (t, x, y, p) arrays, one per window (what data_gather/ hands to
        form_eventframe). Returns vel [T,3]."""
        torch = self.torch
        frames = []
        for rows in rows_per_window:
            fr = self.ev_utils.form_eventframe(rows, H, W, all_events=True)                       # utils/ev_utils.py:150-161
            x = torch.from_numpy(fr).view(1, 1, H, W).float()
            q = torch.quantile(x.abs(), .97)                                                      # evfly_ros/run.py:250
            frames.append(torch.clip(x / q, -1.0, 1.0))                                           # run.py:253
        x = torch.cat(frames)
        with torch.no_grad():
            vel, _ = self.model([x, torch.full((x.shape[0], 1), desvel), [None, None], None])    # learner/evaluation_tools.py:62-66
        return vel
