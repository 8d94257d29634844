"""evfly_b200 -- B200-native (sm_100a) implementation of evfly's perception hot path.

Event accumulation -> frame normalisation -> depth-pretext UNet/ConvLSTM -> ViT-LSTM velocity
forward, behind the reference's own Python surface (`form_eventframe`, the `nn.Module`
constructors / state_dict keys / forward signatures). All compute runs in hand-written CUDA
kernels reached through the C ABI of libevfly_b200.so (include/evfly_b200.h); there is no CPU
fallback. See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"


def set_precision(model, precision: str):
    """'fp32': exact CUDA-core path (matches the reference within rtol 1e-5, the default);
    'bf16': tcgen05 tensor-core path (rtol 1e-2). Applies to every sub-module that has both."""
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    for m in model.modules():
        if hasattr(type(m), "precision"):
            m.precision = precision
    return model
