"""evfly_b200 -- B200-native (sm_100a) implementation of evfly's perception hot path.

Event accumulation -> frame normalisation -> depth-pretext UNet/ConvLSTM -> ViT-LSTM velocity
forward, behind the reference's own Python surface (`form_eventframe`, the `nn.Module`
constructors / state_dict keys / forward signatures). All compute runs in hand-written CUDA
kernels reached through the C ABI of libevfly_b200.so (include/evfly_b200.h); there is no CPU
fallback. See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"
