"""Shared machinery of the drop-in nn.Modules: packed-weight caching and input plumbing."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class PackedModule(nn.Module):
    """An nn.Module whose parameters are only CONTAINERS (so constructor, initialisation,
    state_dict keys, .to(), .eval(), .parameters() behave exactly like the reference class) and
    whose forward runs libevfly_b200 kernels on weights packed once per parameter version."""

    def _weights_signature(self):
        sig = []
        for t in self.parameters(recurse=True):
            sig.append((t.data_ptr(), t._version))
        for t in self.buffers(recurse=True):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def packed(self):
        sig = self._weights_signature()
        cache = self.__dict__.get("_evfly_pack")
        if cache is None or cache[0] != sig:
            with torch.no_grad():
                cache = (sig, self._pack())
            self.__dict__["_evfly_pack"] = cache
        return cache[1]

    def _pack(self):  # -> anything; called under no_grad with parameters on their device
        raise NotImplementedError

    def _device(self) -> torch.device:
        p = next(self.parameters())
        if not p.is_cuda:
            raise _lib.EvflyError(
                f"{type(self).__name__} runs on a CUDA device only (B200, sm_100a); move it with .to('cuda'). "
                "evfly_b200 has no CPU fallback.")
        return p.device

    def _check_inference(self):
        """Forward-only, eval-folded kernels (Dropout off, BatchNorm running statistics folded into the conv, spectral
        norm without power iteration): a module left in train() mode would silently diverge from the reference, which
        uses train-mode semantics there, so that is refused whatever the grad mode; with eval() and grad enabled the
        outputs carry no autograd graph, which a fine-tuning loop must be told about."""
        if self.training:
            raise _lib.EvflyError(
                f"{type(self).__name__}: forward-only inference kernels with eval-mode semantics. Call .eval() first, as every "
                "inference call site of the reference does (run.py:261, run_competition.py:535, learner.py:755).")
        if torch.is_grad_enabled() and not PackedModule._warned_grad:
            PackedModule._warned_grad = True
            import warnings
            warnings.warn(f"{type(self).__name__}: evfly_b200 kernels are forward-only; outputs carry no autograd graph. "
                          "Run inference under torch.no_grad() (training through these modules is not supported).", stacklevel=3)

    _warned_grad = False


def to_dev(t, device):
    """Host->device plumbing for inputs the callers hand over as CPU tensors."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def sn_effective_weight(lin: nn.Module) -> torch.Tensor:
    """Eval-mode weight of an old-style spectral_norm-wrapped Linear: W_orig / (u^T W_orig v)
    (SURVEY.md 8(b)); folded once at pack time."""
    w = lin.weight_orig
    sigma = torch.dot(lin.weight_u, torch.mv(w.reshape(w.shape[0], -1), lin.weight_v))
    return (w / sigma).contiguous()


def bn_affine(bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm as y = x*scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    return scale.contiguous(), shift.contiguous()


def pack_lstm(lstm: nn.LSTM):
    """Per layer: (W_ih [4H,in], b_ih+b_hh or None, W_hh^T [H,4H], bf16 W_hh pairs or None, W_hh [4H,H])."""
    layers = []
    for l in range(lstm.num_layers):
        w_ih = getattr(lstm, f"weight_ih_l{l}").contiguous()
        w_hh_t = getattr(lstm, f"weight_hh_l{l}").t().contiguous()
        b = None
        if lstm.bias:
            b = (getattr(lstm, f"bias_ih_l{l}") + getattr(lstm, f"bias_hh_l{l}")).contiguous()
        pairs = None
        H = lstm.hidden_size
        if H % 2 == 0 and 4 * H <= 1024 and (H // 2) * 4 * H * 4 + 24 * H <= 220 * 1024:
            from . import tc
            pairs = tc.pack_lstm_whh_pairs(getattr(lstm, f"weight_hh_l{l}"))
        layers.append((w_ih, b, w_hh_t, pairs, getattr(lstm, f"weight_hh_l{l}").contiguous()))
    return layers


def run_lstm(ops, packed_layers, seq, state, hidden, smem_weights=False, n_seq=1):
    """nn.LSTM forward on an unbatched sequence seq [T,in] with state (h0,c0) [L,H] or None
    -> (out [T,H], (h [L,H], c [L,H])).
    n_seq > 1 (extension for trajectory batches): seq is the time-major batch [T*n_seq, in] (row t*n_seq + s),
    states are [L, n_seq, H] like PyTorch's batched LSTM, out is [T*n_seq, H]; one CTA per sequence."""
    L = len(packed_layers)
    dev = seq.device
    st_shape = (L, hidden) if n_seq == 1 else (L, n_seq, hidden)
    h_out = torch.empty(st_shape, dtype=torch.float32, device=dev)
    c_out = torch.empty(st_shape, dtype=torch.float32, device=dev)
    h0 = c0 = None
    if state is not None:
        h0, c0 = to_dev(state[0], dev), to_dev(state[1], dev)
        assert tuple(h0.shape) == st_shape and tuple(c0.shape) == st_shape, "LSTM state shape"
    lib = _lib.load()
    inp = seq
    T_all = seq.shape[0] // n_seq
    if T_all <= 4 and n_seq <= 8:
        # short sequences (batch-1 streaming): a persistent single-CTA scan would stream W_hh through one SM;
        # instead each step's gate GEMV runs on the multi-CTA small-M Linear, then one pointwise kernel
        for l, (w_ih, b, w_hh_t, pairs, w_hh) in enumerate(packed_layers):
            gx = ops.linear(inp, w_ih, b)                                   # [T*n_seq, 4H]
            hs = torch.empty((T_all * n_seq, hidden), dtype=torch.float32, device=dev)
            h_prev = None if h0 is None else h0[l].reshape(n_seq, hidden)
            c_prev = None if c0 is None else c0[l].reshape(n_seq, hidden)
            for t in range(T_all):
                g_t = gx[t * n_seq:(t + 1) * n_seq]
                if h_prev is not None:
                    g_t = ops.linear(h_prev, w_hh, None, res2d=g_t)
                last = t == T_all - 1
                h_t = hs[t * n_seq:(t + 1) * n_seq]
                c_t = c_out[l].reshape(n_seq, hidden) if last else torch.empty((n_seq, hidden), dtype=torch.float32, device=dev)
                _lib.check(lib.evfly_lstm_pointwise_f32(g_t.data_ptr(), None if c_prev is None else c_prev.data_ptr(), c_t.data_ptr(),
                                                        h_t.data_ptr(), h_out[l].data_ptr() if last else None, n_seq, hidden,
                                                        _lib.stream_ptr()), "evfly_lstm_pointwise_f32")
                h_prev, c_prev = h_t, c_t
            inp = hs
        return inp, (h_out, c_out)
    for l, (w_ih, b, w_hh_t, pairs, _w_hh) in enumerate(packed_layers):
        gx = ops.linear(inp, w_ih, b)
        T = gx.shape[0] // n_seq
        hs = torch.empty((T * n_seq, hidden), dtype=torch.float32, device=dev)
        args = (None if h0 is None else h0[l].data_ptr(), None if c0 is None else c0[l].data_ptr(),
                _lib.ptr(hs), h_out[l].data_ptr(), c_out[l].data_ptr(), T, hidden, n_seq, _lib.stream_ptr())
        if smem_weights and pairs is not None and T >= 16:     # bf16 W_hh resident in shared memory (bf16 path, long sequences)
            _lib.check(lib.evfly_lstm_seq_smemw(_lib.ptr(gx), pairs.data_ptr(), *args), "evfly_lstm_seq_smemw")
        else:
            _lib.check(lib.evfly_lstm_seq_f32(_lib.ptr(gx), _lib.ptr(w_hh_t), *args), "evfly_lstm_seq_f32")
        inp = hs
    return inp, (h_out, c_out)
