"""Drop-in for evfly's learner/learner_models.py: DynamicConvNet, DynamicFCNet, VelPredictor,
OrigUNet, OrigUNet_w_VITFLY_ViTLSTM with the reference's constructor signatures, state_dict keys
and forward() contracts; the math runs in libevfly_b200 kernels.

Reference behaviours reproduced on purpose (they change results):
  * form_BEV == 0 feeds [pos, pos], not [|neg|, pos]: `zeros_like(x).expand(-1,2,-1,-1)` aliases
    the two channels (learner_models.py:479-481);
  * DynamicConvNet with invert_pool_input registers 'invert_{i}' twice under one name, so only the
    FIRST InvertLayer (before the pool) exists: the block computes pool(-x) (:76-93);
  * form_input zeroes |x| < cutoff IN the caller's tensor (:477).
The dead classes ConvUNet / OrigUNet_w_ConvNet_w_VelPred cannot be constructed in the reference
either and are not provided.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import LSTM

from . import ops, tc, vitfly_models
from ._modbase import PackedModule, bn_affine, pack_lstm, run_lstm, to_dev
from .ConvLSTM_pytorch.convlstm import ConvLSTM
from .vitfly_models import conv_transpose, pack_conv_transpose


def _conv_out_size(size, k, s, p=0):
    return (size + 2 * p - k) // s + 1


def find_output_size(model, input_size):
    """learner_models.py:8-12 runs a mock forward on torch.rand(input_size). The size is computed
    analytically here (no device needed at construction), but the same random tensor is drawn so
    that parameters created afterwards see the same RNG stream as in the reference."""
    torch.rand(input_size)
    return model.output_shape(input_size)


class InvertLayer(nn.Module):
    def forward(self, x):
        return -x


_ACT_MODULES = {'relu': nn.ReLU, 'sigmoid': nn.Sigmoid, 'tanh': nn.Tanh, 'leaky_relu': nn.LeakyReLU}


class DynamicConvNet(PackedModule):
    def __init__(self, in_channels, num_layers, kernel_sizes, kernel_strides, out_channels, activations, pool_type='max', pool_kernels=None, pool_strides=None, conv_function='conv2d', device=None, logger=None, invert_pool_input=False):
        super(DynamicConvNet, self).__init__()
        mylogger = logger if logger is not None else print
        self.layers = nn.Sequential()
        assert len(kernel_sizes) == num_layers, "The length of kernel_sizes should match num_layers"
        assert len(kernel_strides) == num_layers, "The length of kernel_strides should match num_layers"
        assert len(out_channels) == num_layers, "The length of out_channels should match num_layers"
        assert len(activations) == num_layers, "The length of activations should match num_layers"
        if pool_kernels is None:
            pool_kernels = [2] * num_layers
        if pool_strides is None:
            pool_strides = [2] * num_layers
        if conv_function == 'conv2d':
            self.conv_function = nn.Conv2d
        elif conv_function == 'upconv2d':
            self.conv_function = nn.ConvTranspose2d
        else:
            raise NotImplementedError(f'conv_function {conv_function} not implemented. Either use conv2d or upconv2d.')
        self._spec = []
        current_in_channels = in_channels
        for i in range(num_layers):
            self.layers.add_module(f'{conv_function}_{i}', self.conv_function(in_channels=current_in_channels, out_channels=out_channels[i], kernel_size=kernel_sizes[i], stride=kernel_strides[i], bias=False))
            self.layers.add_module(f'batchnorm_{i}', nn.BatchNorm2d(out_channels[i]))
            if activations[i] in _ACT_MODULES:
                self.layers.add_module(f'activation_{i}', _ACT_MODULES[activations[i]]())
            elif activations[i] != 'none':
                raise NotImplementedError(f'activation {activations[i]} not implemented. Either use relu, sigmoid, tanh, or leaky_relu.')
            if invert_pool_input:
                self.layers.add_module(f'invert_{i}', InvertLayer())
            pool = None
            if conv_function == 'conv2d':
                if pool_type == 'max':
                    self.layers.add_module(f'pool_{i}', nn.MaxPool2d(kernel_size=pool_kernels[i], stride=pool_strides[i]))
                    pool = ('max', pool_kernels[i], pool_strides[i])
                elif pool_type == 'avg':
                    self.layers.add_module(f'pool_{i}', nn.AvgPool2d(kernel_size=pool_kernels[i], stride=pool_strides[i]))
                    pool = ('avg', pool_kernels[i], pool_strides[i])
                elif pool_type != 'none':
                    raise NotImplementedError(f'pool_type {pool_type} not implemented. Either use max or avg.')
            if invert_pool_input:
                # same name as above: replaces that entry in place, so the Sequential holds ONE InvertLayer,
                # positioned before the pool
                self.layers.add_module(f'invert_{i}', InvertLayer())
            self._spec.append(dict(name=f'{conv_function}_{i}', bn=f'batchnorm_{i}', act=activations[i], k=kernel_sizes[i], s=kernel_strides[i],
                                   invert=bool(invert_pool_input), pool=pool, transposed=conv_function == 'upconv2d'))
            current_in_channels = out_channels[i]
        mylogger(f'[DynamicConvNet] Initialized DynamicConvNet with in_channels={in_channels}, num_layers={num_layers}, kernel_sizes={kernel_sizes}, kernel_strides={kernel_strides}, out_channels={out_channels}, activations={activations}, pool_type={pool_type}, pool_kernels={pool_kernels}, pool_strides={pool_strides}, conv_function={conv_function}')

    def output_shape(self, input_size):
        n, _, h, w = input_size
        c = None
        for sp in self._spec:
            conv = getattr(self.layers, sp['name'])
            c = conv.out_channels
            if sp['transposed']:
                h, w = (h - 1) * sp['s'] + sp['k'], (w - 1) * sp['s'] + sp['k']
            else:
                h, w = _conv_out_size(h, sp['k'], sp['s']), _conv_out_size(w, sp['k'], sp['s'])
            if sp['pool'] is not None:
                h, w = _conv_out_size(h, sp['pool'][1], sp['pool'][2]), _conv_out_size(w, sp['pool'][1], sp['pool'][2])
        return torch.Size([n, c, h, w])

    def _pack(self):
        packed = []
        for sp in self._spec:
            conv, bn = getattr(self.layers, sp['name']), getattr(self.layers, sp['bn'])
            scale, shift = bn_affine(bn)     # conv (no bias) -> BN folds into the conv
            packed.append(((conv.weight * scale.view(-1, 1, 1, 1)).contiguous(), shift))
        return packed

    def forward(self, x):
        self._check_inference()
        x = to_dev(x, self._device())
        pk = self.packed()
        for sp, (w, b) in zip(self._spec, pk):
            if sp['transposed']:
                raise NotImplementedError("DynamicConvNet(upconv2d) is only used by the reference's dead ConvUNet class")
            x = ops.conv2d(x, w, b, stride=sp['s'], act=sp['act'])
            if sp['pool'] is not None:
                x = ops.pool2d(x, sp['pool'][1], sp['pool'][2], sp['pool'][0], negate_in=sp['invert'])
            elif sp['invert']:
                x = ops.map4d(x, mul=-1.0)
        return x


class DynamicFCNet(PackedModule):
    def __init__(self, input_features, num_layers, layer_sizes, activations, dropout_p=None, device=None, logger=None):
        super(DynamicFCNet, self).__init__()
        mylogger = logger if logger is not None else print
        self.layers = nn.Sequential()
        assert len(layer_sizes) == num_layers, "The length of layer_sizes should match num_layers"
        assert len(activations) == num_layers, "The length of activations should match num_layers"
        self._spec = []
        current_input_features = input_features
        for i, layer_size in enumerate(layer_sizes):
            self.layers.add_module(f'fc_{i}', nn.Linear(current_input_features, layer_size))
            if dropout_p is not None and dropout_p > 0:
                self.layers.add_module(f'dropout_{i}', nn.Dropout(p=dropout_p))     # identity in eval
            if activations[i] not in _ACT_MODULES:
                raise NotImplementedError(f'activation {activations[i]} not implemented. Either use relu, sigmoid, tanh, or leaky_relu.')
            self.layers.add_module(f'activation_{i}', _ACT_MODULES[activations[i]]())
            self._spec.append((f'fc_{i}', activations[i]))
            current_input_features = layer_size
        mylogger(f'[DynamicFCNet] Initialized DynamicFCNet with input_features={input_features}, num_layers={num_layers}, layer_sizes={layer_sizes}, activations={activations}, dropout_p={dropout_p}')

    def forward(self, x):
        self._check_inference()
        x = to_dev(x, self._device())
        for name, act in self._spec:
            fc = getattr(self.layers, name)
            x = ops.linear(x, fc.weight, fc.bias, act=act)
        return x


class VelPredictor(PackedModule):
    def __init__(self, fc_params=None, input_size=512, num_out=3, device=None, logger=None):
        super().__init__()
        self.mylogger = logger if logger is not None else print
        self.input_size = input_size
        self.num_out = num_out
        self.device = device
        self.mylogger(f'[VelPredictor] Initializing VelPredictor with input_size={input_size} and num_out={num_out}')
        if fc_params is None:
            fc_params = {'num_layers': 3, 'layer_sizes': [128, 32, num_out], 'activations': ['leaky_relu', 'leaky_relu', 'tanh'], 'dropout_p': 0.1}
        self.fcnet = DynamicFCNet(input_features=input_size, num_layers=fc_params['num_layers'], layer_sizes=fc_params['layer_sizes'], activations=fc_params['activations'], dropout_p=fc_params['dropout_p'], logger=logger, device=device)

    def forward(self, X):
        x = to_dev(X[0], self._device())
        x = self.fcnet(x.reshape(x.shape[0], -1))
        if self.num_out == 1:        # [sqrt(1-y^2), y, 0]   (learner_models.py:321-333)
            x = ops.velpred_unit(x)
        elif self.num_out == 2:
            raise NotImplementedError("VelPredictor(num_out=2) is never constructed by the reference's OrigUNet (it passes num_out=1)")
        return x, None


# (big, small) of OrigUNet.skip per decoder level, learner_models.py:553-579
_SKIP_SIZES = [((25, 35), (16, 26)), ((58, 79), (24, 44)), ((124, 167), (40, 80)), ((256, 342), (72, 152))]


class OrigUNet(PackedModule):
    def __init__(self, num_in_channels=2, num_out_channels=1, num_recurrent=0, enc_params=None, dec_params=None, input_shape=[1, 2, 260, 346], device=None, logger=None, velpred=0, fc_params=None, form_BEV=0, is_deployment=False, is_large=False, evs_min_cutoff=1e-3, skip_type='crop'):
        super().__init__()
        mylogger = logger if logger is not None else print
        self.num_in_channels = num_in_channels
        self.num_out_channels = num_out_channels
        self.num_recurrent = num_recurrent
        self.input_shape = input_shape
        self.input_h, self.input_w = input_shape[-2], input_shape[-1]
        self.velpred = velpred
        self.fc_params = fc_params
        self.enc_params = enc_params
        self.device = device
        self.form_BEV = form_BEV
        self.evs_min_cutoff = evs_min_cutoff
        self.skip_type = skip_type
        self.decoder_numch_scalar = 1 if self.skip_type == 'none' else 2
        if self.form_BEV == 1 or self.form_BEV == 2:
            self.num_in_channels = 1
        elif self.form_BEV != 0:
            raise ValueError(f'form_BEV should be 0/1/2, but is {self.form_BEV}')
        self.is_deployment = is_deployment
        mylogger(f'[OrigUNet] Initializing OrigUNet with num_in_channels={self.num_in_channels}, num_out_channels={self.num_out_channels}, num_recurrent={self.num_recurrent}, form_BEV={self.form_BEV}, is_deployment={self.is_deployment}, evs_min_cutoff={self.evs_min_cutoff}, skip_type={self.skip_type}')

        self.unet_e11 = nn.Conv2d(self.num_in_channels, 32, kernel_size=3, padding=0)
        self.unet_e12 = nn.Conv2d(32, 32, kernel_size=3, padding=0)
        self.unet_pool1 = nn.MaxPool2d(kernel_size=2, stride=2,)
        self.unet_e21 = nn.Conv2d(32, 64, kernel_size=3, padding=0)
        self.unet_e22 = nn.Conv2d(64, 64, kernel_size=3, padding=0)
        self.unet_pool2 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.unet_e31 = nn.Conv2d(64, 128, kernel_size=3, padding=0)
        self.unet_e32 = nn.Conv2d(128, 128, kernel_size=3, padding=0)
        self.unet_pool3 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.unet_e41 = nn.Conv2d(128, 256, kernel_size=3, padding=0)
        self.unet_e42 = nn.Conv2d(256, 256, kernel_size=3, padding=0)
        self.unet_pool4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.unet_e51 = nn.Conv2d(256, 512, kernel_size=3, padding=0)
        self.unet_e52 = nn.Conv2d(512, 512, kernel_size=3, padding=0)
        self.unet_upconv1 = nn.ConvTranspose2d(512, 256, kernel_size=2, stride=2,)
        self.middle_shape = (1, 512, 8, 13)
        self.unet_d11 = nn.Conv2d(self.decoder_numch_scalar * 256, 256, kernel_size=3, padding=0)
        self.unet_d12 = nn.Conv2d(256, 256, kernel_size=3, padding=0)
        self.unet_upconv2 = nn.ConvTranspose2d(256, 128, kernel_size=2, stride=2,)
        self.unet_d21 = nn.Conv2d(self.decoder_numch_scalar * 128, 128, kernel_size=3, padding=0)
        self.unet_d22 = nn.Conv2d(128, 128, kernel_size=3, padding=0)
        self.unet_upconv3 = nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2,)
        self.unet_d31 = nn.Conv2d(self.decoder_numch_scalar * 64, 64, kernel_size=3, padding=0)
        self.unet_d32 = nn.Conv2d(64, 64, kernel_size=3, padding=0)
        self.unet_upconv4 = nn.ConvTranspose2d(64, 32, kernel_size=2, stride=2,)
        self.unet_d41 = nn.Conv2d(self.decoder_numch_scalar * 32, 32, kernel_size=3, padding=0)
        self.unet_d42 = nn.Conv2d(32, 32, kernel_size=3, padding=0)
        self.unet_out = nn.Conv2d(32, self.num_out_channels, kernel_size=1)
        self.nonlin = nn.ReLU()
        self.decoded_shape = (1, 1, 68, 148)

        if self.num_recurrent[0] > 0:
            mylogger(f'[OrigUNet] Using {self.num_recurrent[0]} recurrent layers')
            self.lstm = ConvLSTM(input_dim=self.middle_shape[1], hidden_dim=[self.middle_shape[1]] * self.num_recurrent[0], num_layers=self.num_recurrent[0], kernel_size=(1, 1), bias=False, batch_first=True, return_all_layers=False)

        if self.velpred > 0:
            if self.velpred == 1:
                mylogger(f'[OrigUNet] self.velpred == 1; Using velocity predictor with a ConvNet encoder and FC head.')
                in_ch, input_shape_enc = 1, torch.Size([1, 1, input_shape[-2], input_shape[-1]])
            elif self.velpred == 11:
                mylogger(f'[OrigUNet] self.velpred == 11; Using velocity predictor with a ConvNet encoder and FC head.')
                in_ch, input_shape_enc = self.decoded_shape[1], torch.Size(self.decoded_shape)
            elif self.velpred == 2:
                mylogger(f'[OrigUNet] self.velpred == 2; Using velocity predictor with a ConvNet encoder and ConvNet head.')
                in_ch, input_shape_enc = self.middle_shape[1], torch.Size(self.middle_shape)
            else:
                raise UnboundLocalError("velpred must be 0, 1, 11 or 2")   # the reference fails with a NameError-like error here too
            self.convnet_velpred = DynamicConvNet(in_channels=in_ch, num_layers=enc_params['num_layers'], kernel_sizes=enc_params['kernel_sizes'], kernel_strides=enc_params['kernel_strides'], out_channels=enc_params['out_channels'], activations=enc_params['activations'], pool_type=enc_params['pool_type'], pool_kernels=enc_params['pool_kernels'], pool_strides=enc_params['pool_strides'], conv_function=enc_params['conv_function'], invert_pool_input=enc_params['invert_pool_inputs'], logger=mylogger, device=device)
            mylogger(f'[OrigUNet] Input size to velpred: {list(input_shape_enc)}')
            self.convnet_velpred_outsize = find_output_size(self.convnet_velpred, input_shape_enc)
            mylogger(f'[OrigUNet] Calculated self.convnet_velpred_outsize = {self.convnet_velpred_outsize}')
            feat = self.convnet_velpred_outsize[1] * self.convnet_velpred_outsize[2] * self.convnet_velpred_outsize[3]
            if self.num_recurrent[1] > 0:
                self.lstm_velpred = LSTM(input_size=feat, hidden_size=feat, num_layers=self.num_recurrent[1], dropout=0.1)
                mylogger(f'[OrigUNet] LSTM for velocity prediction has {sum(p.numel() for p in self.lstm_velpred.parameters() if p.requires_grad):,} parameters.')
            self.velpred_head = VelPredictor(fc_params=fc_params, input_size=feat, num_out=1, device=device, logger=mylogger)
            mylogger(f'[OrigUNet] ConvNet for velocity prediction has {sum(p.numel() for p in self.convnet_velpred.parameters() if p.requires_grad):,} parameters.')
            mylogger(f'[OrigUNet] FCNet for velocity prediction has {sum(p.numel() for p in self.velpred_head.fcnet.parameters() if p.requires_grad):,} parameters.')

    # ---- weights ----------------------------------------------------------------------------
    precision = 'fp32'     # 'fp32': exact CUDA-core path (rtol 1e-5); 'bf16': tcgen05 tensor-core path (rtol 1e-2)

    def _pack(self):
        pk = {f"up{i}": pack_conv_transpose(getattr(self, f"unet_upconv{i}")) for i in range(1, 5)}
        if self.velpred > 0 and self.num_recurrent[1] > 0:
            pk["lstm_velpred"] = pack_lstm(self.lstm_velpred)
        # bf16 operands of the tensor-core path: [Cout][tap][Cin] for the 3x3 convs,
        # [phase][Cout][Cin] for the transposed convs, x/h halves of the ConvLSTM gate conv
        bf = {}
        for name in ("e12", "e21", "e22", "e31", "e32", "e41", "e42", "e51", "e52",
                     "d11", "d12", "d21", "d22", "d31", "d32", "d41", "d42"):
            bf[name] = tc.pack_conv3x3_weight(getattr(self, "unet_" + name).weight)
        for i in range(1, 5):
            bf[f"up{i}"] = tc.pack_convt2x2_weight(getattr(self, f"unet_upconv{i}").weight)
        bf["out"] = tc.pack_conv1x1_weight(self.unet_out.weight)
        bf["out_w1"] = self.unet_out.weight.reshape(-1).to(tc.BF16).float().contiguous()     # bf16-rounded, for d42's epilogue
        if self.num_recurrent[0] > 0:
            bf["lstm"] = []
            for cell in self.lstm.cell_list:
                w = cell.conv.weight
                w2 = w.reshape(w.shape[0], -1)      # 1x1 kernel; rows gate-major -> interleaved (fused cell epilogue)
                bf["lstm"].append((tc.pack_convlstm_gate_weight(w2[:, :cell.input_dim]), tc.pack_convlstm_gate_weight(w2[:, cell.input_dim:])))
        pk["bf16"] = bf
        pk["stage"] = self._pack_stage(bf) if self._stage_ok() and self.unet_e11.weight.is_cuda else None
        return pk

    # ---- stage-level C ABI (csrc/stages.cu): the whole forward enqueued by ONE call -------------------
    def _stage_ok(self):
        """The shipped configuration (learner/configs/*.txt:39-47), for which evfly_unet_forward is written."""
        return (self.form_BEV == 2 and self.skip_type == 'interp' and self.num_in_channels == 1 and self.num_out_channels == 1 and self.velpred == 0
                and not self.is_deployment and list(self.num_recurrent)[0] == 1)

    def _pack_stage(self, bf):
        from . import _lib
        w = _lib.UNetWeights()
        keep = []           # tensors the struct points into

        def p(t):
            t = t.contiguous()
            keep.append(t)
            return t.data_ptr()
        w.e11_w, w.e11_b = p(self.unet_e11.weight.float()), p(self.unet_e11.bias.float())
        for i, name in enumerate(("e12", "e21", "e22", "e31", "e32", "e41", "e42", "e51", "e52", "d11", "d12", "d21", "d22", "d31", "d32", "d41", "d42")):
            w.conv_w[i], w.conv_b[i] = p(bf[name]), p(getattr(self, "unet_" + name).bias.float())
        for i in range(4):
            w.up_w[i], w.up_b[i] = p(bf[f"up{i + 1}"]), p(getattr(self, f"unet_upconv{i + 1}").bias.float())
        w.out_w, w.out_b = p(bf["out_w1"]), p(self.unet_out.bias.float())
        w.lstm_wx, w.lstm_wh = p(bf["lstm"][0][0]), p(bf["lstm"][0][1])
        return w, keep

    _stage_ws: dict = {}

    def _unet_stage(self, frames, state, stage, n_traj):
        """OrigUNet.forward through evfly_unet_forward: one C call enqueues form_input, encoder, ConvLSTM scan, decoder
        and the output resize on the current stream (same kernels, same order as _unet_bf16)."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        N, _, H, W = frames.shape
        dev = frames.device
        h, w = H, W
        for _ in range(4):
            h, w = (h - 4) // 2, (w - 4) // 2
        vh5, vw5 = h - 4, w - 4
        vh, vw = vh5, vw5
        for _ in range(4):
            vh, vw = 2 * vh - 4, 2 * vw - 4
        need = lib.evfly_unet_workspace_bytes(N, n_traj, H, W)
        if need <= 0:
            raise _lib.EvflyError(f"evfly_unet_forward: unsupported shape N={N} n_traj={n_traj} H={H} W={W}")
        key = (dev, torch.cuda.current_stream().cuda_stream)
        ws = OrigUNet._stage_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = OrigUNet._stage_ws[key] = torch.empty((need,), dtype=torch.uint8, device=dev)
        depth = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
        yu = torch.empty((N, 1, vh, vw), dtype=torch.float32, device=dev)
        hT = torch.empty((n_traj, 512, vh5, vw5), dtype=torch.float32, device=dev)
        cT = torch.empty_like(hT)
        h0 = c0 = None
        if state is not None:
            h0, c0 = to_dev(state[0][0], dev), to_dev(state[0][1], dev)
            assert tuple(h0.shape) == tuple(hT.shape) and tuple(c0.shape) == tuple(cT.shape), "ConvLSTM state shape"
        _lib.check(lib.evfly_unet_forward(C.byref(stage[0]), _lib.ptr(frames), N, n_traj, H, W, float(self.evs_min_cutoff), _lib.ptr(h0), _lib.ptr(c0),
                                          _lib.ptr(hT), _lib.ptr(cT), _lib.ptr(depth), _lib.ptr(yu), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "evfly_unet_forward")
        return None, [[hT, cT]], yu, depth

    # ---- pieces of the reference API ------------------------------------------------------------
    def form_input(self, x):
        """learner_models.py:476-494 (in place on x, like the reference)."""
        return ops.form_input(x, self.form_BEV, float(self.evs_min_cutoff))

    def form_output(self, x):
        if self.num_out_channels == 2:
            raise NotImplementedError("num_out_channels == 2 is not used by any shipped configuration")
        return ops.resize_bilinear(x, (self.input_h, self.input_w), align_corners=False), x

    def skip(self, y, big, small, out_view=None):
        if self.skip_type == 'crop':
            crop = y[:, :, big[0] // 2 - small[0] // 2: big[0] // 2 + small[0] // 2, big[1] // 2 - small[1] // 2: big[1] // 2 + small[1] // 2]
            return ops.map4d(crop, out_view)
        elif self.skip_type == 'interp':
            return ops.resize_bilinear(y, (small[0], small[1]), align_corners=False, out_view=out_view)
        elif self.skip_type == 'none':
            return None
        raise ValueError(f'[LEARNER_MODELS/ORIGUNET] skip_type should be crop/interp/none, but is {self.skip_type}.')

    def _c(self, name, x, act="relu", **kw):
        m = getattr(self, name)
        return ops.conv2d(x, m.weight, m.bias, act=act, **kw)

    # ---- exact path: CUDA-core fp32 kernels --------------------------------------------------------
    def _unet_fp32(self, im, state, pk, n_traj=1):
        N, dev = im.shape[0], im.device
        # encoder: (3x3 valid conv + ReLU) x 2 per level, 2x2 max-pool between levels
        y_e1 = self._c("unet_e12", self._c("unet_e11", im))
        y_e2 = self._c("unet_e22", self._c("unet_e21", ops.pool2d(y_e1, 2, 2)))
        y_e3 = self._c("unet_e32", self._c("unet_e31", ops.pool2d(y_e2, 2, 2)))
        y_e4 = self._c("unet_e42", self._c("unet_e41", ops.pool2d(y_e3, 2, 2)))
        y_e5 = self._c("unet_e52", self._c("unet_e51", ops.pool2d(y_e4, 2, 2)))

        h_unet = None
        if self.num_recurrent[0] > 0:
            if n_traj == 1:
                y_e5_lstm, h_unet = self.lstm(y_e5.unsqueeze(0), state)
                y_e5 = y_e5_lstm[0].squeeze(0)
            else:   # time-major batch of trajectories [T*n, C, h, w] -> [n, T, C, h, w] for the (batch_first) ConvLSTM
                T = N // n_traj
                seq = y_e5.view(T, n_traj, *y_e5.shape[1:]).transpose(0, 1)
                y_e5_lstm, h_unet = self.lstm(seq, state)
                y_e5 = y_e5_lstm[0].transpose(0, 1).reshape(N, *y_e5.shape[1:])

        y_upconv = None
        y_interp = None
        if not self.is_deployment or (self.is_deployment and (self.velpred == 1 or self.velpred == 11)):
            y = y_e5
            for lvl, enc in enumerate((y_e4, y_e3, y_e2, y_e1), start=1):
                up = getattr(self, f"unet_upconv{lvl}")
                C = up.out_channels
                big, small = _SKIP_SIZES[lvl - 1]
                oh, ow = 2 * y.shape[2], 2 * y.shape[3]
                if self.skip_type == 'none':
                    cat = torch.empty((N, C, oh, ow), dtype=torch.float32, device=dev)
                    conv_transpose(y, pk[f"up{lvl}"], up.bias, cat)
                else:   # torch.cat((skipped_enc, upconv(y)), 1) assembled in place
                    cat = torch.empty((N, 2 * C, oh, ow), dtype=torch.float32, device=dev)
                    self.skip(enc, big, small, out_view=cat[:, :C])
                    conv_transpose(y, pk[f"up{lvl}"], up.bias, cat[:, C:])
                y = self._c(f"unet_d{lvl}2", self._c(f"unet_d{lvl}1", cat))
            y_upconv = self._c("unet_out", y, act=None)
            y_interp, y_upconv = self.form_output(y_upconv)
        return (lambda: y_e5), h_unet, y_upconv, y_interp

    # ---- fast path: bf16 NHWC pitch grids on the tensor cores -----------------------------------
    def _unet_bf16(self, im, state, W, n_traj=1):
        N, dev = im.shape[0], im.device
        b = lambda name: getattr(self, "unet_" + name).bias
        cv = lambda g, name: tc.conv3x3(g, W[name], b(name), relu=True, compact=True)      # wide layers write compact grids (no don't-care rows downstream)
        # 'interp' skip: y_e1..y_e3 are only sampled by the decoder's bilinear resize to the heights below (the decoder level that
        # takes y_e{k} works at twice the valid height of its input, which shrinks by 4 per level), so the fused-pool convs write
        # only the rows it reads
        skip_oh = {}
        if self.skip_type == 'interp' and (not self.is_deployment or self.velpred in (1, 11)):
            h = im.shape[2]
            for _ in range(4):
                h = (h - 4) // 2
            vh = h - 4                                   # valid height of y_e5
            for k in (4, 3, 2, 1):
                skip_oh[k] = 2 * vh
                vh = 2 * vh - 4
        cvp = lambda g, name, k=None: tc.conv3x3_pool(g, W[name], b(name), relu=True, skip_rows=skip_oh.get(k))     # conv + MaxPool2d(2), fused where it can be
        if self.form_BEV == 2 and tc.FUSE_STEM and tc.USE_HALO and tc.FUSE_POOL and im.shape[1] == 1:
            # binary mask: unet_e11 is a 512-entry table lookup inside the e12 kernel, e11 never goes to HBM
            y_e1, p1 = tc.stem_e12_pool(im, self.unet_e11.weight, self.unet_e11.bias, W["e12"], b("e12"),
                                        frames_cutoff=float(self.evs_min_cutoff) if getattr(self, "_fold_form_input", False) else None,
                                        skip_rows=skip_oh.get(1))
        else:
            y_e1, p1 = cvp(tc.stem_conv3x3(im, self.unet_e11.weight, self.unet_e11.bias), "e12", 1)
        y_e2, p2 = cvp(cv(p1, "e21"), "e22", 2)
        y_e3, p3 = cvp(cv(p2, "e31"), "e32", 3)
        y_e4, p4 = cvp(cv(p3, "e41"), "e42")
        y_e5 = cv(cv(p4, "e51"), "e52")

        h_unet = None
        if self.num_recurrent[0] > 0:
            y_e5, h_unet = self._convlstm_bf16(y_e5, state, W["lstm"], n_traj)

        y_upconv = None
        y_interp = None
        if not self.is_deployment or (self.is_deployment and (self.velpred == 1 or self.velpred == 11)):
            y = y_e5
            for lvl, enc in enumerate((y_e4, y_e3, y_e2, y_e1), start=1):
                up = getattr(self, f"unet_upconv{lvl}")
                C = up.out_channels
                big, small = _SKIP_SIZES[lvl - 1]
                oh, ow = 2 * y.vh, 2 * y.vw
                if self.skip_type == 'none':
                    cat = torch.empty((N, oh, ow, C), dtype=tc.BF16, device=dev)
                    tc.conv_transpose2x2(y, W[f"up{lvl}"], up.bias, cat, 0)
                else:
                    cat = torch.empty((N, oh, ow, 2 * C), dtype=tc.BF16, device=dev)
                    if self.skip_type == 'crop':
                        tc.crop_into(enc, big[0] // 2 - small[0] // 2, big[1] // 2 - small[1] // 2, oh, ow, cat, 0)
                    elif self.skip_type == 'interp':
                        tc.resize_bilinear_into(enc, oh, ow, cat, 0)
                    else:
                        raise ValueError(f'[LEARNER_MODELS/ORIGUNET] skip_type should be crop/interp/none, but is {self.skip_type}.')
                    tc.conv_transpose2x2(y, W[f"up{lvl}"], up.bias, cat, C)
                if lvl == 4 and self.num_out_channels == 1 and tc.USE_HALO:
                    # unet_d42 with unet_out (1x1 to one channel) in its epilogue: the 32-channel activation is never written
                    y1 = cv(tc.Grid(cat, oh, ow), "d41")
                    out32 = tc.conv3x3_out1(y1, W["d42"], b("d42"), W["out_w1"], self.unet_out.bias)
                    y = tc.Grid(out32.view(N, y1.Hp, y1.Wp, 1), oh - 4, ow - 4)
                else:
                    y = cv(cv(tc.Grid(cat, oh, ow), f"d{lvl}1"), f"d{lvl}2")
            if self.num_out_channels != 1:
                raise NotImplementedError("num_out_channels == 2 is not used by any shipped configuration")
            if y.data.dtype == torch.float32:
                out32 = y.data
            else:
                out32 = torch.empty((y.rows, 1), dtype=torch.float32, device=dev)
                tc.gemm(y.data.view(y.rows, y.C), W["out"], self.unet_out.bias, out_f32=out32)
            y_upconv = tc.grid_to_nchw(out32.view(N, y.Hp, y.Wp, 1), y.vh, y.vw)
            y_interp = ops.resize_bilinear(y_upconv, (self.input_h, self.input_w), align_corners=False)
        return (lambda: tc.grid_to_nchw(y_e5.data, y_e5.vh, y_e5.vw)), h_unet, y_upconv, y_interp

    def _convlstm_bf16(self, g, state, Wl, n_traj=1):
        """ConvLSTM (1x1 kernel, no bias) over the time axis on the pitch grid: the x half of the gate conv is ONE
        tensor-core GEMM over all steps; the h half is one small GEMM per step whose epilogue adds the x gates and
        performs the cell update (fp32 c, bf16 h), all T steps enqueued by one C-ABI call.
        g holds T*n_traj frames in time-major order (frame t*n_traj + s): the n_traj trajectories advance together,
        so a step's GEMM has n_traj * Hp*Wp rows. States are [n_traj, Ch, vh, vw] (n_traj = 1: the reference's)."""
        Hp, Wp, dev = g.Hp, g.Wp, g.data.device
        T = g.N // n_traj
        P = n_traj * Hp * Wp            # rows advanced per step
        states = []
        cur = g
        for li, (wx, wh) in enumerate(Wl):
            Ch = wh.shape[1]
            c = torch.zeros((P, Ch), dtype=torch.float32, device=dev)
            h_all = torch.empty((T + 1, P, Ch), dtype=tc.BF16, device=dev)      # block 0 = h_0, block t+1 = h_t
            if state is not None:
                hs, cs = to_dev(state[li][0], dev), to_dev(state[li][1], dev)        # [n_traj,Ch,vh,vw] each
                h_all[0].copy_(tc.nchw_to_grid(hs, Hp, Wp).data.view(P, Ch))
                ops.map4d(cs.permute(0, 2, 3, 1), c.view(n_traj, Hp, Wp, Ch)[:, :g.vh, :g.vw])
            else:
                h_all[0].zero_()
            if not tc.convlstm_scan_fused(cur.data.view(T * P, cur.C), wx, h_all, wh, c, T, P, Ch):    # x-gates inside the step
                gx = torch.empty((T * P, 4 * Ch), dtype=torch.float32, device=dev)
                tc.gemm(cur.data.view(T * P, cur.C), wx, None, out_f32=gx)
                tc.convlstm_scan(h_all, wh, gx, c, T, P, Ch)     # T fused step kernels enqueued from C++
            out = tc.Grid(h_all[1:].view(T * n_traj, Hp, Wp, Ch), g.vh, g.vw)
            h_last = tc.grid_to_nchw(h_all[T].view(n_traj, Hp, Wp, Ch), g.vh, g.vw)
            c_last = torch.empty((n_traj, Ch, g.vh, g.vw), dtype=torch.float32, device=dev)
            ops.map4d(c.view(n_traj, Hp, Wp, Ch)[:, :g.vh, :g.vw].permute(0, 3, 1, 2), c_last)
            states.append([h_last, c_last])
            cur = out
        return cur, states[-1:]

    def forward_trajectories(self, x, n_traj):
        """Extension for config 4: n_traj independent trajectories of equal length advance together.
        x[0] holds T*n_traj frames in TIME-MAJOR order (frame t*n_traj + s belongs to trajectory s); recurrent states
        carry a leading n_traj dimension ([n_traj,512,8,13] for the ConvLSTM). Same outputs as forward()."""
        return self.forward(x, n_traj=n_traj)

    def forward(self, x, n_traj=1):
        """x = [frames [N,1,H,W], desvel (unused), [h_unet, h_velpred] or None]
        -> (vel [N,3], (y_interp, y_upconv, (h_unet, h_velpred)))   (learner_models.py:521-616)"""
        self._check_inference()
        dev = self._device()
        pk = self.packed()
        im = x[0] = to_dev(x[0], dev)
        N = im.shape[0]
        if x[2] is None:
            x[2] = (None, None)
        use_stage = self.precision == 'bf16' and tc.USE_STAGE_ABI and pk["stage"] is not None and tuple(im.shape[1:]) == (1, self.input_h, self.input_w)
        # binary input on the bf16 path: form_input is folded into the stem's pattern extraction (same in-place cutoff)
        self._fold_form_input = (not use_stage and self.precision == 'bf16' and self.form_BEV == 2 and tc.FUSE_STEM and tc.USE_HALO and tc.FUSE_POOL
                                 and im.shape[1] == 1)
        if not use_stage and not self._fold_form_input and (self.num_in_channels == 2 or self.form_BEV > 0):
            im = self.form_input(im)

        if use_stage:
            y_e5_nchw, h_unet, y_upconv, y_interp = self._unet_stage(im, x[2][0], pk["stage"], n_traj)
        elif self.precision == 'bf16':
            y_e5_nchw, h_unet, y_upconv, y_interp = self._unet_bf16(im, x[2][0], pk["bf16"], n_traj)
        else:
            y_e5_nchw, h_unet, y_upconv, y_interp = self._unet_fp32(im, x[2][0], pk, n_traj)

        y_vel = torch.zeros((N, 3), dtype=torch.float32, device=dev)   # default [1,0,0]: forward, full speed
        y_vel[:, 0] = 1.0
        h_velpred = None
        if self.velpred > 0:
            src = {1: y_interp, 11: y_upconv, 2: y_e5_nchw}[self.velpred]
            if callable(src):
                src = src()
            feat = self.convnet_velpred(src)
            feat = feat.reshape(N, -1)
            if self.num_recurrent[1] > 0:
                feat, h_velpred = run_lstm(ops, pk["lstm_velpred"], feat, x[2][1], self.lstm_velpred.hidden_size, n_seq=n_traj)
            y_vel, _ = self.velpred_head([feat])
        return y_vel, (y_interp, y_upconv, (h_unet, h_velpred))


class OrigUNet_w_VITFLY_ViTLSTM(nn.Module):
    def __init__(self, num_in_channels=2, num_out_channels=1, num_recurrent=0, enc_params=None, dec_params=None, input_shape=[1, 2, 260, 346], device=None, logger=None, old_model=False, velpred=False, fc_params=None, form_BEV=0, is_deployment=False, evs_min_cutoff=1e-3, skip_type='crop'):
        super().__init__()
        # evs -> depth
        self.origunet = OrigUNet(num_in_channels=num_in_channels, num_out_channels=num_out_channels, num_recurrent=num_recurrent, enc_params=enc_params, dec_params=dec_params, input_shape=input_shape, device=device, logger=logger, velpred=velpred, fc_params=fc_params, form_BEV=form_BEV, is_deployment=is_deployment, evs_min_cutoff=evs_min_cutoff, skip_type=skip_type)
        # depth -> vel
        self.vitfly_vitlstm = vitfly_models.LSTMNetVIT()
        print(f'[OrigUNet_w_VITFLY_ViTLSTM] Number of parameters: {sum(p.numel() for p in self.parameters()):,}')

    def forward_trajectories(self, X, n_traj):
        """Extension for config 4: n_traj trajectories advance together; frames / desvel are time-major
        (row t*n_traj + s), ConvLSTM states [n_traj,512,8,13], LSTM states [3,n_traj,128]."""
        return self.forward(X, n_traj=n_traj)

    def forward(self, X, n_traj=1):
        """X = [frames, desvel, [h_unet, None] or None, (h,c) or None]
        -> (vel, (depth, y_upconv, ((h_unet, h_velpred), (h,c))))   (learner_models.py:629-636)"""
        x = X[0]
        _, (x_depth, y_upconv, (h_unet, h_velpred)) = self.origunet.forward([x, None, X[2]], n_traj=n_traj)
        # * 2 roughly matches the depth scale VITFLY_ViTLSTM was trained on (:634)
        v = self.vitfly_vitlstm
        if v._stage_usable(x_depth.shape[0]):
            # clamp + resize + both ViT stages + LSTM + head enqueued by one C call (evfly_vit_lstm_forward)
            v._check_inference()
            x_vel, h_vitlstm = v.forward_from_depth(x_depth, X[1], None, X[3], n_traj=n_traj, premap_clamp=True)
            return x_vel, (x_depth, y_upconv, ((h_unet, h_velpred), h_vitlstm))
        if tuple(x_depth.shape[-2:]) != (60, 90):
            # vitfly's refine_inputs (vitfly_models.py:28-29) would resize the clamped depth to 60x90 next: do both in
            # one pass over the four samples of each output pixel instead of materialising the full-size clamped image
            x_depth_input = ops.resize_bilinear(x_depth, (60, 90), align_corners=False, pre=(2.0, 0.0, 1.0))
        else:
            x_depth_input = ops.map4d(x_depth, mul=2.0, lo=0.0, hi=1.0)
        x_vel, h_vitlstm = self.vitfly_vitlstm.forward([x_depth_input, X[1], None, X[3]], n_traj=n_traj)
        return x_vel, (x_depth, y_upconv, ((h_unet, h_velpred), h_vitlstm))
