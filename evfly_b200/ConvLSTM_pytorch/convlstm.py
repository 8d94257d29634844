"""Drop-in for evfly's learner/ConvLSTM_pytorch/convlstm.py (ConvLSTMCell, ConvLSTM): same
constructors, `cell_list.{i}.conv.{weight,bias}` keys, forward(input_tensor, hidden_state)
-> (layer_output_list, last_state_list). Gate order i,f,o,g (convlstm.py:44).

The gate convolution conv([x;h]) is split into its x half, batched over all T steps in one
launch, and its h half, which is the only part that is sequential; the sum rides in the h-conv's
residual epilogue and the cell update is one pointwise kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._modbase import PackedModule, to_dev


class ConvLSTMCell(PackedModule):

    def __init__(self, input_dim, hidden_dim, kernel_size, bias):
        super(ConvLSTMCell, self).__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        self.conv = nn.Conv2d(in_channels=self.input_dim + self.hidden_dim,
                              out_channels=4 * self.hidden_dim,
                              kernel_size=self.kernel_size,
                              padding=self.padding,
                              bias=self.bias)

    def _pack(self):
        w = self.conv.weight
        return {"wx": w[:, :self.input_dim].contiguous(), "wh": w[:, self.input_dim:].contiguous()}

    def gates_x(self, x_seq):
        """x half of the gate conv for a whole [n,Cin,h,w] stack of steps (+ bias)."""
        assert self.kernel_size[0] == self.kernel_size[1], "square kernels only"
        pk = self.packed()
        return ops.conv2d(x_seq, pk["wx"], self.conv.bias, pad=self.padding[0])

    def step(self, gx_t, h_cur, c_inout, h_out):
        """gx_t [b,4Ch,h,w] (overwritten with the full gate pre-activations), c updated in place."""
        pk = self.packed()
        ops.conv2d(h_cur, pk["wh"], None, pad=self.padding[0], res_view=gx_t, out_view=gx_t)
        for b in range(gx_t.shape[0]):
            ops.convlstm_pointwise(gx_t[b], c_inout[b], h_out[b])

    def forward(self, input_tensor, cur_state):
        self._check_inference()
        dev = self._device()
        h_cur, c_cur = to_dev(cur_state[0], dev), to_dev(cur_state[1], dev)
        gx = self.gates_x(to_dev(input_tensor, dev))
        c_next = c_cur.clone()
        h_next = torch.empty_like(h_cur)
        self.step(gx, h_cur, c_next, h_next)
        return h_next, c_next

    def init_hidden(self, batch_size, image_size):
        height, width = image_size
        return (torch.zeros(batch_size, self.hidden_dim, height, width, device=self.conv.weight.device),
                torch.zeros(batch_size, self.hidden_dim, height, width, device=self.conv.weight.device))


class ConvLSTM(PackedModule):

    def __init__(self, input_dim, hidden_dim, kernel_size, num_layers,
                 batch_first=False, bias=True, return_all_layers=False):
        super(ConvLSTM, self).__init__()
        self._check_kernel_size_consistency(kernel_size)
        kernel_size = self._extend_for_multilayer(kernel_size, num_layers)
        hidden_dim = self._extend_for_multilayer(hidden_dim, num_layers)
        if not len(kernel_size) == len(hidden_dim) == num_layers:
            raise ValueError('Inconsistent list length.')
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.num_layers = num_layers
        self.batch_first = batch_first
        self.bias = bias
        self.return_all_layers = return_all_layers
        cell_list = []
        for i in range(0, self.num_layers):
            cur_input_dim = self.input_dim if i == 0 else self.hidden_dim[i - 1]
            cell_list.append(ConvLSTMCell(input_dim=cur_input_dim, hidden_dim=self.hidden_dim[i],
                                          kernel_size=self.kernel_size[i], bias=self.bias))
        self.cell_list = nn.ModuleList(cell_list)

    def forward(self, input_tensor, hidden_state=None):
        """input (t,b,c,h,w) or (b,t,c,h,w) -> (layer_output_list, last_state_list), as
        convlstm.py:120-176 (stateful: hidden_state = [[h,c], ...] per layer)."""
        self._check_inference()
        dev = self._device()
        x = to_dev(input_tensor, dev)
        if not self.batch_first:
            x = x.permute(1, 0, 2, 3, 4).contiguous()
        b, T, _, h, w = x.shape
        if hidden_state is None:
            hidden_state = self._init_hidden(batch_size=b, image_size=(h, w))
        layer_output_list, last_state_list = [], []
        cur = x
        for li, cell in enumerate(self.cell_list):
            Ch = self.hidden_dim[li]
            h_cur = to_dev(hidden_state[li][0], dev)
            c_run = to_dev(hidden_state[li][1], dev).clone()          # the caller's state is not mutated
            gx_all = cell.gates_x(cur.reshape(b * T, cur.shape[2], h, w)).view(b, T, 4 * Ch, h, w)
            out = torch.empty((b, T, Ch, h, w), dtype=torch.float32, device=dev)
            for t in range(T):
                cell.step(gx_all[:, t], h_cur, c_run, out[:, t])
                h_cur = out[:, t]
            cur = out
            layer_output_list.append(out)
            last_state_list.append([h_cur.contiguous() if b > 1 else h_cur.reshape(b, Ch, h, w), c_run])
        if not self.return_all_layers:
            layer_output_list = layer_output_list[-1:]
            last_state_list = last_state_list[-1:]
        return layer_output_list, last_state_list

    def _init_hidden(self, batch_size, image_size):
        return [self.cell_list[i].init_hidden(batch_size, image_size) for i in range(self.num_layers)]

    @staticmethod
    def _check_kernel_size_consistency(kernel_size):
        if not (isinstance(kernel_size, tuple) or
                (isinstance(kernel_size, list) and all([isinstance(elem, tuple) for elem in kernel_size]))):
            raise ValueError('`kernel_size` must be tuple or list of tuples')

    @staticmethod
    def _extend_for_multilayer(param, num_layers):
        if not isinstance(param, list):
            param = [param] * num_layers
        return param
