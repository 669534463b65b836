"""ConvGRUCell / ConvGRU with the reference's signatures and state_dict keys (reference
Module/ConvGRU.py:5-133).  A whole clip runs through ``ConvGRU.forward_sequence`` (one library call per
layer: batched x-half implicit GEMM + the sequential h-half loop on the device); the per-step ``forward``
of the reference is kept for drop-in use.  Stacks whose layers all keep their full state run as a wavefront: layer l
on stream l, a chunk of frames behind layer l-1 (``ops.GRUStackFn``)."""
import torch
import torch.nn as nn
from torch.nn import init

from .. import ops


class ConvGRUCell(nn.Module):

    def __init__(self, input_size, hidden_size, kernel_size, activation=torch.sigmoid):
        super().__init__()
        if activation is not torch.sigmoid:
            raise NotImplementedError("the CUDA ConvGRU cell implements the sigmoid gate activation only")
        padding = kernel_size // 2
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.reset_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.update_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.out_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.activation = activation
        init.orthogonal_(self.reset_gate.weight)
        init.orthogonal_(self.update_gate.weight)
        init.orthogonal_(self.out_gate.weight)
        init.constant_(self.reset_gate.bias, 0.)
        init.constant_(self.update_gate.bias, 0.)
        init.constant_(self.out_gate.bias, 0.)

    def sequence(self, x, h0=None, T_bcast=0):
        """x (B,T,Cx,H,W) [or (B,Cx,H,W) repeated T_bcast times] -> h (B,T,Ch,H,W)."""
        return ops.GRULayerFn.apply(x, h0, self.update_gate.weight, self.reset_gate.weight, self.out_gate.weight,
                                    self.update_gate.bias, self.reset_gate.bias, self.out_gate.bias, T_bcast)

    def forward(self, x, prev_state=None):
        return self.sequence(x.unsqueeze(1), prev_state)[:, 0]


class ConvGRU(nn.Module):

    def __init__(self, input_size, hidden_sizes, kernel_sizes, n_layers):
        super().__init__()
        self.input_size = input_size
        if type(hidden_sizes) != list:
            self.hidden_sizes = [hidden_sizes] * n_layers
        else:
            assert len(hidden_sizes) == n_layers, '`hidden_sizes` must have the same length as n_layers'
            self.hidden_sizes = hidden_sizes
        if type(kernel_sizes) != list:
            self.kernel_sizes = [kernel_sizes] * n_layers
        else:
            assert len(kernel_sizes) == n_layers, '`kernel_sizes` must have the same length as n_layers'
            self.kernel_sizes = kernel_sizes
        self.n_layers = n_layers
        cells = nn.ModuleList()
        for i in range(self.n_layers):
            input_dim = self.input_size if i == 0 else self.hidden_sizes[i - 1]
            cells.append(ConvGRUCell(input_dim, self.hidden_sizes[i], self.kernel_sizes[i]))
        self.cells = cells

    def forward(self, x, hidden=None):
        """One time step through all layers; returns the list of new hidden states (reference :104-133)."""
        if hidden is None:
            hidden = [None] * self.n_layers
        input_ = x
        output = []
        for i in range(self.n_layers):
            upd = self.cells[i](input_, hidden[i])
            output.append(upd)
            input_ = upd
        return output

    def forward_sequence(self, x, T_bcast=0):
        """The Generator's frame loop (Generator.py:87-106) in one go: the last layer's hidden state for
        every frame, (B,T,C_last,H,W), from zero initial state."""
        if T_bcast:
            B, _, H, W = x.shape
            T = T_bcast
        else:
            B, T, _, H, W = x.shape
        sigs = [(c.input_size, c.hidden_size, H, W, c.update_gate.kernel_size[0]) for c in self.cells]
        chunk = ops.gru_wavefront_chunk(B, T, sigs)
        if chunk:          # the layers as a wavefront over chunks of frames, one stream per layer (ops.GRUStackFn)
            params = [p for c in self.cells for p in (c.update_gate.weight, c.reset_gate.weight, c.out_gate.weight,
                                                      c.update_gate.bias, c.reset_gate.bias, c.out_gate.bias)]
            return ops.GRUStackFn.apply(x, T_bcast, chunk, *params)
        h = self.cells[0].sequence(x, None, T_bcast)
        for i in range(1, self.n_layers):
            h = self.cells[i].sequence(h)
        return h
