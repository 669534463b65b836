"""3-D non-local attention modules with the reference's signatures and state_dict keys (reference
Module/Attention.py:8-185).  Stand-alone: the reference imports them into Generator.py but never wires
them into Generator.forward."""
import torch
import torch.nn as nn
from torch.nn import functional as F
from torch.nn import init

from .. import ops


def _conv1(x, m):
    return ops.conv(x, m.weight, m.bias)


class SeparableAttnCell(nn.Module):

    def __init__(self, in_dim, attn_id=None, activation=F.relu, pooling_factor=2, padding_mode='constant',
                 padding_value=0):
        super().__init__()
        self.attn_id = attn_id
        self.activation = activation
        self.query_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim // 2, kernel_size=1)
        self.key_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim // 2, kernel_size=1)
        self.value_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.pooling = nn.MaxPool3d(kernel_size=(2, 1, 1), stride=(pooling_factor, 1, 1))
        self.pooling_factor = pooling_factor
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        self.gamma = nn.Parameter(torch.zeros((1,)))
        self.softmax = nn.Softmax(dim=-1)

    def init_conv(self, conv, glu=True):
        init.xavier_uniform_(conv.weight)
        if conv.bias is not None:
            conv.bias.data.zero_()

    def forward(self, x):
        batch_size, C, T, W, H = x.size()
        assert T % 2 == 0 and W % 2 == 0 and H % 2 == 0, "T, W, H is not even"
        if self.pooling_factor != 2:
            raise NotImplementedError("pooling_factor must be 2 (window 2, stride 2 along the attended axis)")
        xs, xr = ops.fork(x, 2)
        if self.attn_id == 'T':
            attn_dim, out = T, xs
        elif self.attn_id == 'W':
            attn_dim, out = W, ops.Permute5Fn.apply(xs, (0, 1, 3, 2, 4))
        else:
            attn_dim, out = H, ops.Permute5Fn.apply(xs, (0, 1, 4, 3, 2))
        oq, ok, ov = ops.fork(out, 3)
        half = attn_dim // 2
        # the reference reinterprets memory with raw .view()s: q as (B, A, L), k/v as (B, L', A/2)
        query = _conv1(oq, self.query_conv).view(batch_size, attn_dim, -1)
        key = ops.MaxPoolFn.apply(_conv1(ok, self.key_conv), 2, 1, 1).view(batch_size, -1, half)
        value = ops.MaxPoolFn.apply(_conv1(ov, self.value_conv), 2, 1, 1).view(batch_size, -1, half)
        out = ops.AttnCoreFn.apply(query, key, value, True)            # (B, CWH.., A)
        if self.attn_id == 'T':
            out = ops.Permute5Fn.apply(out.view(batch_size, C, W, H, T), (0, 1, 4, 2, 3))
        elif self.attn_id == 'W':
            out = ops.Permute5Fn.apply(out.view(batch_size, C, T, H, W), (0, 1, 2, 4, 3))
        else:
            out = out.view(batch_size, C, T, W, H)
        return ops.ScaleResidualFn.apply(out, xr, self.gamma)


class SeparableAttn(nn.Module):

    def __init__(self, in_dim, activation=F.relu, pooling_factor=2, padding_mode='constant', padding_value=0):
        super().__init__()
        self.model = nn.Sequential(
            SeparableAttnCell(in_dim, 'T', activation, pooling_factor, padding_mode, padding_value),
            SeparableAttnCell(in_dim, 'W', activation, pooling_factor, padding_mode, padding_value),
            SeparableAttnCell(in_dim, 'H', activation, pooling_factor, padding_mode, padding_value))

    def forward(self, x):
        return self.model(x)


class SelfAttention(nn.Module):

    def __init__(self, in_dim, activation=F.relu, pooling_factor=2):
        super().__init__()
        self.activation = activation
        self.query_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim // 2, kernel_size=1)
        self.key_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim // 2, kernel_size=1)
        self.value_conv = nn.Conv3d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.pooling = nn.MaxPool3d(kernel_size=2, stride=pooling_factor)
        self.pooling_factor = pooling_factor ** 3
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def init_conv(self, conv, glu=True):
        init.xavier_uniform_(conv.weight)
        if conv.bias is not None:
            conv.bias.data.zero_()

    def forward(self, x):
        if len(x.size()) == 4:
            batch_size, C, W, H = x.size()
            T = 1
        else:
            batch_size, C, T, W, H = x.size()
        assert T % 2 == 0 and W % 2 == 0 and H % 2 == 0, "T, W, H is not even"
        if self.pooling_factor != 8:
            raise NotImplementedError("pooling_factor must be 2 (2x2x2 max-pool, stride 2)")
        N = T * W * H
        xq, xk, xv, xr = ops.fork(x, 4)
        query = _conv1(xq, self.query_conv).view(batch_size, -1, N)
        key = ops.MaxPoolFn.apply(_conv1(xk, self.key_conv), 2, 2, 2).view(batch_size, -1, N // 8)
        value = ops.MaxPoolFn.apply(_conv1(xv, self.value_conv), 2, 2, 2).view(batch_size, -1, N // 8)
        out = ops.AttnCoreFn.apply(query, key, value, False).view(batch_size, C, T, W, H)
        return ops.ScaleResidualFn.apply(out, xr, self.gamma)
