"""Spatial / Temporal discriminators with the reference's signatures and state_dict keys (reference
Module/Discriminators.py:82-447).  Only the configurations the reference instantiates are accelerated:
GBlock / Res3dBlock with bn=False (the bn=True branch needs the reference's 148-wide HyperBN and is never
taken, Discriminators.py:229-238,388-397)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from .. import ops
from .Normalization import SpectralNorm


def init_conv(conv, glu=True):
    init.xavier_uniform_(conv.weight)
    if conv.bias is not None:
        conv.bias.data.zero_()


class SelfAttention(nn.Module):
    """2-D non-local block (reference :82-119): plain (not SN) 1x1 q/k/v, softmax(Q^T K) with no 1/sqrt(d),
    gamma * out + x."""

    def __init__(self, in_dim, activation=F.relu):
        super().__init__()
        self.chanel_in = in_dim
        self.activation = activation
        self.query_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)
        init_conv(self.query_conv)
        init_conv(self.key_conv)
        init_conv(self.value_conv)

    def forward(self, x):
        B, C, W, H = x.shape
        xq, xk, xv, xr = ops.fork(x, 4)
        q = ops.conv(xq, self.query_conv.weight, self.query_conv.bias).view(B, -1, W * H)
        k = ops.conv(xk, self.key_conv.weight, self.key_conv.bias).view(B, -1, W * H)
        v = ops.conv(xv, self.value_conv.weight, self.value_conv.bias).view(B, -1, W * H)
        out = ops.AttnCoreFn.apply(q, k, v, False).view(B, C, W, H)
        return ops.ScaleResidualFn.apply(out, xr, self.gamma)


class _ResDown(nn.Module):
    """Shared body of GBlock (2-D) and Res3dBlock (3-D), reference :151-211 / :305-366."""
    _conv = None
    _pool = (1, 2, 2)

    def __init__(self, in_channel, out_channel, kernel_size, padding=1, stride=1, n_class=None, bn=True,
                 activation=F.relu, upsample=True, downsample=False):
        super().__init__()
        if activation is not F.relu:
            raise NotImplementedError("the fused CUDA block implements ReLU only")
        self.conv0 = SpectralNorm(self._conv(in_channel, out_channel, kernel_size, stride, padding, bias=True))
        self.conv1 = SpectralNorm(self._conv(out_channel, out_channel, kernel_size, stride, padding, bias=True))
        self.skip_proj = False
        if in_channel != out_channel or upsample or downsample:
            self.conv_sc = SpectralNorm(self._conv(in_channel, out_channel, 1, 1, 0))
            self.skip_proj = True
        self.upsample = upsample
        self.downsample = downsample
        self.activation = activation
        self.bn = bn
        if bn:
            raise NotImplementedError("bn=True needs the reference's 148-wide HyperBN; the reference never "
                                      "instantiates it (Discriminators.py:229-238)")
        if upsample and self._conv is nn.Conv3d:
            raise NotImplementedError("3-D upsampling is not used by the reference")

    def forward(self, input, condition=None):
        up = 1 if self.upsample else 0
        if self.skip_proj:
            xa, xb = ops.fork(input, 2)
            skip = self.conv_sc.conv(xb)               # 1x1: commutes with the nearest upsample
        else:
            xa, skip = ops.fork(input, 2)
        out = self.conv0.conv(xa, in_relu=1, in_up=up)
        out = self.conv1.conv(out, res=skip, in_relu=1, res_up=up if self.skip_proj else 0)
        if self.downsample:
            out = ops.AvgPoolFn.apply(out, *self._pool)
        return out


class GBlock(_ResDown):
    _conv = nn.Conv2d
    _pool = (1, 2, 2)

    def __init__(self, in_channel, out_channel, kernel_size=[3, 3], padding=1, stride=1, n_class=None, bn=True,
                 activation=F.relu, upsample=True, downsample=False):
        super().__init__(in_channel, out_channel, kernel_size, padding, stride, n_class, bn, activation, upsample,
                         downsample)


class Res3dBlock(_ResDown):
    _conv = nn.Conv3d
    _pool = (2, 2, 2)

    def __init__(self, in_channel, out_channel, kernel_size=[3, 3, 3], padding=1, stride=1, n_class=None, bn=True,
                 activation=F.relu, upsample=True, downsample=False):
        super().__init__(in_channel, out_channel, kernel_size, padding, stride, n_class, bn, activation, upsample,
                         downsample)


def _pre(mod, x, pool):
    """pre_conv (SN-conv, ReLU, SN-conv, AvgPool) + pre_skip (1x1 on the pooled input), :248-249 / :407-408."""
    xa, xb = ops.fork(x, 2)
    out = mod.pre_conv[0].conv(xa)
    out = mod.pre_conv[2].conv(out, in_relu=1)
    out = ops.AvgPoolFn.apply(out, *pool)
    xs = ops.AvgPoolFn.apply(xb, *pool)
    return mod.pre_skip.conv(xs, res=out)


def _head(mod, out, class_id, T):
    m, e = mod.linear.module, mod.embed.module
    return ops.DHeadFn.apply(out, m.weight_bar, m.bias, m.weight_u.data, m.weight_v.data, e.weight_bar,
                             e.weight_u.data, e.weight_v.data, class_id, T)


class SpatialDiscriminator(nn.Module):

    def __init__(self, chn=128, n_class=4):
        super().__init__()
        self.pre_conv = nn.Sequential(SpectralNorm(nn.Conv2d(3, 2 * chn, 3, padding=1)),
                                      nn.ReLU(),
                                      SpectralNorm(nn.Conv2d(2 * chn, 2 * chn, 3, padding=1)),
                                      nn.AvgPool2d(2))
        self.pre_skip = SpectralNorm(nn.Conv2d(3, 2 * chn, 1))
        self.conv1 = GBlock(2 * chn, 4 * chn, bn=False, upsample=False, downsample=True)
        self.attn = SelfAttention(4 * chn)
        self.conv2 = nn.Sequential(
            GBlock(4 * chn, 8 * chn, bn=False, upsample=False, downsample=True),
            GBlock(8 * chn, 16 * chn, bn=False, upsample=False, downsample=True),
            GBlock(16 * chn, 16 * chn, bn=False, upsample=False, downsample=True))
        self.linear = SpectralNorm(nn.Linear(16 * chn, 1))
        self.embed = nn.Embedding(n_class, 16 * chn)
        self.embed.weight.data.uniform_(-0.1, 0.1)
        self.embed = SpectralNorm(self.embed)

    def forward(self, x, class_id, taps=None):
        """x (B,T,3,H,W), class_id (B,) -> per-frame scores (B*T,) (no sum over T)."""
        batch_size, T, C, W, H = x.size()
        x = x.reshape(batch_size * T, C, H, W)
        out = _pre(self, x, (1, 2, 2))
        out = self.conv1(out)
        if taps is not None:
            taps["conv1"] = out
        out = self.attn(out)
        if taps is not None:
            taps["attn"] = out
        out = self.conv2(out)
        if taps is not None:
            taps["conv2"] = out
        return _head(self, out, class_id, T)


class TemporalDiscriminator(nn.Module):

    def __init__(self, chn=128, n_class=4):
        super().__init__()
        self.pre_conv = nn.Sequential(SpectralNorm(nn.Conv3d(3, 2 * chn, 3, padding=1)),
                                      nn.ReLU(),
                                      SpectralNorm(nn.Conv3d(2 * chn, 2 * chn, 3, padding=1)),
                                      nn.AvgPool3d(2))
        self.pre_skip = SpectralNorm(nn.Conv3d(3, 2 * chn, 1))
        self.res3d = Res3dBlock(2 * chn, 4 * chn, bn=False, upsample=False, downsample=True)
        self.self_attn = SelfAttention(4 * chn)
        self.conv = nn.Sequential(
            GBlock(4 * chn, 8 * chn, bn=False, upsample=False, downsample=True),
            GBlock(8 * chn, 16 * chn, bn=False, upsample=False, downsample=True),
            GBlock(16 * chn, 16 * chn, bn=False, upsample=False, downsample=True))
        self.linear = SpectralNorm(nn.Linear(16 * chn, 1))
        self.embed = nn.Embedding(n_class, 16 * chn)
        self.embed.weight.data.uniform_(-0.1, 0.1)
        self.embed = SpectralNorm(self.embed)

    def forward(self, x, class_id, taps=None):
        """x (B,3,T,H,W) (already phi-downsampled), class_id (B,) -> (B*(T//4),)."""
        out = _pre(self, x, (2, 2, 2))
        out = self.res3d(out)
        if taps is not None:
            taps["res3d"] = out
        out = ops.Permute5Fn.apply(out, (0, 2, 1, 3, 4))      # per-frame 2-D from here on (Q8)
        B, T, C, W, H = out.size()
        out = out.view(B * T, C, W, H)
        out = self.self_attn(out)
        if taps is not None:
            taps["attn"] = out
        out = self.conv(out)
        return _head(self, out, class_id, T)
