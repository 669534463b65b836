"""GResBlock with the reference's signature and state_dict keys (reference Module/GResBlock.py:9-86)."""
import torch.nn as nn
from torch.nn import functional as F

from .. import ops
from .Normalization import ConditionalNorm, SpectralNorm


class GResBlock(nn.Module):

    def __init__(self, in_channel, out_channel, kernel_size=None, padding=1, stride=1, n_class=96, bn=True,
                 activation=F.relu, upsample_factor=2, downsample_factor=1):
        super().__init__()
        if activation is not F.relu:
            raise NotImplementedError("the fused CUDA block implements ReLU only")
        self.upsample_factor = upsample_factor if downsample_factor == 1 else 1
        self.downsample_factor = downsample_factor
        self.activation = activation
        self.bn = bn if downsample_factor == 1 else False
        if kernel_size is None:
            kernel_size = [3, 3]
        self.conv0 = SpectralNorm(nn.Conv2d(in_channel, out_channel, kernel_size, stride, padding, bias=True))
        self.conv1 = SpectralNorm(nn.Conv2d(out_channel, out_channel, kernel_size, stride, padding, bias=True))
        self.skip_proj = True
        self.conv_sc = SpectralNorm(nn.Conv2d(in_channel, out_channel, 1, 1, 0))
        if bn:
            self.CBNorm1 = ConditionalNorm(in_channel, n_class)
            self.CBNorm2 = ConditionalNorm(out_channel, n_class)

    def forward(self, x, condition=None):
        """x (BT,C,W,H); condition (BT,n_class) -- or (R,n_class) with R | BT, image n using row n % R, which is
        what Generator.py:109-110's ``condition.repeat(T,1)`` amounts to."""
        if self.upsample_factor not in (1, 2):
            raise NotImplementedError("upsample_factor must be 1 or 2")
        up = 1 if self.upsample_factor == 2 else 0
        xa, xb = ops.fork(x, 2)
        # conv_sc is 1x1, so it commutes with the nearest upsample (Q14): run it at low resolution and
        # add its upsampled output in conv1's epilogue.
        skip = self.conv_sc.conv(xb)
        if self.bn:
            # each (CBN -> ReLU -> [up] -> conv) pair is one autograd node that keeps only its pre-norm input
            out = self.CBNorm1.fused_conv(xa, condition, self.conv0, up=up)
            out = self.CBNorm2.fused_conv(out, condition, self.conv1, res=skip, res_up=up)
        else:
            out = self.conv0.conv(xa, in_relu=1, in_up=up)
            out = self.conv1.conv(out, res=skip, in_relu=1, res_up=up)
        if self.downsample_factor != 1:
            d = self.downsample_factor
            out = ops.AvgPoolFn.apply(out, 1, d, d)
        return out
