"""SpectralNorm / ConditionalNorm with the reference's signatures and state_dict keys
(reference Module/Normalization.py:10-88), running on the dvdgan_b200 CUDA kernels."""
import torch
from torch import nn
from torch.nn import Parameter

from .. import ops


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


def _conv_is_supported(m):
    k = m.kernel_size
    return (all(s == 1 for s in m.stride) and all(d == 1 for d in m.dilation) and m.groups == 1
            and all(kk % 2 == 1 for kk in k) and tuple(m.padding) == tuple(kk // 2 for kk in k)
            and m.padding_mode == "zeros")


class SpectralNorm(nn.Module):
    """Wraps ``module`` so that ``module.<name>`` = ``<name>_bar / sigma`` with sigma from ``power_iterations``
    power-iteration steps per forward (state advances on EVERY forward, train or eval; sigma is differentiable;
    inner-module hooks never fire because the inner ``forward`` is bypassed -- reference :19-31, :62-64).
    Parameters registered on the inner module: ``<name>_u``, ``<name>_v`` (requires_grad=False), ``<name>_bar``."""

    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module = module
        self.name = name
        self.power_iterations = power_iterations
        if not self._made_params():
            self._make_params()

    def _made_params(self):
        return all(hasattr(self.module, self.name + s) for s in ("_u", "_v", "_bar"))

    def _make_params(self):
        w = getattr(self.module, self.name)
        height = w.data.shape[0]
        width = w.view(height, -1).data.shape[1]
        u = Parameter(w.data.new(height).normal_(0, 1), requires_grad=False)
        v = Parameter(w.data.new(width).normal_(0, 1), requires_grad=False)
        u.data = l2normalize(u.data)
        v.data = l2normalize(v.data)
        w_bar = Parameter(w.data)
        del self.module._parameters[self.name]
        self.module.register_parameter(self.name + "_u", u)
        self.module.register_parameter(self.name + "_v", v)
        self.module.register_parameter(self.name + "_bar", w_bar)

    # -- fused entry points used by the blocks ------------------------------------------------
    def _wuv(self):
        m = self.module
        return getattr(m, self.name + "_bar"), getattr(m, self.name + "_u"), getattr(m, self.name + "_v")

    def _extra_iterations(self):
        w, u, v = self._wuv()
        for _ in range(self.power_iterations - 1):
            ops.sn_sigma(w.detach(), u.data, v.data)

    def conv(self, x, res=None, in_relu=0, in_up=0, out_act=0, res_up=0):
        """SN conv / linear with fused input ReLU, upsample, residual add and output activation."""
        m = self.module
        if isinstance(m, (nn.Conv2d, nn.Conv3d)) and not _conv_is_supported(m):
            raise NotImplementedError("dvdgan_b200 convolutions are stride-1, same-padded, odd-kernel, groups=1")
        w, u, v = self._wuv()
        self._extra_iterations()
        return ops.conv(x, w, m.bias, u.data, v.data, res, in_relu, in_up, out_act, res_up)

    def forward(self, *args):
        m = self.module
        if self.name == "weight" and isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)) and len(args) == 1:
            return self.conv(args[0])
        w, u, v = self._wuv()
        self._extra_iterations()
        w_sn = ops.SNWeightFn.apply(w, u.data, v.data)
        if self.name == "weight" and isinstance(m, nn.Embedding) and len(args) == 1 and m.padding_idx is None \
                and m.max_norm is None:
            idx = args[0]
            return ops.EmbeddingFn.apply(idx.reshape(-1), w_sn).view(*idx.shape, w_sn.shape[1])
        # generic inner module: its own forward (a PyTorch op) consumes the normalised weight
        setattr(m, self.name, w_sn)
        return m.forward(*args)


class ConditionalNorm(nn.Module):
    """BatchNorm2d(affine=False) followed by a per-row affine from ``Linear(n_condition -> 2C)``
    (reference :66-88; note the column-sliced init)."""

    def __init__(self, in_channel, n_condition=96):
        super().__init__()
        self.in_channel = in_channel
        self.bn = nn.BatchNorm2d(self.in_channel, affine=False)
        self.embed = nn.Linear(n_condition, self.in_channel * 2)
        self.embed.weight.data[:, :self.in_channel].normal_(1, 0.02)
        self.embed.weight.data[:, self.in_channel:].zero_()

    def fused(self, x, class_id, relu=False, up=0):
        """class_id: (R, n_condition) with R == N, or R dividing N (row of image n is n % R)."""
        gb = ops.conv(class_id, self.embed.weight, self.embed.bias)
        bn = self.bn
        use_batch = self.training or bn.running_mean is None
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return ops.CBNFn.apply(x, gb, bn.running_mean, bn.running_var, bn.num_batches_tracked, relu, up,
                               use_batch, momentum, bn.eps)

    def fused_conv(self, x, class_id, sn_conv, up=0, res=None, res_up=0):
        """ConditionalNorm -> ReLU -> [nearest x2] -> ``sn_conv`` (a SpectralNorm-wrapped conv) [+ residual] without
        keeping the normalised activation for the backward (ops.CBNConvFn)."""
        gb = ops.conv(class_id, self.embed.weight, self.embed.bias)
        bn = self.bn
        use_batch = self.training or bn.running_mean is None
        momentum = 0.1 if bn.momentum is None else bn.momentum
        m = sn_conv.module
        w, u, v = sn_conv._wuv()
        sn_conv._extra_iterations()
        return ops.CBNConvFn.apply(x, gb, bn.running_mean, bn.running_var, bn.num_batches_tracked, up, use_batch,
                                   momentum, bn.eps, w, m.bias, u.data, v.data, res, res_up)

    def forward(self, x, class_id):
        return self.fused(x, class_id)
