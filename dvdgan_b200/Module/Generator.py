"""Generator with the reference's constructor/forward signature and state_dict keys (reference
Module/Generator.py:13-120): Embedding + Linear -> 4 x [3-layer ConvGRU over T -> 2 GResBlocks] -> ReLU ->
SN-conv3x3 -> tanh."""
import torch
import torch.nn as nn

from .. import ops
from .Attention import SelfAttention, SeparableAttn
from .ConvGRU import ConvGRU
from .GResBlock import GResBlock
from .Normalization import SpectralNorm


class Generator(nn.Module):

    def __init__(self, in_dim=120, latent_dim=4, n_class=4, ch=32, n_frames=48, hierar_flag=False, attention=False):
        """``attention=True`` (not a reference argument, default off) wires in the two non-local blocks the reference
        imports but leaves commented out (Generator.py:28-36): the pooled 3-D ``SelfAttention(8*ch)`` after the first
        ConvGRU stage and ``SeparableAttn(4*ch)`` in front of the last GResBlock.  Their parameters are extra
        state_dict keys (``self_attn.*``, ``sep_attn.*``); with the default the key set is the reference's."""
        super().__init__()
        self.in_dim = in_dim
        self.latent_dim = latent_dim
        self.n_class = n_class
        self.ch = ch
        self.hierar_flag = hierar_flag
        self.n_frames = n_frames

        self.embedding = nn.Embedding(n_class, in_dim)
        self.affine_transfrom = nn.Linear(in_dim * 2, latent_dim * latent_dim * 8 * ch)

        def gru(c, hs, ks):
            return ConvGRU(c, hidden_sizes=hs, kernel_sizes=ks, n_layers=3)

        def res(ci, co, **kw):
            return GResBlock(ci, co, n_class=in_dim * 2, **kw)

        self.conv = nn.ModuleList([
            gru(8 * ch, [8 * ch, 16 * ch, 8 * ch], [3, 5, 3]),
            res(8 * ch, 8 * ch, upsample_factor=1),
            res(8 * ch, 8 * ch),
            gru(8 * ch, [8 * ch, 16 * ch, 8 * ch], [3, 5, 3]),
            res(8 * ch, 8 * ch, upsample_factor=1),
            res(8 * ch, 8 * ch),
            gru(8 * ch, [8 * ch, 16 * ch, 8 * ch], [3, 5, 3]),
            res(8 * ch, 8 * ch, upsample_factor=1),
            res(8 * ch, 4 * ch),
            gru(4 * ch, [4 * ch, 8 * ch, 4 * ch], [3, 5, 5]),
            res(4 * ch, 4 * ch, upsample_factor=1),
            res(4 * ch, 2 * ch),
        ])
        self.colorize = SpectralNorm(nn.Conv2d(2 * ch, 3, kernel_size=(3, 3), padding=1))
        self.attention = bool(attention)
        if self.attention:
            self.self_attn = SelfAttention(8 * ch)
            self.sep_attn = SeparableAttn(4 * ch)

    def _attend(self, attn, y, B, T):
        """y (B*T, C, W, H) with b-major rows -> the 3-D block over (B, C, T, W, H) -> back."""
        _, C, W, H = y.shape
        v = ops.Permute5Fn.apply(y.view(B, T, C, W, H), (0, 2, 1, 3, 4))
        v = attn(v)
        return ops.Permute5Fn.apply(v, (0, 2, 1, 3, 4)).view(B * T, C, W, H)

    def forward(self, x, class_id, taps=None):
        """x: z (B,in_dim) float32; class_id (B,) int64 -> (B, n_frames, 3, 16*latent_dim, 16*latent_dim)."""
        if self.hierar_flag is True:
            raise NotImplementedError("hierar_flag=True is broken in the reference (tuple cat, SURVEY Q15)")
        B, T = x.shape[0], self.n_frames
        class_emb = ops.EmbeddingFn.apply(class_id, self.embedding.weight)
        cond = torch.cat((x, class_emb), dim=1)        # (B, 2*in_dim): tiny host-side glue
        y = ops.conv(cond, self.affine_transfrom.weight, self.affine_transfrom.bias)
        y = y.view(B, 8 * self.ch, self.latent_dim, self.latent_dim)
        for k, conv in enumerate(self.conv):
            if isinstance(conv, ConvGRU):
                if k == 0:
                    y = conv.forward_sequence(y, T_bcast=T)             # same input every frame (Q13)
                else:
                    _, C, W, H = y.shape
                    y = conv.forward_sequence(y.view(B, T, C, W, H))
                _, _, C, W, H = y.shape
                y = y.view(B * T, C, W, H)                              # b-major rows: b*T + t
                if self.attention and k == 0:
                    y = self._attend(self.self_attn, y, B, T)
            else:
                if self.attention and k == len(self.conv) - 1:
                    y = self._attend(self.sep_attn, y, B, T)
                # the reference conditions row i on sample i % B (condition.repeat(T,1), Q1)
                y = conv(y, cond)
            if taps is not None:
                taps[f"stage{k}"] = y
        if taps is not None:
            y = self.colorize.conv(y, in_relu=1)
            taps["pre_tanh"] = y
            y = ops.ActFn.apply(y, 2)
        else:
            y = self.colorize.conv(y, in_relu=1, out_act=2)
        BT, C, W, H = y.shape
        return y.view(B, T, C, W, H)
