"""CPU oracle for the DVD-GAN G + Ds + Dt training step.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*, never the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``dvdgan_b200/`` imports it.

It is a functional restatement (plain functions over a ``dict[str, Tensor]`` that has the
reference's ``state_dict`` keys) of the reference's PyTorch modules.  The arithmetic itself
lives in PyTorch (third-party, not vendored: ``Readme.md:9`` asks for "pytorch 1.12", this
image has 2.11.0) -- ``F.conv2d/conv3d/linear/batch_norm/avg_pool*/max_pool3d/interpolate``,
``torch.bmm/softmax/mv/dot`` -- exactly the calls the reference makes, in the same order, so
the restatement is bit-identical to the reference on CPU.

Parity pinning: the reference ships no tests and no golden vectors, so
this oracle is pinned against outputs of the reference itself, imported from
``/root/reference`` by ``tests/golden/make_golden.py`` (committed) which wrote the fixtures
under ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks every fixture bit-exactly.

All ``file:line`` citations are into the reference tree.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Normalization.py
# --------------------------------------------------------------------------------------


def l2normalize(v, eps=1e-12):
    """Module/Normalization.py:7-8."""
    return v / (v.norm() + eps)


def spectral_norm_weight(sd, prefix, power_iterations=1):
    """Module/Normalization.py:19-31 (``SpectralNorm._update_u_v``).

    ``prefix`` ends with ``'module.'``.  Mutates ``weight_u`` / ``weight_v`` in ``sd`` (the
    reference updates ``.data`` on every forward, train or eval) and returns ``W_bar / sigma``
    with autograd flowing through ``W_bar`` only (sigma is differentiable, Q3).
    """
    u = sd[prefix + "weight_u"]
    v = sd[prefix + "weight_v"]
    w = sd[prefix + "weight_bar"]
    height = w.shape[0]
    # ``.data =`` (not copy_) exactly as the reference does: it swaps the storage without bumping
    # the autograd version counter, so a graph built by an EARLIER forward of the same module
    # back-propagates d(sigma)/dW = u v^T with the LATEST u, v (quirk Q17).
    wd = w.data.view(height, -1)
    for _ in range(power_iterations):
        v.data = l2normalize(torch.mv(torch.t(wd), u.data))
        u.data = l2normalize(torch.mv(wd, v.data))
    sigma = u.dot(w.view(height, -1).mv(v))
    return w / sigma.expand_as(w)


def conditional_norm(sd, prefix, x, cond, training=True, momentum=0.1, eps=1e-5):
    """Module/Normalization.py:78-88 (``ConditionalNorm.forward``).

    BatchNorm2d(affine=False) in train mode updates ``running_mean`` / ``running_var`` /
    ``num_batches_tracked`` in ``sd``.  gamma/beta are per *row* of ``cond``.
    """
    C = x.shape[1]
    rm = sd[prefix + "bn.running_mean"]
    rv = sd[prefix + "bn.running_var"]
    if training:
        sd[prefix + "bn.num_batches_tracked"] += 1
    out = F.batch_norm(x, rm, rv, None, None, training, momentum, eps)
    embed = F.linear(cond, sd[prefix + "embed.weight"], sd[prefix + "embed.bias"])
    gamma, beta = embed.chunk(2, 1)
    gamma = gamma.view(-1, C, 1, 1)
    beta = beta.view(-1, C, 1, 1)
    return gamma * out + beta


# --------------------------------------------------------------------------------------
# ConvGRU.py
# --------------------------------------------------------------------------------------


def convgru_cell(sd, prefix, x, prev_state=None):
    """Module/ConvGRU.py:29-54 (``ConvGRUCell.forward``); padding = k // 2 (line 13)."""
    wu, bu = sd[prefix + "update_gate.weight"], sd[prefix + "update_gate.bias"]
    wr, br = sd[prefix + "reset_gate.weight"], sd[prefix + "reset_gate.bias"]
    wo, bo = sd[prefix + "out_gate.weight"], sd[prefix + "out_gate.bias"]
    pad = wu.shape[-1] // 2
    if prev_state is None:
        state_size = [x.shape[0], wu.shape[0]] + list(x.shape[2:])
        prev_state = torch.zeros(state_size, dtype=torch.float32)  # ConvGRU.py:41-44 (fp32; cat promotes)
    stacked = torch.cat([x, prev_state], dim=1)
    update = torch.sigmoid(F.conv2d(stacked, wu, bu, padding=pad))
    reset = torch.sigmoid(F.conv2d(stacked, wr, br, padding=pad))
    out_inputs = torch.tanh(F.conv2d(torch.cat([x, prev_state * reset], dim=1), wo, bo, padding=pad))
    return prev_state * (1 - update) + out_inputs * update


def convgru(sd, prefix, x, hidden=None, n_layers=3):
    """Module/ConvGRU.py:104-133 (``ConvGRU.forward``): returns the list of new hiddens."""
    if hidden is None:
        hidden = [None] * n_layers
    inp = x
    output = []
    for i in range(n_layers):
        h = convgru_cell(sd, f"{prefix}cells.{i}.", inp, hidden[i])
        output.append(h)
        inp = h
    return output


# --------------------------------------------------------------------------------------
# GResBlock.py
# --------------------------------------------------------------------------------------


def _sn_conv2d(sd, prefix, x, padding):
    w = spectral_norm_weight(sd, prefix + "module.")
    return F.conv2d(x, w, sd[prefix + "module.bias"], stride=1, padding=padding)


def gresblock(sd, prefix, x, condition, upsample_factor=2, training=True):
    """Module/GResBlock.py:42-86 with bn=True, downsample_factor=1 (the only use, Generator.py:41-54)."""
    BT, C, W, H = x.shape
    out = conditional_norm(sd, prefix + "CBNorm1.", x, condition, training)
    out = F.relu(out)
    if upsample_factor != 1:
        out = F.interpolate(out, scale_factor=upsample_factor)
    out = _sn_conv2d(sd, prefix + "conv0.", out, 1)
    out = out.view(BT, -1, W * upsample_factor, H * upsample_factor)
    out = conditional_norm(sd, prefix + "CBNorm2.", out, condition, training)
    out = F.relu(out)
    out = _sn_conv2d(sd, prefix + "conv1.", out, 1)
    skip = x
    if upsample_factor != 1:
        skip = F.interpolate(skip, scale_factor=upsample_factor)
    skip = _sn_conv2d(sd, prefix + "conv_sc.", skip, 0)
    y = out + skip
    return y.view(BT, -1, W * upsample_factor, H * upsample_factor)


# --------------------------------------------------------------------------------------
# Generator.py
# --------------------------------------------------------------------------------------

# Generator.py:38-55 -- (kind, upsample_factor); ConvGRU stages at k = 0, 3, 6, 9.
G_LAYOUT = [("gru", 0), ("res", 1), ("res", 2), ("gru", 0), ("res", 1), ("res", 2),
            ("gru", 0), ("res", 1), ("res", 2), ("gru", 0), ("res", 1), ("res", 2)]


def generator_forward(sd, z, class_id, n_frames, ch, latent_dim=4, training=True, taps=None):
    """Module/Generator.py:63-120 (hierar_flag=False).  ``taps`` (optional dict) receives
    intermediate tensors: 'stage{k}' after each entry of ``self.conv`` and 'pre_tanh'."""
    class_emb = F.embedding(class_id, sd["embedding.weight"])
    cond1 = torch.cat((z, class_emb), dim=1)
    y = F.linear(cond1, sd["affine_transfrom.weight"], sd["affine_transfrom.bias"])
    y = y.view(-1, 8 * ch, latent_dim, latent_dim)
    for k, (kind, up) in enumerate(G_LAYOUT):
        if kind == "gru":
            if k > 0:
                _, C, W, H = y.shape
                y = y.view(-1, n_frames, C, W, H).contiguous()
            frame_list = []
            for i in range(n_frames):
                if k == 0:
                    xi = y                                   # Q13: same input every frame
                else:
                    xi = y[:, i, :, :, :].squeeze(1)
                frame_list.append(convgru(sd, f"conv.{k}.", xi, frame_list[i - 1] if i > 0 else None))
            y = torch.cat([f[-1].unsqueeze(0) for f in frame_list], dim=0)  # T,B,C,W,H
            y = y.permute(1, 0, 2, 3, 4).contiguous()
            B, T, C, W, H = y.shape
            y = y.view(-1, C, W, H)
        else:
            condition = torch.cat([z, class_emb], dim=1).repeat(n_frames, 1)   # Q1: t-major rows
            y = gresblock(sd, f"conv.{k}.", y, condition, up, training)
        if taps is not None:
            taps[f"stage{k}"] = y
    y = F.relu(y)
    y = _sn_conv2d(sd, "colorize.", y, 1)
    if taps is not None:
        taps["pre_tanh"] = y
    y = torch.tanh(y)
    BT, C, W, H = y.shape
    return y.view(-1, n_frames, C, W, H)


# --------------------------------------------------------------------------------------
# Discriminators.py
# --------------------------------------------------------------------------------------


def d_self_attention(sd, prefix, x):
    """Module/Discriminators.py:100-119 (2-D ``SelfAttention``; plain convs, no 1/sqrt(d))."""
    B, C, W, H = x.shape
    q = F.conv2d(x, sd[prefix + "query_conv.weight"], sd[prefix + "query_conv.bias"])
    k = F.conv2d(x, sd[prefix + "key_conv.weight"], sd[prefix + "key_conv.bias"])
    v = F.conv2d(x, sd[prefix + "value_conv.weight"], sd[prefix + "value_conv.bias"])
    pq = q.view(B, -1, W * H).permute(0, 2, 1)
    pk = k.view(B, -1, W * H)
    energy = torch.bmm(pq, pk)
    attention = torch.softmax(energy, dim=-1)
    pv = v.view(B, -1, W * H)
    out = torch.bmm(pv, attention.permute(0, 2, 1)).view(B, C, W, H)
    return sd[prefix + "gamma"] * out + x


def gblock_down(sd, prefix, x):
    """Module/Discriminators.py:180-211 with bn=False, upsample=False, downsample=True."""
    out = F.relu(x)
    out = _sn_conv2d(sd, prefix + "conv0.", out, 1)
    out = F.relu(out)
    out = _sn_conv2d(sd, prefix + "conv1.", out, 1)
    out = F.avg_pool2d(out, 2)
    skip = _sn_conv2d(sd, prefix + "conv_sc.", x, 0)
    skip = F.avg_pool2d(skip, 2)
    return out + skip


def _sn_conv3d(sd, prefix, x, padding):
    w = spectral_norm_weight(sd, prefix + "module.")
    return F.conv3d(x, w, sd[prefix + "module.bias"], stride=1, padding=padding)


def res3dblock_down(sd, prefix, x):
    """Module/Discriminators.py:335-366 with bn=False, upsample=False, downsample=True."""
    out = F.relu(x)
    out = _sn_conv3d(sd, prefix + "conv0.", out, 1)
    out = F.relu(out)
    out = _sn_conv3d(sd, prefix + "conv1.", out, 1)
    out = F.avg_pool3d(out, 2)
    skip = _sn_conv3d(sd, prefix + "conv_sc.", x, 0)
    skip = F.avg_pool3d(skip, 2)
    return out + skip


def _d_head(sd, out, class_id, T):
    """Discriminators.py:264-291 / 421-447: ReLU, sum over HxW, SN-Linear + SN-Embedding projection."""
    out = F.relu(out)
    out = out.view(out.size(0), out.size(1), -1).sum(2)
    w_lin = spectral_norm_weight(sd, "linear.module.")
    out_linear = F.linear(out, w_lin, sd["linear.module.bias"]).squeeze(1)
    class_id = class_id.view(-1, 1).repeat(1, T).view(-1)
    w_emb = spectral_norm_weight(sd, "embed.module.")
    embed = F.embedding(class_id, w_emb)
    prod = (out * embed).sum(1)
    return out_linear + prod


def spatial_discriminator(sd, x, class_id, taps=None):
    """Module/Discriminators.py:242-291.  x (B,T,3,H,W) -> (B*T,) per-frame scores (Q7)."""
    B, T, C, W, H = x.shape
    x = x.view(B * T, C, H, W)
    out = _sn_conv2d(sd, "pre_conv.0.", x, 1)
    out = F.relu(out)
    out = _sn_conv2d(sd, "pre_conv.2.", out, 1)
    out = F.avg_pool2d(out, 2)
    out = out + _sn_conv2d(sd, "pre_skip.", F.avg_pool2d(x, 2), 0)
    out = gblock_down(sd, "conv1.", out)
    if taps is not None:
        taps["conv1"] = out
    out = d_self_attention(sd, "attn.", out)
    if taps is not None:
        taps["attn"] = out
    for i in range(3):
        out = gblock_down(sd, f"conv2.{i}.", out)
    if taps is not None:
        taps["conv2"] = out
    return _d_head(sd, out, class_id, T)


def temporal_discriminator(sd, x, class_id, taps=None):
    """Module/Discriminators.py:400-447.  x (B,3,T,H,W) -> (B*(T//4),)."""
    out = _sn_conv3d(sd, "pre_conv.0.", x, 1)
    out = F.relu(out)
    out = _sn_conv3d(sd, "pre_conv.2.", out, 1)
    out = F.avg_pool3d(out, 2)
    out = out + _sn_conv3d(sd, "pre_skip.", F.avg_pool3d(x, 2), 0)
    out = res3dblock_down(sd, "res3d.", out)
    if taps is not None:
        taps["res3d"] = out
    out = out.permute(0, 2, 1, 3, 4).contiguous()
    B, T, C, W, H = out.shape
    out = out.view(B * T, C, W, H)
    out = d_self_attention(sd, "self_attn.", out)
    if taps is not None:
        taps["attn"] = out
    for i in range(3):
        out = gblock_down(sd, f"conv.{i}.", out)
    return _d_head(sd, out, class_id, T)


# --------------------------------------------------------------------------------------
# Attention.py (3-D non-local attention; stand-alone modules, not wired into Generator.forward)
# --------------------------------------------------------------------------------------


def attention3d(sd, prefix, x, pooling_factor=2):
    """Module/Attention.py:153-185 (``SelfAttention``, 5-D input; the 4-D path is dead)."""
    B, C, T, W, H = x.shape
    assert T % 2 == 0 and W % 2 == 0 and H % 2 == 0, "T, W, H is not even"
    N = T * W * H
    pf = pooling_factor ** 3
    q = F.conv3d(x, sd[prefix + "query_conv.weight"], sd[prefix + "query_conv.bias"])
    q = q.view(B, -1, N).permute(0, 2, 1)
    k = F.conv3d(x, sd[prefix + "key_conv.weight"], sd[prefix + "key_conv.bias"])
    k = F.max_pool3d(k, kernel_size=2, stride=pooling_factor).view(B, -1, N // pf)
    dist = torch.bmm(q, k)
    attn = torch.softmax(dist, dim=-1)
    v = F.conv3d(x, sd[prefix + "value_conv.weight"], sd[prefix + "value_conv.bias"])
    v = F.max_pool3d(v, kernel_size=2, stride=pooling_factor).view(B, -1, N // pf)
    out = torch.bmm(v, attn.permute(0, 2, 1)).view(B, C, T, W, H)
    return sd[prefix + "gamma"] * out + x


def separable_attn_cell(sd, prefix, x, attn_id, pooling_factor=2):
    """Module/Attention.py:63-111 -- note the raw ``.view`` reinterpretations (memory-order semantics)."""
    B, C, T, W, H = x.shape
    assert T % 2 == 0 and W % 2 == 0 and H % 2 == 0, "T, W, H is not even"
    if attn_id == "T":
        attn_dim, out = T, x[:]
    elif attn_id == "W":
        attn_dim, out = W, x.transpose(2, 3)
    else:
        attn_dim, out = H, x.transpose(2, 4)
    pool = lambda t: F.max_pool3d(t, kernel_size=(2, 1, 1), stride=(pooling_factor, 1, 1))
    q = F.conv3d(out, sd[prefix + "query_conv.weight"], sd[prefix + "query_conv.bias"]).view(B, attn_dim, -1)
    k = F.conv3d(out, sd[prefix + "key_conv.weight"], sd[prefix + "key_conv.bias"])
    k = pool(k).view(B, -1, attn_dim // pooling_factor)
    dist = torch.bmm(q, k)
    score = torch.softmax(dist, dim=-1)
    v = F.conv3d(out, sd[prefix + "value_conv.weight"], sd[prefix + "value_conv.bias"])
    v = pool(v).view(B, -1, attn_dim // pooling_factor)
    out = torch.bmm(v, score.transpose(2, 1))
    if attn_id == "T":
        out = out.view(B, C, W, H, T).permute(0, 1, 4, 2, 3)
    elif attn_id == "W":
        out = out.view(B, C, T, H, W).permute(0, 1, 2, 4, 3)
    else:
        out = out.view(B, C, T, W, H)
    return sd[prefix + "gamma"] * out + x


def separable_attn(sd, prefix, x):
    """Module/Attention.py:8-20: T, W, H cells in sequence (``model.{0,1,2}``)."""
    for i, a in enumerate("TWH"):
        x = separable_attn_cell(sd, f"{prefix}model.{i}.", x, a)
    return x


# --------------------------------------------------------------------------------------
# utils.py / trainer.py
# --------------------------------------------------------------------------------------


def sample_k_frames(data, video_length, k_sample):
    """utils.py:60-63 (consumes one ``torch.randperm`` from the default CPU generator)."""
    frame_idx = torch.randperm(video_length)
    srt, _ = frame_idx[:k_sample].sort()
    return data[:, srt, :, :, :]


def vid_downsample(data):
    """utils.py:77-83 (phi): 2x2 avg-pool per frame + permute to (B,C,T,H/2,W/2)."""
    B, T, C, H, W = data.shape
    x = F.avg_pool2d(data.view(B * T, C, H, W), kernel_size=2)
    _, _, H, W = x.shape
    return x.view(B, T, C, H, W).permute(0, 2, 1, 3, 4).contiguous()


def calc_loss(x, real_flag, adv_loss="hinge"):
    """trainer.py:114-121."""
    if real_flag is True:
        x = -x
    if adv_loss == "wgan-gp":
        return torch.mean(x)
    return torch.relu(1.0 + x).mean()


def trainable_keys(sd):
    """Keys Adam sees (trainer.py:136-141 filters ``requires_grad``): everything floating that
    is not an SN u/v vector or a BatchNorm buffer."""
    out = []
    for k, v in sd.items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            continue
        if "bn.running_" in k or k.endswith("num_batches_tracked"):
            continue
        if not torch.is_floating_point(v):
            continue
        out.append(k)
    return out


class OracleTrainer:
    """trainer.py:189-307 restated: one alternating D_s -> D_t -> G update per ``step`` call.

    State dicts are taken by reference and updated in place.  RNG consumption order on the
    default CPU generator matches SURVEY.md section 3.1 (randperm, randn, randint, randperm).
    """

    def __init__(self, g_sd, ds_sd, dt_sd, *, n_frames, k_sample, n_class, batch_size, g_chn,
                 z_dim=120, latent_dim=4, adv_loss="hinge", g_lr=5e-5, d_lr=5e-5, beta1=0.0, beta2=0.9):
        self.g, self.ds, self.dt = g_sd, ds_sd, dt_sd
        self.n_frames, self.k_sample, self.n_class = n_frames, k_sample, n_class
        self.batch_size, self.g_chn, self.z_dim, self.latent_dim = batch_size, g_chn, z_dim, latent_dim
        self.adv_loss = adv_loss
        for sd in (g_sd, ds_sd, dt_sd):
            for k in trainable_keys(sd):
                sd[k].requires_grad_(True)
        mk = lambda sd, lr: torch.optim.Adam([sd[k] for k in trainable_keys(sd)], lr, (beta1, beta2))
        self.g_opt, self.ds_opt, self.dt_opt = mk(g_sd, g_lr), mk(ds_sd, d_lr), mk(dt_sd, d_lr)

    def reset_grad(self):
        self.ds_opt.zero_grad()
        self.dt_opt.zero_grad()
        self.g_opt.zero_grad()

    def step(self, real_videos, real_labels):
        """real_videos (B,C,T,H,W) as the loader yields them (trainer.py:227 permutes)."""
        real_videos = real_videos.permute(0, 2, 1, 3, 4).contiguous()
        real_s = sample_k_frames(real_videos, self.n_frames, self.k_sample)
        z = torch.randn(self.batch_size, self.z_dim)
        z_class = torch.randint(low=0, high=self.n_class, size=(self.batch_size,))
        fake = generator_forward(self.g, z, z_class, self.n_frames, self.g_chn, self.latent_dim)
        fake_s = sample_k_frames(fake, self.n_frames, self.k_sample)
        ds_real = spatial_discriminator(self.ds, real_s, real_labels)
        ds_fake = spatial_discriminator(self.ds, fake_s.detach(), z_class)
        ds_loss = calc_loss(ds_real, True, self.adv_loss) + calc_loss(ds_fake, False, self.adv_loss)
        self.reset_grad()
        ds_loss.backward()
        self.ds_opt.step()
        real_d = vid_downsample(real_videos)
        fake_d = vid_downsample(fake)
        dt_real = temporal_discriminator(self.dt, real_d, real_labels)
        dt_fake = temporal_discriminator(self.dt, fake_d.detach(), z_class)
        dt_loss = calc_loss(dt_real, True, self.adv_loss) + calc_loss(dt_fake, False, self.adv_loss)
        self.reset_grad()
        dt_loss.backward()
        self.dt_opt.step()
        g_s = spatial_discriminator(self.ds, fake_s, z_class)
        g_t = temporal_discriminator(self.dt, fake_d, z_class)
        g_loss = calc_loss(g_s, True, self.adv_loss) + calc_loss(g_t, True, self.adv_loss)
        self.reset_grad()
        g_loss.backward()
        self.g_opt.step()
        return {"ds_loss": float(ds_loss.detach()), "dt_loss": float(dt_loss.detach()), "g_loss": float(g_loss.detach())}
