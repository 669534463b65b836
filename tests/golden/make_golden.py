"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run once in the build container (needs /root/reference, which is not present on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/*.pt.  Shims, none touching reference files:
  1. a stub ``tensorboardX`` module (imported at the top of several reference files, not installed);
  2. ``.cuda()`` no-ops so ``trainer.py:349-351`` runs on CPU;
  3. a fake loader yielding synthetic (B,3,T,H,W) clips.
Fixture hygiene: every attention ``gamma`` is set non-zero (its init 0 would hide q/k/v bugs);
state_dicts are snapshotted BEFORE each forward because SpectralNorm mutates u/v on every call.
"""
import argparse
import copy
import os
import sys
import types

import torch

REF = os.environ.get("DVDGAN_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_shims():
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
    sys.modules["tensorboardX"] = tbx
    sys.path.insert(0, REF)


def snap(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def grads_of(module):
    return {k: (p.grad.detach().clone() if p.grad is not None else None)
            for k, p in module.named_parameters()}


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def make_generator():
    from Module.Generator import Generator
    torch.manual_seed(1234)
    B, T, ch, n_class = 3, 4, 2, 3
    G = Generator(in_dim=120, latent_dim=4, n_class=n_class, ch=ch, n_frames=T)
    G.train()
    taps = {}
    hooks = []
    for k, m in enumerate(G.conv):
        hooks.append(m.register_forward_hook(
            lambda mod, inp, out, k=k: taps.__setitem__(f"stage{k}", out)))
    hooks.append(G.colorize.register_forward_hook(lambda mod, inp, out: taps.__setitem__("pre_tanh", out)))
    z = torch.randn(B, 120)
    cls = torch.tensor([2, 0, 1])
    pre = snap(G)
    out = G(z, cls)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    # ConvGRU stages return lists per frame inside the loop; keep only tensor-valued taps
    taps = {k: v.detach().clone() for k, v in taps.items() if torch.is_tensor(v)}
    for h in hooks:
        h.remove()
    post = snap(G)
    changed = lambda a, b: {k: v for k, v in b.items() if not torch.equal(a[k], v)}
    # eval-mode forward (running stats used; u/v still advance, Q3); its "pre" state is ``post``
    G.eval()
    with torch.no_grad():
        out_e = G(z, cls)
    post_e = snap(G)
    save("generator.pt", dict(cfg=dict(B=B, T=T, ch=ch, n_class=n_class, latent_dim=4),
                              sd_pre=pre, sd_post_changed=changed(pre, post), z=z, class_id=cls,
                              out=out.detach(), taps=taps, loss_weight=wgt, grads=grads_of(G),
                              eval_out=out_e, eval_sd_post_changed=changed(post, post_e)))


def _set_gammas(net, val):
    for n, p in net.named_parameters():
        if n.endswith("gamma"):
            p.data.fill_(val)


def make_discriminators():
    from Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
    torch.manual_seed(4321)
    n_class = 3
    Ds = SpatialDiscriminator(chn=2, n_class=n_class)
    _set_gammas(Ds, 0.7)
    x = (torch.rand(2, 3, 3, 64, 64) * 2 - 1).requires_grad_(True)
    cls = torch.tensor([1, 2])
    pre = snap(Ds)
    out = Ds(x, cls)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    save("spatial_d.pt", dict(cfg=dict(chn=2, n_class=n_class), sd_pre=pre, sd_post=snap(Ds),
                              x=x.detach(), class_id=cls, out=out.detach(), loss_weight=wgt,
                              grads=grads_of(Ds), dx=x.grad.clone()))
    # a second forward from the advanced u/v state (pins the stateful power iteration)
    with torch.no_grad():
        out2 = Ds(x.detach(), cls)
    save("spatial_d_second.pt", dict(out=out2, sd_post=snap(Ds)))

    Dt = TemporalDiscriminator(chn=2, n_class=n_class)
    _set_gammas(Dt, -0.5)
    x = (torch.rand(2, 3, 8, 32, 32) * 2 - 1).requires_grad_(True)
    pre = snap(Dt)
    out = Dt(x, cls)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    save("temporal_d.pt", dict(cfg=dict(chn=2, n_class=n_class), sd_pre=pre, sd_post=snap(Dt),
                               x=x.detach(), class_id=cls, out=out.detach(), loss_weight=wgt,
                               grads=grads_of(Dt), dx=x.grad.clone()))


def make_blocks():
    from Module.ConvGRU import ConvGRU, ConvGRUCell
    from Module.GResBlock import GResBlock
    from Module.Normalization import ConditionalNorm, SpectralNorm
    from Module.Attention import SelfAttention, SeparableAttn
    import torch.nn as nn
    torch.manual_seed(99)
    fx = {}
    # ConvGRUCell k=3 and k=5, with and without previous state
    for name, (cin, ch, k) in {"cell_k3": (5, 6, 3), "cell_k5": (4, 8, 5)}.items():
        cell = ConvGRUCell(cin, ch, k)
        for p in cell.parameters():          # biases are zero-initialised; make them matter
            if p.dim() == 1:
                p.data.normal_(0, 0.1)
        x = torch.randn(3, cin, 6, 6, requires_grad=True)
        h = torch.randn(3, ch, 6, 6, requires_grad=True)
        y0 = cell(x)
        y1 = cell(x, h)
        wgt = torch.randn_like(y1)
        (y1 * wgt).sum().backward()
        fx[name] = dict(sd=snap(cell), x=x.detach(), h=h.detach(), y_nostate=y0.detach(), y=y1.detach(),
                        loss_weight=wgt, grads=grads_of(cell), dx=x.grad.clone(), dh=h.grad.clone())
    # multi-layer ConvGRU over a short sequence (the Generator's inner loop)
    gru = ConvGRU(4, hidden_sizes=[4, 8, 4], kernel_sizes=[3, 5, 3], n_layers=3)
    xs = torch.randn(2, 3, 4, 8, 8, requires_grad=True)   # B,T,C,H,W
    hid, outs = None, []
    for t in range(3):
        hid = gru(xs[:, t], hid)
        outs.append(hid[-1])
    y = torch.stack(outs, 1)
    wgt = torch.randn_like(y)
    (y * wgt).sum().backward()
    fx["gru_seq"] = dict(sd=snap(gru), x=xs.detach(), y=y.detach(), loss_weight=wgt,
                         grads=grads_of(gru), dx=xs.grad.clone())
    # ConditionalNorm
    cn = ConditionalNorm(5, 7)
    x = torch.randn(6, 5, 4, 4, requires_grad=True)
    c = torch.randn(6, 7, requires_grad=True)
    pre = snap(cn)
    y = cn(x, c)
    wgt = torch.randn_like(y)
    (y * wgt).sum().backward()
    fx["cbn"] = dict(sd_pre=pre, sd_post=snap(cn), x=x.detach(), cond=c.detach(), y=y.detach(),
                     loss_weight=wgt, grads=grads_of(cn), dx=x.grad.clone(), dcond=c.grad.clone())
    # GResBlock, both upsample factors
    for up in (1, 2):
        blk = GResBlock(6, 4, n_class=10, upsample_factor=up)
        x = torch.randn(4, 6, 4, 4, requires_grad=True)
        c = torch.randn(2, 10).repeat(2, 1).requires_grad_(True)
        pre = snap(blk)
        y = blk(x, c)
        wgt = torch.randn_like(y)
        (y * wgt).sum().backward()
        fx[f"gres_up{up}"] = dict(sd_pre=pre, sd_post=snap(blk), x=x.detach(), cond=c.detach(), y=y.detach(),
                                  loss_weight=wgt, grads=grads_of(blk), dx=x.grad.clone(), dcond=c.grad.clone())
    # SpectralNorm around Conv2d / Conv3d / Linear / Embedding, two consecutive calls
    sn_cases = {
        "sn_conv2d": (SpectralNorm(nn.Conv2d(3, 5, 3, padding=1)), torch.randn(2, 3, 5, 5)),
        "sn_conv3d": (SpectralNorm(nn.Conv3d(2, 4, 3, padding=1)), torch.randn(2, 2, 4, 5, 5)),
        "sn_linear": (SpectralNorm(nn.Linear(6, 1)), torch.randn(4, 6)),
        "sn_embed": (SpectralNorm(nn.Embedding(4, 6)), torch.tensor([3, 0, 0, 2])),
    }
    for name, (m, x) in sn_cases.items():
        pre = snap(m)
        if x.is_floating_point():
            x.requires_grad_(True)
        y = m(x)
        wgt = torch.randn_like(y)
        (y * wgt).sum().backward()
        mid = snap(m)
        with torch.no_grad():
            y2 = m(x.detach())
        fx[name] = dict(sd_pre=pre, sd_mid=mid, sd_post=snap(m), x=x.detach(), y=y.detach(), y2=y2,
                        loss_weight=wgt, grads=grads_of(m),
                        dx=(x.grad.clone() if x.is_floating_point() else None))
    # 3-D attention modules (stand-alone; Attention.py)
    sa = SelfAttention(8)
    _set_gammas(sa, 0.9)
    x = torch.randn(2, 8, 4, 4, 6, requires_grad=True)
    y = sa(x)
    wgt = torch.randn_like(y)
    (y * wgt).sum().backward()
    fx["attn3d"] = dict(sd=snap(sa), x=x.detach(), y=y.detach(), loss_weight=wgt, grads=grads_of(sa),
                        dx=x.grad.clone())
    sp = SeparableAttn(4)
    _set_gammas(sp, 0.6)
    x = torch.randn(2, 4, 4, 6, 8, requires_grad=True)
    y = sp(x)
    wgt = torch.randn_like(y)
    (y * wgt).sum().backward()
    fx["sep_attn"] = dict(sd=snap(sp), x=x.detach(), y=y.detach(), loss_weight=wgt, grads=grads_of(sp),
                          dx=x.grad.clone())
    save("blocks.pt", fx)


def make_helpers():
    from utils import sample_k_frames, vid_downsample
    torch.manual_seed(7)
    data = torch.randn(2, 6, 3, 8, 8)
    torch.manual_seed(11)
    s = sample_k_frames(data, 6, 3)
    torch.manual_seed(11)
    s_all = sample_k_frames(data, 6, 64)    # k >= T: all frames, sorted -> identity order
    d = vid_downsample(data)
    save("helpers.pt", dict(data=data, seed=11, sample_k3=s, sample_all=s_all, phi=d))


class _Loader:
    def __init__(self, clips, labels):
        self.clips, self.labels = clips, labels

    def __len__(self):
        return len(self.clips)

    def __iter__(self):
        return iter(zip(self.clips, self.labels))


def make_step():
    """trainer.py, config 1 of BASELINE.json (4 frames, B=2, run at 64x64: Dt cannot do 32x32, Q12)."""
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    import trainer as ref_trainer
    n_steps = 2
    cfg = argparse.Namespace(
        model="dvd-gan", adv_loss="hinge", imsize=64, g_num=5, z_dim=120, g_chn=2, ds_chn=2, dt_chn=2,
        n_frames=4, g_conv_dim=64, d_conv_dim=64, lr_schr="const", lambda_gp=10, total_epoch=1, d_iters=1,
        g_iters=1, batch_size=2, num_workers=0, g_lr=5e-5, d_lr=5e-5, lr_decay=0.9999, beta1=0.0, beta2=0.9,
        pretrained_model=None, n_class=2, k_sample=2, dataset="synthetic", use_tensorboard=False,
        test_batch_size=1, image_path="", log_path="/tmp/dvd_golden/log", model_save_path="/tmp/dvd_golden/m",
        sample_path="/tmp/dvd_golden/s", log_epoch=10 ** 6, sample_epoch=10 ** 6, model_save_epoch=10 ** 6,
        version="golden", gpus="", parallel=False)
    torch.manual_seed(2024)
    clips = [torch.rand(2, 3, 4, 64, 64) * 2 - 1 for _ in range(n_steps)]
    labels = [torch.randint(0, 2, (2,)) for _ in range(n_steps)]
    tr = ref_trainer.Trainer(_Loader(clips, labels), cfg)
    _set_gammas(tr.D_s, 0.3)
    _set_gammas(tr.D_t, -0.4)
    pre = dict(G=snap(tr.G), Ds=snap(tr.D_s), Dt=snap(tr.D_t))
    losses = []
    torch.manual_seed(77)       # seed for the step's own RNG draws (randperm / randn / randint)
    # NB trainer.py:195 draws fixed_z = randn(test_batch_size*n_class, z_dim) before the first step.
    # capture the three losses of every step by wrapping torch.Tensor.backward
    bw = torch.Tensor.backward

    def backward(self, *a, **k):
        losses.append(float(self.detach()))
        return bw(self, *a, **k)
    torch.Tensor.backward = backward
    try:
        tr.train()
    finally:
        torch.Tensor.backward = bw
    assert len(losses) == 3 * n_steps, losses
    post = dict(G=snap(tr.G), Ds=snap(tr.D_s), Dt=snap(tr.D_t))
    save("step.pt", dict(cfg=vars(cfg), clips=clips, labels=labels, rng_seed=77, sd_pre=pre, sd_post=post,
                         losses=losses))


if __name__ == "__main__":
    _install_shims()
    torch.set_num_threads(1)          # summation order independent of the machine's core count
    make_helpers()
    make_blocks()
    make_discriminators()
    make_generator()
    make_step()
