"""Full-size golden fixtures from the UNMODIFIED reference (CPU), for BASELINE.json configs 2-5 at B = 1.

    python tests/golden/make_golden_full.py [c3 c4 c5 c2step]

ch = 32 networks are 137 M / 10 M / 10 M parameters (0.6 GB), and a 48-frame 128x128 clip is 9 MB, so these fixtures
hold (a) the SEED the networks are constructed from (tests rebuild them with this repo's constructors, whose same-seed
parity with the reference is pinned by tests/test_gpu_parity.py::test_state_dict_roundtrip_and_init) plus a per-tensor
fingerprint of the initial state, (b) the small inputs, and (c) SAMPLES of every output / per-stage activation /
gradient at index sets derived from the tensor's name (``sample_idx``), together with each tensor's full L2 norm.
Writes tests/golden/full_<name>.pt (a few MB each)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from full_fixture import sample_idx  # noqa: E402
from make_golden import _Loader, _install_shims, _set_gammas, save  # noqa: E402


def sample(name, t, k):
    f = t.detach().reshape(-1)
    return dict(v=f[sample_idx(name, f.numel(), k)].clone(), norm=float(f.double().norm()), n=f.numel())


def fingerprint(module):
    return {k: float(v.double().abs().sum()) for k, v in module.state_dict().items() if torch.is_floating_point(v)}


def seeded(shape, seed, kind="randn"):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) if kind == "randn" else torch.rand(shape, generator=g) * 2 - 1


CASES = {
    # BASELINE.json configs[2]: 48 frames 128x128;  [3]: 12 frames 256x256, 600 classes (N = 4096 attention tokens
    # in Ds);  [4]: 128 frames 64x64
    "c3": dict(ld=8, T=48, n_class=101, k=8, seed=3003),
    "c4": dict(ld=16, T=12, n_class=600, k=8, seed=4004),
    "c5": dict(ld=4, T=128, n_class=101, k=8, seed=5005),
}


def _sampled_rel(a, b):
    """global rel-L2 between two dicts of samples (name -> dict(v=...))"""
    num = sum(float((a[k]["v"].double() - b[k]["v"].double()).norm() ** 2) for k in b)
    den = sum(float(b[k]["v"].double().norm() ** 2) for k in b)
    return (num / max(den, 1e-300)) ** 0.5


def make_case(name, c):
    """The fixture proper comes from a run on all cores; a second run of the same reference code on HALF the cores
    (a different summation order inside MKL-DNN, nothing else) measures the reference's own run-to-run noise on the
    sampled gradients, which the GPU tests use as their yardstick."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    fx = run_case(name, c)
    torch.set_num_threads(max(1, n // 2))
    fx2 = run_case(name + " (half the threads)", c)
    torch.set_num_threads(n)
    fx["noise"] = dict(threads=(n, max(1, n // 2)),
                       G_out=float((fx["G"]["out"]["v"] - fx2["G"]["out"]["v"]).norm() / fx["G"]["out"]["v"].norm()),
                       G_grads=_sampled_rel(fx2["G"]["grads"], fx["G"]["grads"]),
                       Ds_grads=_sampled_rel(fx2["Ds"]["grads"], fx["Ds"]["grads"]),
                       Dt_grads=_sampled_rel(fx2["Dt"]["grads"], fx["Dt"]["grads"]))
    print(name, "reference noise:", fx["noise"], flush=True)
    save(f"full_{name}.pt", fx)


def run_case(name, c):
    from Module.Generator import Generator
    from Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
    t0 = time.time()
    ch, ld, T, ncls = 32, c["ld"], c["T"], c["n_class"]
    side = 16 * ld
    torch.manual_seed(c["seed"])
    G = Generator(in_dim=120, latent_dim=ld, n_class=ncls, ch=ch, n_frames=T)
    Ds = SpatialDiscriminator(chn=ch, n_class=ncls)
    Dt = TemporalDiscriminator(chn=ch, n_class=ncls)
    _set_gammas(Ds, 0.5)
    _set_gammas(Dt, -0.5)
    fx = dict(cfg=dict(ch=ch, ld=ld, T=T, n_class=ncls, k=c["k"], seed=c["seed"], gamma_s=0.5, gamma_t=-0.5),
              fp=dict(G=fingerprint(G), Ds=fingerprint(Ds), Dt=fingerprint(Dt)))
    # ---- G: forward + backward of a linear loss
    G.train()
    taps = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out, k=k: taps.__setitem__(f"stage{k}", out))
             for k, m in enumerate(G.conv)]
    hooks.append(G.colorize.register_forward_hook(lambda mod, inp, out: taps.__setitem__("pre_tanh", out)))
    z = seeded((1, 120), c["seed"] + 1)
    cls = torch.tensor([c["seed"] % ncls])
    out = G(z, cls)
    print(f"{name}: G forward {time.time() - t0:.0f}s", flush=True)
    wgt = seeded(tuple(out.shape), c["seed"] + 2)
    (out * wgt).sum().backward()
    print(f"{name}: G backward {time.time() - t0:.0f}s", flush=True)
    for h in hooks:
        h.remove()
    fx["G"] = dict(z=z, class_id=cls, out=sample("out", out, 1 << 18),
                   taps={k: sample(k, v, 1 << 16) for k, v in taps.items() if torch.is_tensor(v)},
                   grads={k: sample("g." + k, p.grad, 1024) for k, p in G.named_parameters() if p.grad is not None},
                   saturated=float((out.detach().abs() > 0.999).float().mean()))
    del out, wgt, taps
    # ---- Ds on k frames, Dt on the phi-sized clip: forward + backward of a linear loss (synthetic inputs in [-1, 1])
    xs = seeded((1, c["k"], 3, side, side), c["seed"] + 3, "rand").requires_grad_(True)
    o = Ds(xs, cls)
    w = seeded(tuple(o.shape), c["seed"] + 4)
    (o * w).sum().backward()
    fx["Ds"] = dict(out=o.detach().clone(), dx=sample("ds.dx", xs.grad, 1 << 16),
                    grads={k: sample("gs." + k, p.grad, 1024) for k, p in Ds.named_parameters() if p.grad is not None})
    xt = seeded((1, 3, T, side // 2, side // 2), c["seed"] + 5, "rand").requires_grad_(True)
    o = Dt(xt, cls)
    w = seeded(tuple(o.shape), c["seed"] + 6)
    (o * w).sum().backward()
    fx["Dt"] = dict(out=o.detach().clone(), dx=sample("dt.dx", xt.grad, 1 << 16),
                    grads={k: sample("gt." + k, p.grad, 1024) for k, p in Dt.named_parameters() if p.grad is not None})
    print(f"{name}: done {time.time() - t0:.0f}s", flush=True)
    return fx


def make_step_full():
    """Two steps of the reference Trainer at config-2 width (ch = 32, 48 frames, 64x64, 101 classes, k = 8), 1 clip;
    run twice (all cores / half the cores) so that the fixture also holds the reference's own noise on the parameter
    updates (beta1 = 0: the first Adam step is lr * sign(g), so every near-zero gradient whose sign flips under
    summation-order noise moves its parameter by 2 * lr)."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    fx = run_step_full()
    torch.set_num_threads(max(1, n // 2))
    fx2 = run_step_full()
    torch.set_num_threads(n)
    fx["noise"] = dict(threads=(n, max(1, n // 2)),
                       losses=[abs(a - b) for a, b in zip(fx["losses"], fx2["losses"])],
                       delta={k: _sampled_rel(fx2["delta"][k], fx["delta"][k]) for k in fx["delta"]})
    print("c2step reference noise:", fx["noise"], flush=True)
    save("full_c2step.pt", fx)


def run_step_full():
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    import trainer as ref_trainer
    n_steps, seed = 2, 2002
    cfg = argparse.Namespace(
        model="dvd-gan", adv_loss="hinge", imsize=64, g_num=5, z_dim=120, g_chn=32, ds_chn=32, dt_chn=32,
        n_frames=48, g_conv_dim=64, d_conv_dim=64, lr_schr="const", lambda_gp=10, total_epoch=1, d_iters=1,
        g_iters=1, batch_size=1, num_workers=0, g_lr=5e-5, d_lr=5e-5, lr_decay=0.9999, beta1=0.0, beta2=0.9,
        pretrained_model=None, n_class=101, k_sample=8, dataset="synthetic", use_tensorboard=False,
        test_batch_size=1, image_path="", log_path="/tmp/dvd_golden/log", model_save_path="/tmp/dvd_golden/m",
        sample_path="/tmp/dvd_golden/s", log_epoch=10 ** 6, sample_epoch=10 ** 6, model_save_epoch=10 ** 6,
        version="golden_full", gpus="", parallel=False)
    clips = [seeded((1, 3, 48, 64, 64), seed + 10 + i, "rand") for i in range(n_steps)]
    labels = [torch.tensor([(seed + i) % 101]) for i in range(n_steps)]
    torch.manual_seed(seed)
    tr = ref_trainer.Trainer(_Loader(clips, labels), cfg)
    _set_gammas(tr.D_s, 0.3)
    _set_gammas(tr.D_t, -0.4)
    nets = dict(G=tr.G, Ds=tr.D_s, Dt=tr.D_t)
    fp = {k: fingerprint(n) for k, n in nets.items()}
    pre = {k: {n: p.detach().clone() for n, p in net.named_parameters() if p.requires_grad} for k, net in nets.items()}
    losses = []
    torch.manual_seed(77)
    bw = torch.Tensor.backward

    def backward(self, *a, **k):
        losses.append(float(self.detach()))
        print("loss", losses[-1], flush=True)
        return bw(self, *a, **k)
    torch.Tensor.backward = backward
    try:
        tr.train()
    finally:
        torch.Tensor.backward = bw
    delta = {k: {n: sample(f"d.{k}.{n}", p.detach() - pre[k][n], 1024)
                 for n, p in net.named_parameters() if p.requires_grad} for k, net in nets.items()}
    return dict(cfg=vars(cfg), seed=seed, rng_seed=77, n_steps=n_steps, gamma_s=0.3, gamma_t=-0.4, fp=fp,
                losses=losses, delta=delta)


if __name__ == "__main__":
    _install_shims()
    todo = sys.argv[1:] or ["c5", "c4", "c3", "c2step"]
    for t in todo:
        if t == "c2step":
            make_step_full()
        else:
            make_case(t, CASES[t])
