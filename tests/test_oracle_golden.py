"""Pin the CPU oracle (oracle/dvdgan_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Same torch ops in the same order on the
same machine class => bit-exact forward; gradients are compared at 1e-6 relative (autograd
may legally reorder accumulation)."""
import pytest
import torch

from oracle import dvdgan_oracle as O
from conftest import clone_sd

torch.set_num_threads(1)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _check_grads(sd, grads, tol=2e-6, skip_zero=True):
    for k, g in grads.items():
        if g is None:
            continue
        got = sd[k].grad
        assert got is not None, k
        if g.norm() < 1e-7:          # analytically-zero grads (conv0 bias before a BatchNorm)
            assert got.norm() < 1e-5, k
            continue
        assert _rel(got, g) < tol, (k, _rel(got, g))


def test_helpers(golden):
    fx = golden("helpers.pt")
    torch.manual_seed(fx["seed"])
    assert torch.equal(O.sample_k_frames(fx["data"], 6, 3), fx["sample_k3"])
    torch.manual_seed(fx["seed"])
    assert torch.equal(O.sample_k_frames(fx["data"], 6, 64), fx["sample_all"])
    assert torch.equal(fx["sample_all"], fx["data"])      # k >= T: identity order
    assert torch.equal(O.vid_downsample(fx["data"]), fx["phi"])


@pytest.mark.parametrize("name", ["cell_k3", "cell_k5"])
def test_convgru_cell(golden, name):
    fx = golden("blocks.pt")[name]
    sd = clone_sd(fx["sd"], fx["sd"].keys())
    x = fx["x"].clone().requires_grad_(True)
    h = fx["h"].clone().requires_grad_(True)
    assert torch.equal(O.convgru_cell(sd, "", x), fx["y_nostate"])
    y = O.convgru_cell(sd, "", x, h)
    assert torch.equal(y, fx["y"])
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6 and _rel(h.grad, fx["dh"]) < 2e-6


def test_convgru_seq(golden):
    fx = golden("blocks.pt")["gru_seq"]
    sd = clone_sd(fx["sd"], fx["sd"].keys())
    xs = fx["x"].clone().requires_grad_(True)
    hid, outs = None, []
    for t in range(xs.shape[1]):
        hid = O.convgru(sd, "", xs[:, t], hid)
        outs.append(hid[-1])
    y = torch.stack(outs, 1)
    assert torch.equal(y, fx["y"])
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(xs.grad, fx["dx"]) < 2e-6


def test_cbn(golden):
    fx = golden("blocks.pt")["cbn"]
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))
    x = fx["x"].clone().requires_grad_(True)
    c = fx["cond"].clone().requires_grad_(True)
    y = O.conditional_norm(sd, "", x, c)
    assert torch.equal(y, fx["y"])
    for k, v in fx["sd_post"].items():
        assert torch.equal(sd[k].detach(), v), k
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6 and _rel(c.grad, fx["dcond"]) < 2e-6


@pytest.mark.parametrize("up", [1, 2])
def test_gresblock(golden, up):
    fx = golden("blocks.pt")[f"gres_up{up}"]
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))
    x = fx["x"].clone().requires_grad_(True)
    c = fx["cond"].clone().requires_grad_(True)
    y = O.gresblock(sd, "", x, c, up)
    assert torch.equal(y, fx["y"])
    for k, v in fx["sd_post"].items():
        assert torch.equal(sd[k].detach(), v), k
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6 and _rel(c.grad, fx["dcond"]) < 2e-6


@pytest.mark.parametrize("name", ["sn_conv2d", "sn_conv3d", "sn_linear", "sn_embed"])
def test_spectral_norm(golden, name):
    import torch.nn.functional as F
    fx = golden("blocks.pt")[name]
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))

    def fwd(x):
        w = O.spectral_norm_weight(sd, "module.")
        if name == "sn_conv2d":
            return F.conv2d(x, w, sd["module.bias"], padding=1)
        if name == "sn_conv3d":
            return F.conv3d(x, w, sd["module.bias"], padding=1)
        if name == "sn_linear":
            return F.linear(x, w, sd["module.bias"])
        return F.embedding(x, w)
    x = fx["x"].clone()
    if x.is_floating_point():
        x.requires_grad_(True)
    y = fwd(x)
    assert torch.equal(y, fx["y"])
    for k in ("module.weight_u", "module.weight_v"):
        assert torch.equal(sd[k], fx["sd_mid"][k])
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    with torch.no_grad():
        y2 = fwd(fx["x"])
    assert torch.equal(y2, fx["y2"])           # second call from the advanced u/v (Q3)
    for k in ("module.weight_u", "module.weight_v"):
        assert torch.equal(sd[k], fx["sd_post"][k])


def test_attention3d(golden):
    fx = golden("blocks.pt")["attn3d"]
    sd = clone_sd(fx["sd"], fx["sd"].keys())
    x = fx["x"].clone().requires_grad_(True)
    y = O.attention3d(sd, "", x)
    assert torch.equal(y, fx["y"])
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6
    with pytest.raises(AssertionError):
        O.attention3d(sd, "", torch.randn(1, 8, 3, 4, 4))     # odd T (Attention.py:161)


def test_separable_attn(golden):
    fx = golden("blocks.pt")["sep_attn"]
    sd = clone_sd(fx["sd"], fx["sd"].keys())
    x = fx["x"].clone().requires_grad_(True)
    y = O.separable_attn(sd, "", x)
    assert torch.equal(y, fx["y"])
    (y * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6


def test_spatial_discriminator(golden):
    fx = golden("spatial_d.pt")
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))
    x = fx["x"].clone().requires_grad_(True)
    out = O.spatial_discriminator(sd, x, fx["class_id"])
    assert out.shape == (6,)
    assert torch.equal(out, fx["out"])
    for k, v in fx["sd_post"].items():
        assert torch.equal(sd[k].detach(), v), k
    (out * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6
    fx2 = golden("spatial_d_second.pt")
    with torch.no_grad():
        out2 = O.spatial_discriminator(sd, fx["x"], fx["class_id"])
    assert torch.equal(out2, fx2["out"])
    for k, v in fx2["sd_post"].items():
        assert torch.equal(sd[k].detach(), v), k


def test_temporal_discriminator(golden):
    fx = golden("temporal_d.pt")
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))
    x = fx["x"].clone().requires_grad_(True)
    out = O.temporal_discriminator(sd, x, fx["class_id"])
    assert out.shape == (2 * (8 // 4),)
    assert torch.equal(out, fx["out"])
    for k, v in fx["sd_post"].items():
        assert torch.equal(sd[k].detach(), v), k
    (out * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"])
    assert _rel(x.grad, fx["dx"]) < 2e-6


def test_generator(golden):
    fx = golden("generator.pt")
    cfg = fx["cfg"]
    sd = clone_sd(fx["sd_pre"], O.trainable_keys(fx["sd_pre"]))
    taps = {}
    out = O.generator_forward(sd, fx["z"], fx["class_id"], cfg["T"], cfg["ch"], cfg["latent_dim"], taps=taps)
    assert torch.equal(out, fx["out"])
    for k, v in fx["taps"].items():
        assert torch.equal(taps[k], v), k
    for k, v in fx["sd_post_changed"].items():
        assert torch.equal(sd[k].detach(), v), k
    (out * fx["loss_weight"]).sum().backward()
    _check_grads(sd, fx["grads"], tol=2e-5)
    with torch.no_grad():
        out_e = O.generator_forward(sd, fx["z"], fx["class_id"], cfg["T"], cfg["ch"], cfg["latent_dim"],
                                    training=False)
    assert torch.equal(out_e, fx["eval_out"])
    for k, v in fx["eval_sd_post_changed"].items():
        assert torch.equal(sd[k].detach(), v), k


def test_train_step(golden):
    """trainer.py:213-307 for two steps (BASELINE.json configs[0], at 64x64 because of Q12)."""
    fx = golden("step.pt")
    cfg = fx["cfg"]
    g, ds, dt = (clone_sd(fx["sd_pre"][k]) for k in ("G", "Ds", "Dt"))
    tr = O.OracleTrainer(g, ds, dt, n_frames=cfg["n_frames"], k_sample=cfg["k_sample"], n_class=cfg["n_class"],
                         batch_size=cfg["batch_size"], g_chn=cfg["g_chn"], z_dim=cfg["z_dim"],
                         adv_loss=cfg["adv_loss"], g_lr=cfg["g_lr"], d_lr=cfg["d_lr"],
                         beta1=cfg["beta1"], beta2=cfg["beta2"])
    torch.manual_seed(fx["rng_seed"])
    torch.randn(cfg["test_batch_size"] * cfg["n_class"], cfg["z_dim"])     # fixed_z, trainer.py:195
    losses = []
    for clip, lab in zip(fx["clips"], fx["labels"]):
        r = tr.step(clip, lab)
        losses += [r["ds_loss"], r["dt_loss"], r["g_loss"]]
    assert losses == pytest.approx(fx["losses"], rel=1e-6)
    for name, sd in (("G", g), ("Ds", ds), ("Dt", dt)):
        for k, v in fx["sd_post"][name].items():
            if not torch.is_floating_point(v):
                assert torch.equal(sd[k], v), (name, k)
            else:
                assert torch.allclose(sd[k].detach(), v, rtol=1e-5, atol=1e-7), (name, k)
