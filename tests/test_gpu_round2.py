"""GPU parity / behaviour tests added in round 2 (run with -m gpu on a B200): full-size BASELINE.json configurations
against samples of the unmodified reference (tests/golden/full_*.pt), lean ConvGRU BPTT, range safety of the bf16
operand planes, the ONEACC double-buffered tiles, dtype / index checks of the boundary, checkpoints and sampling,
two devices in one process."""
import argparse
import os
import sys

import pytest
import torch

from conftest import GOLDEN, clone_sd

sys.path.insert(0, GOLDEN)
from full_fixture import sample_idx  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _set_gammas(net, val):
    for n, p in net.named_parameters():
        if n.endswith("gamma"):
            p.data.fill_(val)


def _fingerprint_ok(net, fp, what):
    """The fixtures hold the seed the reference nets were built from, not 0.6 GB of weights: check that this machine's
    same-seed construction reproduces the reference's initial state before comparing anything downstream."""
    for k, v in net.state_dict().items():
        if torch.is_floating_point(v):
            got = float(v.double().abs().sum())
            assert got == pytest.approx(fp[k], rel=1e-5, abs=1e-6), f"{what}.{k}: same-seed init differs from the fixture"


def _seeded(shape, seed, kind="randn"):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) if kind == "randn" else torch.rand(shape, generator=g) * 2 - 1


def _cmp_sample(name, got, fx, tol):
    f = got.detach().reshape(-1)
    assert f.numel() == fx["n"], name
    idx = sample_idx(name, f.numel(), fx["v"].numel()).to(f.device)
    e = rel(f[idx], fx["v"])
    assert e < tol, (name, e)
    return e


def _cmp_grads(prefix, net, gfx, tol_each, tol_global, zero_suffix=None):
    """sampled gradients: every tensor within tol_each (absolute floor for analytically-zero ones), the whole
    sampled gradient within tol_global"""
    num = den = 0.0
    for k, p in net.named_parameters():
        if k not in gfx:
            continue
        assert p.grad is not None, k
        f = p.grad.detach().reshape(-1)
        fx = gfx[k]
        idx = sample_idx(prefix + k, f.numel(), fx["v"].numel()).to(f.device)
        got, want = f[idx].double().cpu(), fx["v"].double()
        if zero_suffix and k.endswith(zero_suffix):          # analytically zero: both sides hold rounding noise only
            scale = max(float(q.grad.abs().max()) for q in net.parameters() if q.grad is not None)
            assert float(got.abs().max()) <= 1e-4 * scale, (k, float(got.abs().max()), scale)
            continue
        d = float((got - want).norm())
        num += d * d
        den += float(want.norm()) ** 2
        assert d <= tol_each * float(want.norm()) + 1e-7, (k, d / (float(want.norm()) + 1e-30))
        full = float(f.double().norm())
        assert full == pytest.approx(fx["norm"], rel=10 * tol_each, abs=1e-6), (k, "norm")
    print(prefix, "sampled gradient global rel-L2: %.3e (bound %.3e)" % ((num / max(den, 1e-300)) ** 0.5, tol_global))
    assert (num / max(den, 1e-300)) ** 0.5 < tol_global, (prefix, (num / den) ** 0.5)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs 3, 4, 5 at full width and full clip length (B = 1) vs the unmodified reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_full_size_config_vs_reference(dev, name):
    """ch = 32; c3: 48 frames 128x128, c4: 12 frames 256x256 / 600 classes (N = 4096 attention tokens in Ds),
    c5: 128 frames 64x64.  Forward (output, pre-tanh, every stage) within 1e-3 rel-L2 (north_star).  Backward of a
    linear loss: end-to-end gradients through 4 ConvGRU stages and 16 CBNs are noisy in the REFERENCE itself (ReLU kinks
    flip under summation-order noise), so the fixture carries the reference's own noise -- the same code run
    on half the CPU threads, whose forward differs by 6e-6..3e-5 -- and the global bound is 4x that (measured here:
    1.1-1.3e-2 at a forward difference of 0.8-1.3e-4, i.e. ~3x the reference's own 4e-3; never below 1.5e-2 / 5e-2 per
    tensor).  Every kernel's gradients are held to 1e-3 on identical inputs by the per-kernel tests.  Ds / Dt forward +
    backward on synthetic clips of the same size (ReLU kinks again: 5e-3 on the input gradient)."""
    from dvdgan_b200.Module.Generator import Generator
    from dvdgan_b200.Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
    path = os.path.join(GOLDEN, f"full_{name}.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    fx = torch.load(path, weights_only=False)
    c = fx["cfg"]
    torch.manual_seed(c["seed"])
    G = Generator(in_dim=120, latent_dim=c["ld"], n_class=c["n_class"], ch=c["ch"], n_frames=c["T"])
    Ds = SpatialDiscriminator(chn=c["ch"], n_class=c["n_class"])
    Dt = TemporalDiscriminator(chn=c["ch"], n_class=c["n_class"])
    _set_gammas(Ds, c["gamma_s"])
    _set_gammas(Dt, c["gamma_t"])
    for net, key in ((G, "G"), (Ds, "Ds"), (Dt, "Dt")):
        _fingerprint_ok(net, fx["fp"][key], key)
    side = 16 * c["ld"]
    G.to(dev).train()
    taps = {}
    g = fx["G"]
    out = G(g["z"].to(dev), g["class_id"].to(dev), taps=taps)
    assert out.shape == (1, c["T"], 3, side, side)
    errs = {k: _cmp_sample(k, taps[k], v, 1e-3) for k, v in g["taps"].items()}
    errs["out"] = _cmp_sample("out", out, g["out"], 1e-3)
    print(name, "forward rel-L2:", {k: f"{v:.2e}" for k, v in errs.items()})
    wgt = _seeded(tuple(out.shape), c["seed"] + 2).to(dev)
    (out * wgt).sum().backward()
    noise = fx.get("noise", {})
    print(name, "reference noise (all cores vs half):", noise)
    ng = noise.get("G_grads", 0.0)
    _cmp_grads("g.", G, g["grads"], max(5e-2, 10 * ng), max(1.5e-2, 4 * ng), zero_suffix="conv0.module.bias")
    del out, wgt, taps
    G.cpu()
    torch.cuda.empty_cache()
    cls = g["class_id"].to(dev)
    Ds.to(dev)
    xs = _seeded((1, c["k"], 3, side, side), c["seed"] + 3, "rand").to(dev).requires_grad_(True)
    o = Ds(xs, cls)
    assert rel(o, fx["Ds"]["out"]) < 1e-3
    (o * _seeded(tuple(o.shape), c["seed"] + 4).to(dev)).sum().backward()
    _cmp_sample("ds.dx", xs.grad, fx["Ds"]["dx"], 5e-3)
    # key_conv.bias: softmax is invariant to a constant added to every logit of a row, so this gradient is exactly zero
    _cmp_grads("gs.", Ds, fx["Ds"]["grads"], 5e-3, max(2e-3, 3 * noise.get("Ds_grads", 0.0)), zero_suffix="key_conv.bias")
    Dt.to(dev)
    xt = _seeded((1, 3, c["T"], side // 2, side // 2), c["seed"] + 5, "rand").to(dev).requires_grad_(True)
    o = Dt(xt, cls)
    assert rel(o, fx["Dt"]["out"]) < 1e-3
    (o * _seeded(tuple(o.shape), c["seed"] + 6).to(dev)).sum().backward()
    _cmp_sample("dt.dx", xt.grad, fx["Dt"]["dx"], 5e-3)
    _cmp_grads("gt.", Dt, fx["Dt"]["grads"], 5e-3, max(2e-3, 3 * noise.get("Dt_grads", 0.0)), zero_suffix="key_conv.bias")


def test_full_width_two_steps_vs_reference_trainer(dev):
    """Two G+Ds+Dt steps of the reference Trainer at config-2 width (ch = 32, 48 frames, 64x64, 101 classes, k = 8) on one
    clip: the six losses and samples of every parameter's update."""
    from dvdgan_b200.trainer import Trainer
    path = os.path.join(GOLDEN, "full_c2step.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    fx = torch.load(path, weights_only=False)
    cfg = argparse.Namespace(**fx["cfg"])
    clips = [_seeded((1, 3, 48, 64, 64), fx["seed"] + 10 + i, "rand") for i in range(fx["n_steps"])]
    labels = [torch.tensor([(fx["seed"] + i) % 101]) for i in range(fx["n_steps"])]

    class Loader:
        def __len__(self):
            return len(clips)

        def __iter__(self):
            return iter(zip(clips, labels))
    torch.cuda.set_device(dev)
    torch.manual_seed(fx["seed"])
    tr = Trainer(Loader(), cfg)
    _set_gammas(tr.D_s, fx["gamma_s"])
    _set_gammas(tr.D_t, fx["gamma_t"])
    nets = dict(G=tr.G, Ds=tr.D_s, Dt=tr.D_t)
    for key, net in nets.items():
        _fingerprint_ok(net, fx["fp"][key], key)
    pre = {k: {n: p.detach().clone() for n, p in net.named_parameters() if p.requires_grad} for k, net in nets.items()}
    torch.manual_seed(fx["rng_seed"])
    hist = tr.train()
    losses = [float(h[k]) for h in hist for k in ("ds_loss", "dt_loss", "g_loss")]
    noise = fx.get("noise", {})
    print("losses", losses, "reference", fx["losses"], "reference noise", noise)
    ltol = max(1e-3, 3 * max(noise.get("losses", [0.0])))
    assert losses == pytest.approx(fx["losses"], rel=ltol, abs=ltol)
    for key, net in nets.items():
        num = den = 0.0
        for n, p in net.named_parameters():
            if not p.requires_grad or (key == "G" and n.endswith("conv0.module.bias")):
                continue
            d = (p.detach() - pre[key][n]).reshape(-1)
            s = fx["delta"][key][n]
            idx = sample_idx(f"d.{key}.{n}", d.numel(), s["v"].numel()).to(d.device)
            num += float((d[idx].double().cpu() - s["v"].double()).norm() ** 2)
            den += float(s["v"].double().norm() ** 2)
        # beta1 = 0: the first Adam step is ~ lr * sign(g); an element whose tiny gradient changes sign under fp32
        # summation-order noise moves by 2*lr.  The yardstick is the same statistic between two runs of the reference
        # itself (all cores vs half), x3; never tighter than test_train_step_golden's criterion.
        bound = max(0.15, 3 * noise.get("delta", {}).get(key, 0.0))
        print(key, "sampled update rel-L2: %.3f (bound %.3f)" % ((num / den) ** 0.5, bound))
        assert (num / den) ** 0.5 < bound, (key, (num / den) ** 0.5)


# ---------------------------------------------------------------------------------------------------
# lean BPTT, range safety, ONEACC tiles
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 6, 64, 64, 16, 16, 3), (1, 5, 128, 64, 32, 32, 5)])
def test_convgru_lean_bptt_matches_full_state(dev, shape):
    """ops.set_gru_lean(True) keeps h only and recomputes gates / r*h at the start of the backward: same kernels on the same
    operands, so the gradients agree to summation-order noise with the run that kept the whole state."""
    from dvdgan_b200 import ops
    B, T, Cx, Ch, H, W, k = shape
    torch.manual_seed(5)
    x = torch.randn(B, T, Cx, H, W, device=dev)
    ws = [(torch.randn(Ch, Cx + Ch, k, k, device=dev) * 0.05).requires_grad_(True) for _ in range(3)]
    bs = [(torch.randn(Ch, device=dev) * 0.1).requires_grad_(True) for _ in range(3)]
    wgt = torch.randn(B, T, Ch, H, W, device=dev)
    res = []
    for lean in (False, True):
        ops.set_gru_lean(lean)
        try:
            xx = x.clone().requires_grad_(True)
            mem0 = torch.cuda.memory_allocated()
            h = ops.GRULayerFn.apply(xx, None, ws[0], ws[1], ws[2], bs[0], bs[1], bs[2], 0)
            kept = torch.cuda.memory_allocated() - mem0
            assert h.grad_fn.lean is lean                 # the mode really took effect
            (h * wgt).sum().backward()
            res.append((h.detach().clone(), xx.grad.clone(), [w.grad.clone() for w in ws], [b.grad.clone() for b in bs],
                        kept))
            for t in ws + bs:
                t.grad = None
            del h, xx
        finally:
            ops.set_gru_lean(False)
    full, lean = res
    assert rel(lean[0], full[0]) < 1e-6          # (split-K atomics: summation order varies from run to run)
    assert rel(lean[1], full[1]) < 1e-5
    for a, b in zip(lean[2] + lean[3], full[2] + full[3]):
        assert rel(a, b) < 1e-5
    assert lean[4] < 0.45 * full[4], (lean[4], full[4])          # h (+ the input copy) instead of h + gates + r*h


def test_second_backward_through_gru_raises(dev):
    from dvdgan_b200 import ops
    x = torch.randn(1, 3, 8, 8, 8, device=dev, requires_grad=True)
    w = [(torch.randn(8, 16, 3, 3, device=dev) * 0.1).requires_grad_(True) for _ in range(3)]
    b = [torch.zeros(8, device=dev, requires_grad=True) for _ in range(3)]
    h = ops.GRULayerFn.apply(x, None, w[0], w[1], w[2], b[0], b[1], b[2], 0)
    s = h.sum()
    s.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        s.backward()


@pytest.mark.parametrize("scale", [1e5, 3e8])
def test_forward_range_guard(dev, scale):
    """Forward operands ride fp16 (hi, lo) planes: 22 significant bits, which the 1e-3 contract needs over 48 recurrent
    frames, but |x| <= 65504.  Activations of 1e5 / 3e8 are clamped there -- and COUNTED (never silently): the counter
    is what Trainer.check_numerics turns into a switch to bf16 planes, under which the same conv is exact again."""
    from dvdgan_b200 import _C, ops
    torch.manual_seed(3)
    N, Ci, Co, H = 8, 128, 128, 32
    x = torch.randn(N, Ci, H, H, device=dev) * scale
    w = torch.randn(Co, Ci, 3, 3, device=dev) * 0.05
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    wp = ops.pack_weight(w)
    _C.saturation_count()
    y = ops.conv_raw(x, wp, None, Co, (1, 3, 3), x_kind=1)
    assert _C.saturation_count() > 0 and rel(y, ref) > 1e-2            # clamped, and flagged
    y1 = ops.conv_raw(torch.randn(N, Ci, H, H, device=dev) * 100.0, wp, None, Co, (1, 3, 3), x_kind=1)
    assert _C.saturation_count() == 0 and torch.isfinite(y1).all()      # ordinary magnitudes: nothing counted
    _C.set_option("fwd_bf16", 1)
    try:
        y = ops.conv_raw(x, wp, None, Co, (1, 3, 3), x_kind=1)
    finally:
        _C.set_option("fwd_bf16", 0)
    assert _C.saturation_count() == 0 and rel(y, ref) < 2e-5, rel(y, ref)


@pytest.mark.parametrize("scale", [1e5, 1e-7, 1e-20])
def test_gradient_operands_keep_fp32_range(dev, scale):
    """Everything that holds gradients (dgrad / wgrad operands) is split into bf16 planes, which have fp32's exponent
    range: 1e-7 / 1e-20 gradients (below fp16's 6e-8 subnormal floor) and 1e5 ones come out as accurate as O(1) do."""
    from dvdgan_b200 import ops
    torch.manual_seed(3)
    N, Ci, Co, H = 8, 128, 128, 32
    x = torch.randn(N, Ci, H, H, device=dev)
    dy = torch.randn(N, Co, H, H, device=dev) * scale
    w = torch.randn(Co, Ci, 3, 3, device=dev) * 0.05
    gref = torch.nn.grad.conv2d_weight(x.double().cpu(), w.shape, dy.double().cpu(), padding=1)
    g = ops.unpack_wgrad(ops.wgrad_raw(x, dy, (1, 3, 3)), w)
    assert rel(g, gref) < 2e-5, rel(g, gref)
    dref = torch.nn.grad.conv2d_input(x.shape, w.double().cpu(), dy.double().cpu(), padding=1)
    dx = ops.conv_raw(dy, ops.pack_weight(w, transpose=True), None, Ci, (1, 3, 3))       # x_kind = 0: gradient operand
    assert rel(dx, dref) < 2e-5, rel(dx, dref)


def test_trainer_switches_planes_when_activations_saturate(dev, golden):
    from dvdgan_b200 import _C, ops
    from dvdgan_b200.trainer import Trainer
    fx = golden("step.pt")
    torch.cuda.set_device(dev)
    # ch = 8: wide enough (64 channels and up) for the convolutions to run on the tensor-core engine
    tr = Trainer(None, argparse.Namespace(**dict(fx["cfg"], g_chn=8, ds_chn=8, dt_chn=8)))
    _C.saturation_count()
    wave = dict(ops.GRU_WAVEFRONT)
    try:
        with torch.no_grad():            # blow the first ConvGRU's input up past fp16's range
            tr.G.affine_transfrom.weight.mul_(1e7)
        # (whole-clip ConvGRU calls: in this tiny configuration a chunk of frames of the 4x4 stage is under the 128
        # output pixels the tensor-core engine needs, and the fp32 FFMA engine it falls back to has no planes to saturate)
        ops.GRU_WAVEFRONT["enabled"] = 0
        tr.train_step(fx["clips"][0], fx["labels"][0])
        with pytest.warns(RuntimeWarning, match="bf16 operand planes"):
            assert tr.check_numerics() > 0
        assert _C.get_option("fwd_bf16") == 1
        tr.train_step(fx["clips"][1], fx["labels"][1])
        assert tr.check_numerics() == 0
    finally:
        _C.set_option("fwd_bf16", 0)
        ops.GRU_WAVEFRONT.update(wave)


@pytest.mark.parametrize("case", [
    dict(N=64, Ci=256, Co=512, H=32, k=5),        # the per-timestep ConvGRU shape: persistent 256-wide pairs
    dict(N=96, Ci=128, Co=384, H=32, k=5),        # 192-wide tiles
    dict(N=256, Ci=256, Co=128, H=16, k=5),       # 128-wide tiles
])
def test_oneacc_double_buffered_tiles(dev, case):
    """option "oneacc": persistent CTA-pair tiles with ONE accumulator for the three bf16 products and two accumulator
    sets in TMEM; same contract as the two-accumulator kernels (<= 1e-4 per conv), also in accumulate mode."""
    from dvdgan_b200 import _C, ops
    torch.manual_seed(9)
    N, Ci, Co, H, k = case["N"], case["Ci"], case["Co"], case["H"], case["k"]
    x = torch.randn(N, Ci, H, H, device=dev)
    w = torch.randn(Co, Ci, k, k, device=dev) * (1.0 / (Ci * k * k) ** 0.5)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=k // 2)
    wp = ops.pack_weight(w)
    outs = {}
    for one in (0, 1):
        _C.set_option("oneacc", one)
        try:
            y = ops.conv_raw(x, wp, None, Co, (1, k, k), x_kind=1)
            y2 = ops.conv_raw(x, wp, None, Co, (1, k, k), x_kind=1, out=y.clone(), accumulate=1)
        finally:
            _C.set_option("oneacc", 0)
        outs[one] = (rel(y, ref), rel(y2, 2 * ref))
    print("rel-L2 vs fp64 (two accumulators, oneacc):", outs)
    assert max(outs[0]) < 1e-4 and max(outs[1]) < 1e-4


# ---------------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------------
def test_dtype_checks_at_the_boundary(dev):
    from dvdgan_b200 import ops
    x = torch.randn(2, 3, 8, 8, device=dev)
    w = torch.randn(4, 3, 3, 3, device=dev)
    with pytest.raises(TypeError, match="float32"):
        ops.conv(x.double(), w)
    with pytest.raises(TypeError, match="float32"):
        ops.Permute5Fn.apply(torch.zeros(1, 2, 3, 4, 5, device=dev, dtype=torch.uint8), (0, 2, 1, 3, 4))
    with pytest.raises(TypeError, match="int64"):
        ops.EmbeddingFn.apply(torch.zeros(3, device=dev, dtype=torch.int32), torch.randn(4, 6, device=dev))
    with pytest.raises(TypeError, match="int64"):
        ops.GatherFramesFn.apply(torch.randn(1, 4, 3, 8, 8, device=dev), torch.tensor([0, 2], device=dev).int())


def test_out_of_range_class_ids_are_counted_not_dereferenced(dev):
    from dvdgan_b200 import _C, ops
    from dvdgan_b200.Module.Discriminators import SpatialDiscriminator
    _C.index_errors()                                   # clear
    w = torch.randn(4, 6, device=dev)
    y = ops.EmbeddingFn.apply(torch.tensor([0, 3, 4, -1], device=dev), w)
    assert _C.index_errors() == 2
    assert torch.equal(y[2], w[3]) and torch.equal(y[3], w[0])          # clamped into the table
    Ds = SpatialDiscriminator(chn=2, n_class=3).to(dev)
    with torch.no_grad():
        Ds(torch.randn(2, 2, 3, 32, 32, device=dev), torch.tensor([1, 7], device=dev))
    assert _C.index_errors() == 2                       # the two frames of the clip labelled 7
    assert _C.index_errors() == 0


def test_trainer_rejects_bad_labels_and_casts_clips(dev, golden):
    from dvdgan_b200.trainer import Trainer
    fx = golden("step.pt")
    torch.cuda.set_device(dev)
    tr = Trainer(None, argparse.Namespace(**fx["cfg"]))
    with pytest.raises(IndexError):
        tr.train_step(fx["clips"][0], torch.tensor([0, 2]))            # n_class = 2
    out = tr.train_step(fx["clips"][0].double(), fx["labels"][0].int())   # loaders that yield float64 / int32
    assert all(torch.isfinite(v) for v in out.values())
    with pytest.raises(ValueError, match="clips per rank"):
        tr.train_step(fx["clips"][0][:1], fx["labels"][0][:1])


def test_checkpoint_roundtrip_resume_and_sampling(dev, golden, tmp_path):
    """trainer.py:336-343, 375-382, 322-334: save -> load (also from a DataParallel-prefixed file) keeps parameters as
    arena views and reproduces the next step's losses; the sampling path runs G in eval mode on fixed noise."""
    from dvdgan_b200.trainer import Trainer
    fx = golden("step.pt")
    base = dict(fx["cfg"], model_save_path=str(tmp_path / "m"), sample_path=str(tmp_path / "s"), version="v")
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    tr = Trainer(None, argparse.Namespace(**base))
    torch.manual_seed(1)
    tr.train_step(fx["clips"][0], fx["labels"][0])
    tr.save_models(7)
    files = sorted(os.listdir(tr.model_save_path))
    assert files == ["7_Ds.pth", "7_Dt.pth", "7_G.pth"]
    sd = torch.load(os.path.join(tr.model_save_path, "7_G.pth"))
    assert "conv.0.cells.0.update_gate.weight" in sd and "colorize.module.weight_u" in sd      # reference key names
    torch.manual_seed(2)
    want = tr.train_step(fx["clips"][1], fx["labels"][1])
    # resume in a fresh trainer; the G file is rewritten the way nn.DataParallel would have saved it
    torch.save({"module." + k: v for k, v in sd.items()}, os.path.join(tr.model_save_path, "7_G.pth"))
    torch.manual_seed(123)            # different init: everything must come from the files
    tr2 = Trainer(None, argparse.Namespace(**dict(base, pretrained_model=7)))
    lo, hi = tr2.g_optimizer.flat_p.data_ptr(), tr2.g_optimizer.flat_p.data_ptr() + 4 * tr2.g_optimizer.flat_p.numel()
    assert all(lo <= p.data_ptr() < hi for p in tr2.G.parameters() if p.requires_grad)       # still arena views
    # Like the reference, the files hold neither optimizer state nor step count: Adam restarts in the resumed run.  ds_loss
    # and dt_loss of the next step depend on the loaded weights / spectral-norm state only; g_loss is computed after the
    # two discriminator updates, whose Adam moments differ between the runs (lr 5e-5: a small effect).
    torch.manual_seed(2)
    got = tr2.train_step(fx["clips"][1], fx["labels"][1])
    for k, tol in (("ds_loss", 1e-5), ("dt_loss", 1e-5), ("g_loss", 2e-2)):
        assert float(got[k]) == pytest.approx(float(want[k]), rel=tol, abs=tol), k
    with pytest.raises(KeyError, match="unexpected"):
        Trainer._load_sd(tr2.G, dict(sd, bogus=torch.zeros(1)))
    # sampling: one clip per class, de-normalised to [0, 1], G back in train mode afterwards, u/v advanced (Q3)
    u0 = tr2.G.colorize.module.weight_u.detach().clone()
    vids = tr2.sample(step=7)
    assert vids.shape == (base["n_class"] * base["test_batch_size"], base["n_frames"], 3, 64, 64)
    assert float(vids.min()) >= 0.0 and float(vids.max()) <= 1.0 and tr2.G.training
    assert not torch.equal(u0, tr2.G.colorize.module.weight_u)
    assert len(os.listdir(tr2.sample_path)) == base["n_class"] * base["test_batch_size"]


def test_wgan_gp_loss_branch(dev, golden):
    """adv_loss='wgan-gp' (the reference default; its gradient penalty is commented out, Q6): mean(+-x) losses."""
    from dvdgan_b200.trainer import Trainer
    fx = golden("step.pt")
    torch.cuda.set_device(dev)
    tr = Trainer(None, argparse.Namespace(**dict(fx["cfg"], adv_loss="wgan-gp")))
    x = torch.randn(7, device=dev, requires_grad=True)
    y = torch.randn(5, device=dev, requires_grad=True)
    loss = tr.calc_loss(x, True, y, False)
    assert float(loss) == pytest.approx(float(-x.mean() + y.mean()), rel=1e-6)
    loss.backward()
    assert torch.allclose(x.grad, torch.full_like(x, -1 / 7)) and torch.allclose(y.grad, torch.full_like(y, 1 / 5))
    out = tr.train_step(fx["clips"][0], fx["labels"][0])
    assert all(torch.isfinite(v) for v in out.values())


def test_two_devices_in_one_process():
    """The reference's DataParallel drives several devices from one process: per-device kernel attributes
    and pool settings must follow the current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dvdgan_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(64, 256, 32, 32)
    w = torch.randn(256, 256, 5, 5) * 0.01
    outs = []
    for i in (0, 1):
        with torch.cuda.device(i):
            d = torch.device("cuda", i)
            outs.append(ops.conv_raw(x.to(d), ops.pack_weight(w.to(d)), None, 256, (1, 5, 5), x_kind=1).cpu())
            torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------------
# tensor-core attention without the N x N map
# ---------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v):
    s = torch.einsum("bci,bcj->bij", q, k)
    a = torch.softmax(s, dim=-1)
    return torch.einsum("bcj,bij->bci", v, a)


@pytest.mark.parametrize("shape", [
    # B, dq, dv, Nq, Nk
    (3, 16, 128, 256, 256),       # Ds at 64x64 (chn = 32): N = 256 tokens
    (2, 16, 128, 4096, 4096),     # Ds at 256x256 (BASELINE.json configs[3]): the 67 MB-per-frame map that never exists
    (4, 16, 128, 64, 64),         # Dt at 64x64: half a row tile
    (2, 64, 128, 1024, 128),      # pooled 3-D attention (Attention.py:153-185): Nk = N / 8, dq = C / 2
    (2, 4, 32, 200, 72),          # ragged tiles (rows % 128, cols % 64), narrow channels (zero-padded planes)
    (5, 1, 8, 256, 256),          # the golden Ds / Dt fixtures' width (chn = 2)
])
def test_flash_attention_matches_softmax_attention(dev, shape):
    """forward and all three gradients of softmax(Q^T K) V against float64 torch, and against this library's own
    materialised path; logits of O(10) so that the softmax is peaked (the regime where bf16-split logits would hurt)."""
    from dvdgan_b200 import _C, ops
    B, dq, dv, Nq, Nk = shape
    assert _C.lib().dvd_attn_flash_supported(B, dq, dv, Nq, Nk, 1) == 1
    torch.manual_seed(17)
    q = (torch.randn(B, dq, Nq, device=dev) * (3.0 / dq ** 0.5)).requires_grad_(True)
    k = (torch.randn(B, dq, Nk, device=dev) * (3.0 / dq ** 0.5)).requires_grad_(True)
    v = torch.randn(B, dv, Nk, device=dev, requires_grad=True)
    wgt = torch.randn(B, dv, Nq, device=dev)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out = ops.AttnCoreFn.apply(q, k, v, False)
    (out * wgt).sum().backward()
    peak = torch.cuda.max_memory_allocated()
    got = [out.detach(), q.grad.clone(), k.grad.clone(), v.grad.clone()]
    # nothing of size B * Nq * Nk was ever allocated
    if Nq * Nk >= 1 << 24:
        assert peak - base < 0.5 * 4 * B * Nq * Nk, (peak - base, 4 * B * Nq * Nk)
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qd, kd, vd)
    (ref * wgt.double()).sum().backward()
    want = [ref.detach(), qd.grad, kd.grad, vd.grad]
    errs = [rel(a, b) for a, b in zip(got, want)]
    print(shape, "rel-L2 (out, dq, dk, dv):", ["%.1e" % e for e in errs])
    # forward 2e-5; gradients 2e-4 (dQ / dK are sums of dS, whose rows sum to zero, against k / q: the cancellation
    # amplifies the 2^-16 of the two-plane second-stage operands -- measured 1e-5..1e-4; the repo-wide bound is 1e-3)
    assert errs[0] < 2e-5 and max(errs[1:]) < 2e-4, errs
    # the materialised SIMT path gives the same numbers
    for t in (q, k, v):
        t.grad = None
    _C.set_option("flash_attn", 0)
    try:
        out2 = ops.AttnCoreFn.apply(q, k, v, False)
        (out2 * wgt).sum().backward()
    finally:
        _C.set_option("flash_attn", 1)
    assert rel(out2, out) < 2e-5 and rel(q.grad, got[1]) < 2e-4 and rel(k.grad, got[2]) < 2e-4 and rel(v.grad, got[3]) < 2e-4


def test_flash_attention_falls_back_outside_its_coverage(dev):
    from dvdgan_b200 import _C
    lib = _C.lib()
    assert lib.dvd_attn_flash_supported(2, 16, 128, 100, 96, 0) == 0        # Nq % 8
    assert lib.dvd_attn_flash_supported(2, 96, 128, 256, 256, 0) == 0       # dq > 64
    assert lib.dvd_attn_flash_supported(2, 16, 256, 256, 256, 0) == 1 and lib.dvd_attn_flash_supported(2, 16, 256, 256, 256, 1) == 0


@pytest.mark.parametrize("with_h0", [False, True])
def test_fused_bptt_epilogues_match_separate_kernels(dev, with_h0):
    """option "gru_bwd_fused": the gate-gradient math of the BPTT sweep in the epilogues of the two per-step dgrad GEMMs
    (GruEpi modes 3 / 4) against the separate elementwise kernels it replaces -- same operands, same planes."""
    from dvdgan_b200 import _C, ops
    B, T, Cx, Ch, H, W, k = 48, 5, 128, 128, 32, 32, 5
    torch.manual_seed(21)
    x = torch.randn(B, T, Cx, H, W, device=dev) * 0.7
    h0 = torch.randn(B, Ch, H, W, device=dev) * 0.3 if with_h0 else None
    ws = [(torch.randn(Ch, Cx + Ch, k, k, device=dev) * 0.02).requires_grad_(True) for _ in range(3)]
    bs = [(torch.randn(Ch, device=dev) * 0.1).requires_grad_(True) for _ in range(3)]
    wgt = torch.randn(B, T, Ch, H, W, device=dev)
    res = {}
    for fused in (1, 0):
        _C.set_option("gru_bwd_fused", fused)
        try:
            xx = x.clone().requires_grad_(True)
            hh = h0.clone().requires_grad_(True) if with_h0 else None
            h = ops.GRULayerFn.apply(xx, hh, ws[0], ws[1], ws[2], bs[0], bs[1], bs[2], 0)
            (h * wgt).sum().backward()
            res[fused] = [xx.grad.clone()] + ([hh.grad.clone()] if with_h0 else []) + [w.grad.clone() for w in ws] + \
                         [b.grad.clone() for b in bs]
            for t in ws + bs:
                t.grad = None
        finally:
            _C.set_option("gru_bwd_fused", 0)
    for a, b in zip(res[1], res[0]):
        assert rel(a, b) < 2e-5, rel(a, b)


@pytest.mark.parametrize("shape,with_h0,bwd_fused", [
    ((8, 6, 64, 128, 16, 16, 3), False, 0),      # 4 chains of 2 clips (512 rows each)
    ((8, 5, 128, 64, 32, 32, 5), True, 0),       # with an initial state
    ((48, 4, 128, 128, 32, 32, 5), True, 1),     # persistent pair tiles, fused BPTT epilogues
    ((6, 4, 64, 64, 8, 8, 3), False, 0),         # 4 does not divide the batch: falls back to 3 chains
    ((3, 3, 64, 64, 4, 4, 3), False, 0),         # slices too small for the tensor path: one chain
])
def test_convgru_batch_slices_on_helper_streams(dev, shape, with_h0, bwd_fused):
    """option "gru_streams": the time loops of a ConvGRU layer (forward sweep and BPTT) run as independent chains over
    batch slices on library-owned helper streams, forked from / joined to the caller's stream.  The clips of a batch do
    not interact inside a ConvGRU (ConvGRU.py:29-54), so every chain count gives the one-chain result (to the
    summation-order noise of the split-K atomics), with work queued on the caller's stream before and after the call
    ordered against the chains."""
    from dvdgan_b200 import _C, ops
    B, T, Cx, Ch, H, W, k = shape
    torch.manual_seed(31)
    x = torch.randn(B, T, Cx, H, W, device=dev) * 0.7
    h0 = torch.randn(B, Ch, H, W, device=dev) * 0.3 if with_h0 else None
    ws = [(torch.randn(Ch, Cx + Ch, k, k, device=dev) * 0.03).requires_grad_(True) for _ in range(3)]
    bs = [(torch.randn(Ch, device=dev) * 0.1).requires_grad_(True) for _ in range(3)]
    wgt = torch.randn(B, T, Ch, H, W, device=dev)
    default = _C.get_option("gru_streams")
    res = {}
    _C.set_option("gru_bwd_fused", bwd_fused)
    try:
        for n in (1, 2, 4):
            _C.set_option("gru_streams", n)
            # produced on the caller's stream right before the call / consumed right after it: both orderings matter
            xx = (x * 2.0 - x).requires_grad_(True)
            hh = (h0 + 0.0).requires_grad_(True) if with_h0 else None
            h = ops.GRULayerFn.apply(xx, hh, ws[0], ws[1], ws[2], bs[0], bs[1], bs[2], 0)
            out = h * 1.0
            (out * wgt).sum().backward()
            res[n] = [out.detach().clone(), xx.grad.clone()] + ([hh.grad.clone()] if with_h0 else []) + \
                     [w.grad.clone() for w in ws] + [b.grad.clone() for b in bs]
            for t in ws + bs:
                t.grad = None
    finally:
        _C.set_option("gru_streams", default)
        _C.set_option("gru_bwd_fused", 0)
    for n in (2, 4):
        for i, (a, b) in enumerate(zip(res[n], res[1])):
            assert rel(a, b) < 2e-5, (n, i, rel(a, b))


@pytest.mark.parametrize("shape,hidden,ks,bcast,chunk,bwd_fused", [
    ((4, 7, 64, 8, 8), [64, 128, 64], [3, 5, 3], False, 3, 0),        # chunks of 3, 3, 1 frames
    ((2, 6, 64, 4, 4), [64, 128, 64], [3, 5, 3], True, 2, 0),         # stage 0: one input frame fed to every step (Q13)
    ((48, 5, 128, 32, 32), [128, 128], [5, 3], False, 2, 1),          # persistent pair tiles, fused BPTT epilogues
    ((2, 4, 32, 16, 16), [32, 64, 32], [3, 5, 5], False, 1, 0),       # Ch % 64 != 0: planes split after the sweep
])
def test_convgru_stack_wavefront_matches_layer_by_layer(dev, shape, hidden, ks, bcast, chunk, bwd_fused):
    """ops.GRUStackFn: the layers of a ConvGRU stack on one stream each, layer l a chunk of frames behind layer l-1
    (dvd_convgru_layer_{fwd,bwd}_range), against ConvGRU.forward_sequence layer by layer (one whole-clip call per
    layer).  Same kernels on the same operands: equal to summation-order noise, forward and every gradient."""
    from dvdgan_b200 import _C, ops
    from dvdgan_b200.Module.ConvGRU import ConvGRU
    B, T, Cx, H, W = shape
    torch.manual_seed(41)
    net = ConvGRU(Cx, hidden, ks, len(hidden)).to(dev)
    with torch.no_grad():
        for p in net.parameters():          # (orthogonal init / zero biases are a special case)
            p.add_(torch.randn_like(p) * 0.01)
    x = torch.randn(B, Cx, H, W, device=dev) if bcast else torch.randn(B, T, Cx, H, W, device=dev) * 0.7
    wgt = torch.randn(B, T, hidden[-1], H, W, device=dev)
    saved = dict(ops.GRU_WAVEFRONT)
    res = {}
    _C.set_option("gru_bwd_fused", bwd_fused)
    try:
        for wave in (0, 1):
            ops.GRU_WAVEFRONT.update(enabled=wave, chunk=chunk, max_rows=0)
            xx = (x * 2.0 - x).requires_grad_(True)          # produced on the caller's stream right before the call
            h = net.forward_sequence(xx, T_bcast=T if bcast else 0)
            assert (type(h.grad_fn).__name__ == "GRUStackFnBackward") == bool(wave)
            out = h * 1.0
            (out * wgt).sum().backward()
            res[wave] = [out.detach().clone(), xx.grad.clone()] + [p.grad.clone() for p in net.parameters()]
            net.zero_grad()
    finally:
        ops.GRU_WAVEFRONT.clear()
        ops.GRU_WAVEFRONT.update(saved)
        _C.set_option("gru_bwd_fused", 0)
    for i, (a, b) in enumerate(zip(res[1], res[0])):
        assert rel(a, b) < 2e-5, (i, rel(a, b))


def test_generator_with_optional_attention_blocks(dev):
    """Generator(attention=True) wires in the two non-local blocks the reference leaves commented out
    (Generator.py:28-36).  Their gamma is initialised to 0 (Attention.py), so with the same weights the output is the
    plain Generator's, while the backward pass reaches the attention parameters; the default keeps the reference's
    state_dict key set."""
    from dvdgan_b200.Module.Generator import Generator
    torch.manual_seed(3)
    G0 = Generator(in_dim=120, latent_dim=4, n_class=3, ch=8, n_frames=4)
    torch.manual_seed(3)
    G1 = Generator(in_dim=120, latent_dim=4, n_class=3, ch=8, n_frames=4, attention=True)
    extra = set(G1.state_dict()) - set(G0.state_dict())
    assert extra and all(k.startswith(("self_attn.", "sep_attn.")) for k in extra)
    G1.load_state_dict(G0.state_dict(), strict=False)
    z = torch.randn(2, 120)
    cls = torch.tensor([0, 2])
    G0.to(dev), G1.to(dev)
    y0 = G0(z.to(dev), cls.to(dev))
    y1 = G1(z.to(dev), cls.to(dev))
    # (not bit-equal: the split-K convolutions of such a small problem add with atomics, and the default CBN init
    # amplifies that run-to-run summation-order noise ~17x per block)
    assert rel(y1, y0) < 1e-4
    y1.sum().backward()
    assert float(G1.self_attn.gamma.grad.abs()) > 0 and float(G1.sep_attn.model[0].gamma.grad.abs()) > 0
    with torch.no_grad():           # a non-zero gamma changes the output: the blocks are really in the graph
        G1.self_attn.gamma.fill_(0.5)
    assert rel(G1(z.to(dev), cls.to(dev)), y0) > 1e-3
