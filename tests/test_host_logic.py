"""CPU tests of host-side logic added in round 2: LR schedules vs the torch schedulers the reference uses, the option
table of the C ABI, the ConvGRU state estimate that drives the lean-BPTT decision, the fixture index sampler."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind,make", [
    ("const", lambda o: torch.optim.lr_scheduler.StepLR(o, step_size=10000, gamma=1)),
    ("step", lambda o: torch.optim.lr_scheduler.StepLR(o, step_size=500, gamma=0.98)),
    ("exp", lambda o: torch.optim.lr_scheduler.ExponentialLR(o, gamma=0.9999)),
    ("multi", lambda o: torch.optim.lr_scheduler.MultiStepLR(o, [10000, 30000], gamma=0.3)),
])
def test_lr_schedules_match_reference_schedulers(kind, make):
    """trainer.py:142-158: the schedulers are stepped once per optimizer step; _lr_at(t) is the closed form."""
    from dvdgan_b200.trainer import _lr_at
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=5e-5)
    sch = make(opt)
    checkpoints = {0, 1, 499, 500, 501, 999, 1000, 9999, 10000, 10001, 29999, 30000, 30500}
    for t in range(30501):
        if t in checkpoints:
            assert opt.param_groups[0]["lr"] == pytest.approx(_lr_at(kind, 5e-5, t, 0.9999), rel=1e-9), (kind, t)
        opt.step()
        sch.step()


def test_lr_reduce_raises_like_the_reference():
    from dvdgan_b200.trainer import _lr_at
    with pytest.raises(NotImplementedError):
        _lr_at("reduce", 1.0, 0, 0.9)


def test_options_table():
    from dvdgan_b200 import _C
    assert _C.get_option("pair") == 1 and _C.get_option("simt_only") == 0
    _C.set_option("pair", 0)
    assert _C.get_option("pair") == 0
    _C.set_option("pair", 1)
    with pytest.raises(ValueError, match="unknown option"):
        _C.set_option("no_such_switch", 1)


def test_gru_state_estimate_drives_lean_mode():
    from dvdgan_b200.ops import gru_state_bytes
    full2 = gru_state_bytes(64, 48, 32, 4, lean=False)
    assert 52e9 < full2 < 55e9                       # 53 GB of ConvGRU state at config 2
    assert gru_state_bytes(64, 48, 32, 4, lean=True) * 5 == full2
    # config 3 (32 clips of 48 x 128x128 per GPU) does not fit 180 GB with the full state; lean does
    assert gru_state_bytes(32, 48, 32, 8, lean=False) > 0.35 * 180e9
    assert gru_state_bytes(32, 48, 32, 8, lean=True) < 25e9
    # the policy: layers that free the most bytes per recomputed FLOP go lean first, until the kept state fits
    from dvdgan_b200.ops import gru_lean_policy, gru_layer_state_bytes, gru_layers
    budget = 0.30 * 190e9
    pol, kept = gru_lean_policy(32, 48, 32, 8, budget)
    layers, sizes = gru_layers(32, 8), gru_layer_state_bytes(32, 48, 32, 8)
    assert kept <= budget and kept == sum(s // 5 if l in pol else s for l, s in zip(layers, sizes))
    flops = lambda l: (l[0] + l[1]) * l[1] * l[4] ** 2 * l[2] * l[3]
    # cheaper than recomputing everything, and the expensive 5x5 middle layer of the 32x32 stage keeps its state
    assert sum(flops(l) for l in pol) < 0.6 * sum(flops(l) for l in layers) and (256, 512, 32, 32, 5) not in pol
    assert gru_lean_policy(64, 48, 32, 4, budget)[0] is False           # config 2: nothing recomputed
    assert gru_lean_policy(16, 128, 32, 4, budget)[0] is False          # config 5
    assert gru_lean_policy(32, 48, 32, 8, 1e9)[0] is True               # hopeless budget: everything lean


def test_fixture_sampler_is_deterministic():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from full_fixture import sample_idx
    a, b = sample_idx("out", 10 ** 6, 1000), sample_idx("out", 10 ** 6, 1000)
    assert torch.equal(a, b) and a.max() < 10 ** 6 and not torch.equal(a, sample_idx("pre_tanh", 10 ** 6, 1000))
    assert torch.equal(sample_idx("x", 10, 1000), torch.arange(10))


def test_convgru_stack_wavefront_policy():
    """the wavefront needs every layer's full state: a lean layer (its gates are recomputed per layer) or a stage above
    ``max_rows`` sends the stack down the layer-by-layer path."""
    from dvdgan_b200 import ops
    sigs = [(256, 256, 8, 8, 3), (256, 512, 8, 8, 5), (512, 256, 8, 8, 3)]
    saved = dict(ops.GRU_WAVEFRONT)
    try:
        ops.GRU_WAVEFRONT.update(enabled=1, chunk=8, max_rows=0)
        assert ops.gru_wavefront_chunk(64, 48, sigs) == 8
        assert ops.gru_wavefront_chunk(64, 12, sigs) == 4 and ops.gru_wavefront_chunk(64, 1, sigs) == 0
        assert ops.gru_wavefront_chunk(64, 48, sigs[:1]) == 0
        ops.set_gru_lean({sigs[1]})
        assert ops.gru_wavefront_chunk(64, 48, sigs) == 0
        ops.set_gru_lean(False)
        ops.GRU_WAVEFRONT.update(max_rows=1024)
        assert ops.gru_wavefront_chunk(64, 48, sigs) == 0 and ops.gru_wavefront_chunk(16, 48, sigs) == 8
    finally:
        ops.set_gru_lean(False)
        ops.GRU_WAVEFRONT.clear()
        ops.GRU_WAVEFRONT.update(saved)
