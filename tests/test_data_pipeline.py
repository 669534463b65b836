"""Input pipeline.  CPU: the host restatement of Pillow's resample coefficients against Pillow itself, and
the reference's random-parameter draws.  GPU: dvd_clip_transform against the PIL oracle, bit for bit."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("in_hw,out", [((240, 320), 64), ((200, 200), 64), ((96, 131), 64), ((64, 64), 64),
                                       ((50, 70), 128), ((333, 777), 112)])
def test_resample_coeffs_reproduce_pillow(in_hw, out):
    """dvdgan_b200/data.py restates Pillow's precompute_coeffs + 8-bit two-pass resample: bit-exact, down- and upscaling"""
    from PIL import Image
    from dvdgan_b200.data import resize_u8
    rng = np.random.default_rng(in_hw[0] * 1000 + out)
    img = rng.integers(0, 256, size=(*in_hw, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((out, out), Image.BILINEAR))
    got = resize_u8(img, out, out)
    assert np.array_equal(got, want)


def test_crop_parameter_draws_follow_the_reference_order():
    """same `random` seed -> same scale / corner / flip as Compose.randomize_parameters() of the reference's transforms"""
    from dvdgan_b200.data import draw_crop_params
    scales = [1.0, 1.0 / 2 ** 0.25, 1.0 / 2 ** 0.5]
    random.seed(123)
    box, flip = draw_crop_params(320, 240, scales)
    random.seed(123)
    scale = scales[random.randint(0, len(scales) - 1)]
    tl_x, tl_y, p = random.random(), random.random(), random.random()
    cs = int(240 * scale)
    assert box == (int(round(tl_x * (320 - cs))), int(round(tl_y * (240 - cs))), cs, cs) and flip == int(p < 0.5)


@pytest.mark.gpu
def test_gpu_clip_transform_matches_pil_bit_for_bit():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, ROOT)
    from oracle.clip_transform_oracle import transform_clip
    from dvdgan_b200.data import GpuClipTransform, PrefetchLoader, draw_crop_params
    dev = torch.device("cuda:0")
    B, T, Hs, Ws, S = 5, 6, 240, 320, 64
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, size=(B, T, Hs, Ws, 3), dtype=np.uint8)
    scales = [1.0, 1.0 / 2 ** 0.25, 1.0 / 2 ** 0.5, 64 / 240]           # the last: crop == output size, no resampling
    random.seed(5)
    boxes, flips = zip(*[draw_crop_params(Ws, Hs, scales) for _ in range(B)])
    boxes = list(boxes)
    boxes[-1] = (17, 33, 64, 64)
    tf = GpuClipTransform(S)
    got = tf(torch.from_numpy(frames).to(dev), boxes, list(flips)).cpu()
    assert got.shape == (B, 3, T, S, S)
    for b in range(B):
        x0, y0, w, h = boxes[b]
        want = transform_clip(frames[b], (x0, y0, x0 + w, y0 + h), flips[b], S)
        assert torch.equal(got[b], want), (b, float((got[b] - want).abs().max()))
    assert float(got.min()) >= -1.0 and float(got.max()) <= 1.0
    # the prefetching loader yields the same clips, in order, with int64 labels on the device
    batches = [(torch.from_numpy(frames), torch.arange(B), boxes, list(flips)) for _ in range(3)]
    seen = 0
    for clips, labels in PrefetchLoader(batches, tf, dev):
        assert clips.is_cuda and labels.dtype is torch.int64 and torch.equal(clips.cpu(), got)
        seen += 1
    assert seen == 3
    with pytest.raises(ValueError):
        tf(torch.from_numpy(frames).to(dev), [(300, 0, 64, 64)] * B, [0] * B)
