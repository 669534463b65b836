"""GPU input pipeline: the reference's per-frame PIL transforms (Dataloader/datasets/ucf101.py:177-199 with the Compose of
main.py:33-56: MultiScaleRandomCrop / MultiScaleCornerCrop -> RandomHorizontalFlip -> ToTensor -> Normalize) as ONE CUDA
launch per batch over decoded uint8 frames, plus a prefetching loader that overlaps the host->device copy of the next
batch with the current training step.

JPEG decoding itself stays on the host side of this boundary (PIL in the reference's worker processes, or nvJPEG through
torchvision.io.decode_jpeg): what moves to the GPU is everything between the decoded frame and the (B, C, T, H, W) float
clip.  At 300 clips/s on a node (14 k frames/s) the reference's PIL crop + resize + ToTensor per frame is what starves the
step; the uint8 source is 4.7 x smaller to copy than it looks (240x320x3 bytes vs 64x64x3 floats is 18.8 : 1 in pixels).

The resize is Pillow's ImagingResample restated: ``resample_coeffs`` follows precompute_coeffs / normalize_coeffs_8bpc of
Pillow's Resample.c operation by operation in Python floats (IEEE doubles, no contraction), so the integer tables -- and
with them the kernel's output -- are bit-identical to PIL.Image.resize(..., BILINEAR).
"""
import ctypes
import math
import random

import numpy as np
import torch

from . import _C

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


_COEFF_CACHE = {}


def resample_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs(inSize, 0, inSize, outSize, BILINEAR) + normalize_coeffs_8bpc:
    -> (bounds int32 [out_size][2] = (first source index, tap count), coeffs int32 [out_size][ksize], ksize)."""
    key = (in_size, out_size)
    if key in _COEFF_CACHE:
        return _COEFF_CACHE[key]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            w = _bilinear((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
            v = k[x] * (1 << PRECISION_BITS)
            coeffs[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    _COEFF_CACHE[key] = (bounds, coeffs, ksize)
    return _COEFF_CACHE[key]


def resize_u8(img, out_h, out_w):
    """Host restatement of the kernel's two passes (numpy, integer math) for an (H, W, C) uint8 image: used by the CPU
    tests to pin ``resample_coeffs`` against PIL without a GPU."""
    H, W, _ = img.shape
    xb, xk, _ = resample_coeffs(W, out_w)
    yb, yk, _ = resample_coeffs(H, out_h)
    half = 1 << (PRECISION_BITS - 1)
    tmp = np.empty((H, out_w, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for ox in range(out_w):
        x0, n = xb[ox]
        acc = (src[:, x0:x0 + n, :] * xk[ox, :n, None].astype(np.int64)).sum(1) + half
        tmp[:, ox, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.empty((out_h, out_w, img.shape[2]), dtype=np.uint8)
    t64 = tmp.astype(np.int64)
    for oy in range(out_h):
        y0, n = yb[oy]
        acc = (t64[y0:y0 + n] * yk[oy, :n, None, None].astype(np.int64)).sum(0) + half
        out[oy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def draw_crop_params(width, height, scales, rng=random):
    """The random draws of Compose.randomize_parameters() for [MultiScaleRandomCrop, RandomHorizontalFlip, ...]
    (spatial_transforms.py:333-336, 267-268) in the reference's order, then PIL.Image.crop's rounding of the box:
    -> ((x0, y0, w, h), flip)."""
    scale = scales[rng.randint(0, len(scales) - 1)]
    tl_x = rng.random()
    tl_y = rng.random()
    p = rng.random()
    crop_size = int(min(width, height) * scale)
    x1 = tl_x * (width - crop_size)
    y1 = tl_y * (height - crop_size)
    bx0, by0, bx1, by1 = (int(round(v)) for v in (x1, y1, x1 + crop_size, y1 + crop_size))
    return (bx0, by0, bx1 - bx0, by1 - by0), int(p < 0.5)


class GpuClipTransform:
    """frames uint8 (B, T, Hs, Ws, 3) on the device + per-clip crop boxes / flips -> float32 (B, 3, T, size, size)."""

    def __init__(self, sample_size, norm_value=255.0, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
        self.size = int(sample_size)
        self.norm_value = float(norm_value)
        self.mean = (ctypes.c_float * 3)(*mean)
        self.std = (ctypes.c_float * 3)(*std)

    def tables(self, boxes):
        """int32 host tables for a batch of crop boxes (cached per crop size)."""
        B, S = len(boxes), self.size
        per = [(resample_coeffs(w, S), resample_coeffs(h, S)) for (_, _, w, h) in boxes]
        xks = max(p[0][2] for p in per)
        yks = max(p[1][2] for p in per)
        xb = np.zeros((B, S, 2), np.int32); xk = np.zeros((B, S, xks), np.int32)
        yb = np.zeros((B, S, 2), np.int32); yk = np.zeros((B, S, yks), np.int32)
        for i, ((bx, kx, nx), (by, ky, ny)) in enumerate(per):
            xb[i], yb[i] = bx, by
            xk[i, :, :nx], yk[i, :, :ny] = kx, ky
        return xb, xk, xks, yb, yk, yks

    def __call__(self, frames, boxes, flips):
        if not frames.is_cuda or frames.dtype is not torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3:
            raise TypeError("frames must be a CUDA uint8 tensor (B, T, Hs, Ws, 3)")
        frames = frames.contiguous()
        B, T, Hs, Ws, _ = frames.shape
        for (x0, y0, w, h) in boxes:
            if x0 < 0 or y0 < 0 or w <= 0 or h <= 0 or x0 + w > Ws or y0 + h > Hs:
                raise ValueError(f"crop box {(x0, y0, w, h)} outside the {Ws}x{Hs} frame")
        xb, xk, xks, yb, yk, yks = self.tables(boxes)
        dev = frames.device

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        d_box = up(np.asarray(boxes, np.int32).reshape(B, 4))
        d_flip = up(np.asarray(flips, np.int32).reshape(B))
        d = [up(a) for a in (xb, xk, yb, yk)]
        out = torch.empty((B, 3, T, self.size, self.size), device=dev, dtype=torch.float32)
        _C.call("dvd_clip_transform", frames.data_ptr(), B, T, Hs, Ws, d_box.data_ptr(), d_flip.data_ptr(),
                d[0].data_ptr(), d[1].data_ptr(), xks, d[2].data_ptr(), d[3].data_ptr(), yks, self.size, self.size,
                self.norm_value, self.mean, self.std, out.data_ptr())
        return out


class PrefetchLoader:
    """Wraps an iterable of host batches ``(frames uint8 (B,T,Hs,Ws,3), labels, boxes, flips)`` -- what worker processes
    produce after decoding and drawing the crop parameters -- and yields ``(clips float32 (B,3,T,S,S), labels int64)`` on
    the device.  The copy + transform of batch i+1 run on a side stream while the caller trains on batch i."""

    def __init__(self, batches, transform, device=None):
        self.batches = batches
        self.transform = transform
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.batches)

    def _stage(self, item):
        frames, labels, boxes, flips = item
        with torch.cuda.stream(self.stream):
            if not frames.is_pinned():
                frames = frames.pin_memory()
            d_frames = frames.to(self.device, non_blocking=True)
            clips = self.transform(d_frames, boxes, flips)
            d_labels = torch.as_tensor(labels).to(self.device, dtype=torch.int64, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return clips, d_labels, done, d_frames

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            clips, labels, done, keep = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(done)
            clips.record_stream(torch.cuda.current_stream(self.device))
            labels.record_stream(torch.cuda.current_stream(self.device))
            del keep
            yield clips, labels
