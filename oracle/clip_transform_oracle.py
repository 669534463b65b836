"""TEST INFRASTRUCTURE ONLY (imported by tests/ only; never by dvdgan_b200/): the reference's per-frame spatial transform,
restated with PIL exactly as Dataloader/transform/spatial_transforms.py composes it (main.py:33-56):

    MultiScaleRandomCrop.__call__ (:347-362)  img.crop((x1, y1, x2, y2)).resize((size, size), BILINEAR)
    RandomHorizontalFlip.__call__ (:256-265)  img.transpose(FLIP_LEFT_RIGHT) when p < 0.5
    ToTensor.__call__ (:47-58)                torch.from_numpy(HWC).permute(2,0,1).float().div(norm_value)
    Normalize.__call__ (:108-118)             t.sub_(m).div_(s) per channel
    UCF101.__getitem__ (ucf101.py:192-193)    torch.stack(clip, 0).permute(1, 0, 2, 3)      -> (C, T, H, W)

Pinned by construction: it calls Pillow itself, which is the third-party code the reference delegates the arithmetic to
(Pillow 12.2 in this image)."""
import numpy as np
import torch
from PIL import Image


def transform_clip(frames_u8, crop_xyxy, flip, size, norm_value=255.0, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """frames_u8 (T, H, W, 3) uint8 numpy; crop_xyxy = the FLOAT box the reference hands to img.crop -> (3, T, size, size)."""
    out = []
    for f in frames_u8:
        img = Image.fromarray(f).crop(crop_xyxy).resize((size, size), Image.BILINEAR)
        if flip:
            img = img.transpose(Image.FLIP_LEFT_RIGHT)
        t = torch.from_numpy(np.asarray(img).transpose((2, 0, 1)).copy()).float().div(norm_value)
        for ch, m, s in zip(t, mean, std):
            ch.sub_(m).div_(s)
        out.append(t)
    return torch.stack(out, 0).permute(1, 0, 2, 3).contiguous()
