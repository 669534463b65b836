"""Step helpers with the reference's signatures (reference utils.py:60-63, 77-83)."""
import torch

from . import ops


def sample_k_frames(data, video_length, k_sample):
    """Random sorted k-frame subset.  The permutation is drawn from the default CPU generator exactly like the
    reference (``torch.randperm(video_length)``), so a run with the same seed samples the same frames."""
    frame_idx = torch.randperm(video_length)
    srt, _ = frame_idx[:k_sample].sort()
    return ops.GatherFramesFn.apply(data, srt.to(data.device, non_blocking=True))


def vid_downsample(data):
    """phi: (B,T,C,H,W) -> 2x2 average pool -> (B,C,T,H/2,W/2)."""
    return ops.PhiFn.apply(data)


def denorm(x):
    out = (x + 1) / 2
    return out.clamp_(0, 1)
