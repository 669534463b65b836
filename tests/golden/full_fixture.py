"""Index sampler shared by tests/golden/make_golden_full.py (which writes the samples) and the GPU tests (which read
them): k indices into a flat tensor of n elements as a pure function of (name, n, k)."""
import zlib

import torch


def sample_idx(name, n, k):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    if n <= k:
        return torch.arange(n)
    return torch.randint(n, (k,), generator=g)
