"""dvdgan_b200: B200-native (sm_100a) DVD-GAN training hot path.

Hand-written CUDA kernels behind a C ABI (include/dvdgan_b200.h, dvdgan_b200/csrc), bound as
torch.autograd.Functions (ops.py) under drop-in nn.Modules that keep the reference's
forward() signatures and state_dict keys (Module/*.py, trainer.py, utils.py).
"""
__version__ = "0.1.0"
