"""Alternating D_s -> D_t -> G training step with the reference Trainer's interface (reference
trainer.py:14-391), on the dvdgan_b200 CUDA kernels.

Differences from the reference that do not change the numbers:
  * one process per GPU; when torch.distributed is initialised the batch is sharded across ranks and the flat
    fp32 gradient arena of the network being updated is all-reduced (NCCL, sum) before its Adam step -- the
    reference's nn.DataParallel (trainer.py:353-359) sums replica gradients the same way; BatchNorm statistics
    stay per replica in both;
  * parameters / Adam moments / gradients of each network live in flat arenas: one fused Adam launch and one
    collective per network instead of one per tensor;
  * during the G update the discriminators' parameters do not require grad (the reference computes those
    gradients and throws them away at the next reset_grad, trainer.py:384-387).
RNG draws on the default CPU generator happen in the reference's order.
"""
import contextlib
import os
import time

import torch
import torch.distributed as dist

from . import ops
from .Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
from .Module.Generator import Generator
from .utils import sample_k_frames, vid_downsample


class FlatAdam:
    """torch.optim.Adam(lr, betas, eps=1e-8) over every requires_grad parameter of ``net`` (trainer.py:136-141),
    with params, grads and moments in flat fp32 arenas."""

    def __init__(self, net, lr, betas, eps=1e-8):
        self.params = [p for p in net.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        self.slices = []
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view(p.shape)      # parameters become views of the arena
            self.slices.append((off, k))
            off += k
        self.numel = n
        self.base_lr = lr
        self.lr = lr
        self.betas = betas
        self.eps = eps
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """Copy every parameter gradient into the flat arena (missing grads count as zero): one multi-tensor launch
        per 96 parameters instead of one copy per parameter."""
        import ctypes
        g = self.flat_g
        n = len(self.params)
        srcs = (ctypes.c_void_p * n)()
        keep = []
        for i, p in enumerate(self.params):
            if p.grad is None:
                srcs[i] = None
            else:
                src = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(src)
                srcs[i] = src.data_ptr()
        if not hasattr(self, "_offs"):
            self._offs = (ctypes.c_int64 * n)(*[o for o, _ in self.slices])
            self._cnts = (ctypes.c_int64 * n)(*[k for _, k in self.slices])
        self._gather(srcs, self._offs, self._cnts, n, g)
        return g

    @staticmethod
    def _gather(srcs, offs, cnts, n, flat):
        ops.call("dvd_gather_flat", srcs, offs, cnts, n, ops.ptr(flat))

    @staticmethod
    def _copy(src, flat, off, k):
        ops.call("dvd_axpby", ops.ptr(src), 1.0, 0.0, k, flat.data_ptr() + 4 * off)

    @staticmethod
    def _adam(p, g, m, v, lr, b1, b2, eps, t, scale):
        ops.adam_step(p, g, m, v, lr, b1, b2, eps, t, scale)

    def step(self, world_size=1):
        """gather -> (sum all-reduce across ranks) -> fused Adam with the 1/world_size average folded in."""
        g = self.gather_grads()
        if world_size > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        self.t += 1
        self._adam(self.flat_p, g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                   1.0 / world_size)


def shard_batch(global_batch, world_size, rank):
    """Data-parallel sharding of a step's clips: equal contiguous shards, one per rank."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return slice(rank * per, (rank + 1) * per)


def _lr_at(kind, base, t, decay):
    """Learning rate after t scheduler steps (trainer.py:142-176; stepped once per optimizer step)."""
    if kind == 'const':
        return base
    if kind == 'step':
        return base * 0.98 ** (t // 500)
    if kind == 'exp':
        return base * 0.9999 ** t
    if kind == 'multi':
        return base * 0.3 ** ((t >= 10000) + (t >= 30000))
    raise NotImplementedError("lr_schr='reduce' (ReduceLROnPlateau(verbose=...)) raises in the reference itself on "
                              "torch >= 2.4; use const/step/exp/multi")


@contextlib.contextmanager
def _frozen(*nets):
    params = [p for n in nets for p in n.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


class Trainer(object):
    """Same constructor contract as the reference: ``Trainer(data_loader, config)``; ``config`` carries the
    reference's argparse fields (parameter.py:6-79).  ``latent_dim`` (not a reference flag) may be added for
    128x128 / 256x256 clips."""

    def __init__(self, data_loader, config):
        self.data_loader = data_loader
        c = config
        self.adv_loss = c.adv_loss
        self.z_dim, self.g_chn, self.ds_chn, self.dt_chn = c.z_dim, c.g_chn, c.ds_chn, c.dt_chn
        self.n_frames, self.k_sample, self.n_class = c.n_frames, c.k_sample, c.n_class
        self.batch_size, self.d_iters = c.batch_size, c.d_iters
        self.g_lr, self.d_lr, self.beta1, self.beta2 = c.g_lr, c.d_lr, c.beta1, c.beta2
        self.lr_schr, self.lr_decay = c.lr_schr, getattr(c, "lr_decay", 0.9999)
        self.latent_dim = getattr(c, "latent_dim", 4)
        self.total_epoch = getattr(c, "total_epoch", 1)
        self.log_epoch = getattr(c, "log_epoch", 1)
        self.sample_epoch = getattr(c, "sample_epoch", 10 ** 9)
        self.model_save_epoch = getattr(c, "model_save_epoch", 10 ** 9)
        self.test_batch_size = getattr(c, "test_batch_size", 1)
        self.pretrained_model = getattr(c, "pretrained_model", None)
        version = getattr(c, "version", "")
        self.model_save_path = os.path.join(getattr(c, "model_save_path", "./models"), version)
        if self.adv_loss not in ("hinge", "wgan-gp"):
            raise ValueError("adv_loss must be 'hinge' or 'wgan-gp'")
        if not torch.cuda.is_available():
            raise RuntimeError("dvdgan_b200.Trainer needs a CUDA device: there is no CPU fallback")
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world_size = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.build_model()
        if self.pretrained_model:
            self.load_pretrained_model()

    # ------------------------------------------------------------------ model / optimizers
    def build_model(self):
        self.G = Generator(self.z_dim, latent_dim=self.latent_dim, n_class=self.n_class, ch=self.g_chn,
                           n_frames=self.n_frames).to(self.device)
        self.D_s = SpatialDiscriminator(chn=self.ds_chn, n_class=self.n_class).to(self.device)
        self.D_t = TemporalDiscriminator(chn=self.dt_chn, n_class=self.n_class).to(self.device)
        if self.distributed:            # replicas start identical (DataParallel re-broadcasts every forward)
            for net in (self.G, self.D_s, self.D_t):
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t.data, src=0)
        self.select_opt_schr()

    def select_opt_schr(self):
        betas = (self.beta1, self.beta2)
        self.g_optimizer = FlatAdam(self.G, self.g_lr, betas)
        self.ds_optimizer = FlatAdam(self.D_s, self.d_lr, betas)
        self.dt_optimizer = FlatAdam(self.D_t, self.d_lr, betas)
        _lr_at(self.lr_schr, 1.0, 0, self.lr_decay)     # validates lr_schr

    def _sched_step(self, opt):
        opt.lr = _lr_at(self.lr_schr, opt.base_lr, opt.t, self.lr_decay)

    def reset_grad(self):
        self.ds_optimizer.zero_grad()
        self.dt_optimizer.zero_grad()
        self.g_optimizer.zero_grad()

    def label_sample(self):
        label = torch.randint(low=0, high=self.n_class, size=(self.batch_size,))
        return label.to(self.device)

    def calc_loss(self, x, real_flag, y=None, y_real_flag=None):
        """trainer.py:114-121; with ``y`` the two-term sum loss(x) + loss(y) in one op."""
        hinge = self.adv_loss == 'hinge'
        sx = -1.0 if real_flag is True else 1.0
        sy = -1.0 if y_real_flag is True else 1.0
        return ops.GanLossFn.apply(hinge, sx, x, sy, y)

    # ------------------------------------------------------------------ one step (trainer.py:229-307)
    def train_step(self, real_videos, real_labels):
        """real_videos (B,C,T,H,W) on the device (as the loader yields them), real_labels (B,) int64."""
        real_videos = ops.Permute5Fn.apply(real_videos, (0, 2, 1, 3, 4))       # -> B,T,C,H,W
        for _ in range(self.d_iters):
            real_s = sample_k_frames(real_videos, self.n_frames, self.k_sample)
            z = torch.randn(self.batch_size, self.z_dim).to(self.device)
            z_class = self.label_sample()
            fake_videos = self.G(z, z_class)
            fv_s, fv_t = ops.fork(fake_videos, 2)
            # ---- D_s
            fake_s = sample_k_frames(fv_s, self.n_frames, self.k_sample)
            ds_real = self.D_s(real_s, real_labels)
            ds_fake = self.D_s(fake_s.detach(), z_class)
            ds_loss = self.calc_loss(ds_real, True, ds_fake, False)
            self.reset_grad()
            ds_loss.backward()
            self.ds_optimizer.step(self.world_size)
            self._sched_step(self.ds_optimizer)
            # ---- D_t
            real_d = vid_downsample(real_videos)
            fake_d = vid_downsample(fv_t)
            dt_real = self.D_t(real_d, real_labels)
            dt_fake = self.D_t(fake_d.detach(), z_class)
            dt_loss = self.calc_loss(dt_real, True, dt_fake, False)
            self.reset_grad()
            dt_loss.backward()
            self.dt_optimizer.step(self.world_size)
            self._sched_step(self.dt_optimizer)
        # ---- G (uses the already-updated discriminators; their u/v advance a third time, Q10)
        with _frozen(self.D_s, self.D_t):
            g_s = self.D_s(fake_s, z_class)
            g_t = self.D_t(fake_d, z_class)
            g_loss = self.calc_loss(g_s, True, g_t, True)
            self.reset_grad()
            g_loss.backward()
        self.g_optimizer.step(self.world_size)
        self._sched_step(self.g_optimizer)
        return {"ds_loss": ds_loss.detach(), "dt_loss": dt_loss.detach(), "g_loss": g_loss.detach()}

    # ------------------------------------------------------------------ loop (trainer.py:189-343)
    def epoch2step(self):
        self.epoch = 0
        step_per_epoch = len(self.data_loader)
        self.total_step = self.total_epoch * step_per_epoch
        self.log_step = self.log_epoch * step_per_epoch
        self.model_save_step = self.model_save_epoch * step_per_epoch

    def train(self):
        data_iter = iter(self.data_loader)
        self.epoch2step()
        # consumed to keep the CPU RNG stream aligned with the reference (trainer.py:195)
        self.fixed_z = torch.randn(self.test_batch_size * self.n_class, self.z_dim).to(self.device)
        start = self.pretrained_model + 1 if self.pretrained_model else 1
        start_time = time.time()
        self.D_s.train(); self.D_t.train(); self.G.train()
        history = []
        for step in range(start, self.total_step + 1):
            try:
                real_videos, real_labels = next(data_iter)
            except StopIteration:
                data_iter = iter(self.data_loader)
                real_videos, real_labels = next(data_iter)
                self.epoch += 1
            real_videos = real_videos.to(self.device, non_blocking=True)
            real_labels = real_labels.to(self.device, non_blocking=True)
            out = self.train_step(real_videos, real_labels)
            history.append(out)
            if step % self.log_step == 0 and self.rank == 0:
                elapsed = time.time() - start_time
                start_time = time.time()
                print("Epoch: [%d/%d], Step: [%d/%d], time: %.1fs, ds_loss: %.4f, dt_loss: %.4f, g_loss: %.4f, lr: %.2e"
                      % (self.epoch, self.total_epoch, step, self.total_step, elapsed, float(out["ds_loss"]),
                         float(out["dt_loss"]), float(out["g_loss"]), self.g_optimizer.lr))
            if step % self.model_save_step == 0 and self.rank == 0:
                self.save_models(step)
        return history

    # ------------------------------------------------------------------ checkpoints (trainer.py:336-343,375-382)
    def save_models(self, step):
        os.makedirs(self.model_save_path, exist_ok=True)
        for net, tag in ((self.G, "G"), (self.D_s, "Ds"), (self.D_t, "Dt")):
            torch.save(net.state_dict(), os.path.join(self.model_save_path, '{}_{}.pth'.format(step, tag)))

    @staticmethod
    def _load(net, path):
        Trainer._load_sd(net, torch.load(path, map_location="cpu"), path)

    @staticmethod
    def _load_sd(net, sd, path="<state_dict>"):
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}  # DataParallel files
        own = net.state_dict()
        for k, v in sd.items():
            own[k].copy_(v)          # in place: parameters stay views of the flat arena
        missing = set(own) - set(sd)
        if missing:
            raise KeyError(f"missing keys in {path}: {sorted(missing)[:5]}")

    def load_pretrained_model(self):
        for net, tag in ((self.G, "G"), (self.D_s, "Ds"), (self.D_t, "Dt")):
            self._load(net, os.path.join(self.model_save_path, '{}_{}.pth'.format(self.pretrained_model, tag)))
