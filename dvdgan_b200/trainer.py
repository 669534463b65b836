"""Alternating D_s -> D_t -> G training step with the reference Trainer's interface (reference
trainer.py:14-391), on the dvdgan_b200 CUDA kernels.

Differences from the reference that do not change the numbers:
  * one process per GPU.  ``config.batch_size`` is the GLOBAL batch, as under the reference's nn.DataParallel
    (trainer.py:353-359): when torch.distributed is initialised every rank draws z / labels / frame subsets for the
    global batch from the SAME CPU generator stream (rank 0's state is broadcast at construction) and keeps its
    contiguous shard (``shard_batch``); the flat fp32 gradient arena of the network being updated is summed across
    ranks (NCCL) before its Adam step, the 1/N folded into the Adam kernel.  BatchNorm statistics stay per
    replica in both;
  * ``config.shard_optimizer`` (default off): reduce-scatter of the gradient arena -> Adam on this rank's 1/N slice
    of (params, m, v) -> all-gather of the parameter arena, instead of all-reduce + N identical full Adam steps;
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

from . import _C, ops
from .Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
from .Module.Generator import Generator
from .utils import denorm, sample_k_frames, vid_downsample


class FlatAdam:
    """torch.optim.Adam(lr, betas, eps=1e-8) over every requires_grad parameter of ``net`` (trainer.py:136-141),
    with params, grads and moments in flat fp32 arenas.

    ``shard=(world, rank)``: ZeRO-1 style -- the arena is padded to a multiple of ``world``; ``step`` reduce-scatters
    the gradient arena, runs Adam on this rank's slice only (m / v exist only for the slice) and all-gathers the
    parameter arena in place."""

    def __init__(self, net, lr, betas, eps=1e-8, shard=None):
        self.params = [p for p in net.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.shard = shard if (shard and shard[0] > 1) else None
        world = self.shard[0] if self.shard else 1
        n_pad = (n + world - 1) // world * world
        self.flat_p = torch.zeros(n_pad, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n_pad, device=dev, dtype=torch.float32)
        n_state = n_pad // world
        self.m = torch.zeros(n_state, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n_state, device=dev, dtype=torch.float32)
        if self.shard:
            lo = self.shard[1] * n_state
            self.p_shard = self.flat_p[lo:lo + n_state]
            self.g_shard = torch.zeros(n_state, device=dev, dtype=torch.float32)
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
        self._pending = []
        self._fired = [0] * len(getattr(self, "_buckets", ()))

    # -- overlap of the gradient all-reduce with the backward pass -----------------------------------------------------
    def enable_overlap(self, net):
        """Split the arena into one bucket per top-level child of ``net`` (for G: embedding, affine, the 12 ConvGRU /
        GResBlock stages, colorize -- contiguous in the arena because parameters() walks the module tree in order) and
        start each bucket's all-reduce as soon as the backward pass has produced all of its gradients: the last
        stages' 60 % of the 547 MB arena is reduced while the earlier stages are still back-propagating.  step() then
        only waits.  Buckets whose gradients never show up (unused parameters) are handled by step() as before."""
        if self.shard:
            return                      # the reduce-scatter path reduces the arena in one piece
        index = {id(p): i for i, p in enumerate(self.params)}
        self._buckets = []
        for child in net.children():
            idx = sorted(index[id(p)] for p in child.parameters() if id(p) in index)
            if idx:
                assert idx == list(range(idx[0], idx[-1] + 1)), "bucket parameters are contiguous in the arena"
                self._buckets.append((idx[0], idx[-1] + 1))
        covered = sum(b - a for a, b in self._buckets)
        assert covered == len(self.params), "every parameter belongs to exactly one top-level child"
        self._bucket_of = {}
        for bi, (a, b) in enumerate(self._buckets):
            for i in range(a, b):
                self._bucket_of[i] = bi
        self._fired = [0] * len(self._buckets)
        self._pending = []
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(_param):
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
                return
            bi = self._bucket_of[i]
            self._fired[bi] += 1
            a, b = self._buckets[bi]
            if self._fired[bi] == b - a:
                self._launch_bucket(bi)
        return hook

    def _launch_bucket(self, bi):
        a, b = self._buckets[bi]
        self._gather_range(a, b)
        lo = self.slices[a][0]
        hi = self.slices[b - 1][0] + self.slices[b - 1][1]
        work = dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
        self._pending.append((bi, work))

    def _gather_range(self, a, b):
        import ctypes
        n = b - a
        srcs = (ctypes.c_void_p * n)()
        keep = []
        for j in range(n):
            g = self.params[a + j].grad
            if g is None:
                srcs[j] = None
            else:
                g = g if g.is_contiguous() else g.contiguous()
                keep.append(g)
                srcs[j] = g.data_ptr()
        offs = (ctypes.c_int64 * n)(*[self.slices[a + j][0] for j in range(n)])
        cnts = (ctypes.c_int64 * n)(*[self.slices[a + j][1] for j in range(n)])
        self._gather(srcs, offs, cnts, n, self.flat_g)

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
    def _adam(p, g, m, v, lr, b1, b2, eps, t, scale):
        ops.adam_step(p, g, m, v, lr, b1, b2, eps, t, scale)

    @staticmethod
    def _reduce_scatter(out, full):
        """sum-reduce-scatter of a flat arena; gloo (CPU tests of this host logic) has no reduce_scatter."""
        if dist.get_backend() == "gloo":
            dist.all_reduce(full, op=dist.ReduceOp.SUM)
            n = out.numel()
            out.copy_(full[dist.get_rank() * n:(dist.get_rank() + 1) * n])
        else:
            dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM)

    def step(self, world_size=1):
        """gather -> (sum across ranks) -> fused Adam with the 1/world_size average folded in."""
        b1, b2 = self.betas
        if getattr(self, "_buckets", None) and world_size > 1 and not self.shard:
            # buckets reduced during the backward pass: wait for them; reduce whatever never fired (unused parameters)
            done = set()
            for bi, work in self._pending:
                work.wait()
                done.add(bi)
            self._pending = []
            for bi, (a, b) in enumerate(self._buckets):
                if bi not in done:
                    self._gather_range(a, b)
                    lo, hi = self.slices[a][0], self.slices[b - 1][0] + self.slices[b - 1][1]
                    dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM)
            self.t += 1
            self._adam(self.flat_p, self.flat_g, self.m, self.v, self.lr, b1, b2, self.eps, self.t, 1.0 / world_size)
            return
        g = self.gather_grads()
        self.t += 1
        if self.shard and world_size > 1:
            self._reduce_scatter(self.g_shard, g)
            self._adam(self.p_shard, self.g_shard, self.m, self.v, self.lr, b1, b2, self.eps, self.t, 1.0 / world_size)
            dist.all_gather_into_tensor(self.flat_p, self.p_shard)        # in place: the slice sits at its own offset
            return
        if world_size > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        self._adam(self.flat_p, g, self.m, self.v, self.lr, b1, b2, self.eps, self.t, 1.0 / world_size)


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
    reference's argparse fields (parameter.py:6-79).  Not reference flags, all optional: ``latent_dim`` (128x128 /
    256x256 clips), ``gru_lean`` ('auto' | True | False: ConvGRU BPTT keeps h only and recomputes the gates),
    ``shard_optimizer`` (reduce-scatter / sharded Adam / all-gather), ``sample_path``, ``g_attention`` (the reference's
    commented-out non-local blocks in G), ``sync_bn`` (cross-replica statistics in G's conditional batch norms, the
    reference's TODO), ``dp_shard=(world, rank)`` (act as shard ``rank`` of a ``world``-way data-parallel job WITHOUT a
    process group: same global RNG draws and slicing, no collectives -- for debugging one rank in isolation)."""

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
        self.use_tensorboard = getattr(c, "use_tensorboard", False)
        version = getattr(c, "version", "")
        self.model_save_path = os.path.join(getattr(c, "model_save_path", "./models"), version)
        self.sample_path = os.path.join(getattr(c, "sample_path", "./samples"), version)
        self.log_path = os.path.join(getattr(c, "log_path", "./logs"), version)
        self.shard_optimizer = bool(getattr(c, "shard_optimizer", False))
        self.overlap_allreduce = bool(getattr(c, "overlap_allreduce", True))
        if self.adv_loss not in ("hinge", "wgan-gp"):
            raise ValueError("adv_loss must be 'hinge' or 'wgan-gp'")
        if not torch.cuda.is_available():
            raise RuntimeError("dvdgan_b200.Trainer needs a CUDA device: there is no CPU fallback")
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world_size = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        shard_world, shard_rank = getattr(c, "dp_shard", None) or (self.world_size, self.rank)
        self.shard = shard_batch(self.batch_size, shard_world, shard_rank)      # raises if not divisible
        self.local_batch = self.batch_size // shard_world
        self.g_attention = bool(getattr(c, "g_attention", False))
        ops.set_sync_bn(True if (getattr(c, "sync_bn", False) and self.distributed) else None)
        lean = getattr(c, "gru_lean", "auto")
        if lean == "auto":      # full BPTT state (20 B per hidden element) while it stays under ~30 % of the device;
            # beyond that the layers that free the most memory per recomputed FLOP keep h only (ops.gru_lean_policy)
            budget = 0.30 * torch.cuda.get_device_properties(self.device).total_memory
            lean, self.gru_state_kept = ops.gru_lean_policy(self.local_batch, self.n_frames, self.g_chn,
                                                            self.latent_dim, budget)
        self.gru_lean = lean
        ops.set_gru_lean(self.gru_lean)
        self.writer = None
        if self.use_tensorboard:
            self.build_tensorboard()
        self.build_model()
        if self.pretrained_model:
            self.load_pretrained_model()

    # ------------------------------------------------------------------ model / optimizers
    def build_model(self):
        self.G = Generator(self.z_dim, latent_dim=self.latent_dim, n_class=self.n_class, ch=self.g_chn,
                           n_frames=self.n_frames, attention=self.g_attention).to(self.device)
        self.D_s = SpatialDiscriminator(chn=self.ds_chn, n_class=self.n_class).to(self.device)
        self.D_t = TemporalDiscriminator(chn=self.dt_chn, n_class=self.n_class).to(self.device)
        if self.distributed:            # replicas start identical (DataParallel re-broadcasts every forward)
            for net in (self.G, self.D_s, self.D_t):
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t.data, src=0)
            # ... and draw the same z / labels / frame subsets: one CPU generator stream, rank 0's
            state = torch.get_rng_state().to(self.device)
            dist.broadcast(state, src=0)
            torch.set_rng_state(state.cpu())
        self.select_opt_schr()

    def select_opt_schr(self):
        betas = (self.beta1, self.beta2)
        shard = (self.world_size, self.rank) if (self.shard_optimizer and self.distributed) else None
        self.g_optimizer = FlatAdam(self.G, self.g_lr, betas, shard=shard)
        self.ds_optimizer = FlatAdam(self.D_s, self.d_lr, betas, shard=shard)
        self.dt_optimizer = FlatAdam(self.D_t, self.d_lr, betas, shard=shard)
        if self.distributed and getattr(self, "overlap_allreduce", True):
            self.g_optimizer.enable_overlap(self.G)        # 547 MB at ch = 32; the D arenas (40 MB) are not worth it
        _lr_at(self.lr_schr, 1.0, 0, self.lr_decay)     # validates lr_schr

    def _sched_step(self, opt):
        opt.lr = _lr_at(self.lr_schr, opt.base_lr, opt.t, self.lr_decay)

    def reset_grad(self):
        self.ds_optimizer.zero_grad()
        self.dt_optimizer.zero_grad()
        self.g_optimizer.zero_grad()

    def label_sample(self):
        """trainer.py:84-88 for the GLOBAL batch; this rank's shard goes to the device."""
        label = torch.randint(low=0, high=self.n_class, size=(self.batch_size,))
        return label[self.shard].to(self.device)

    def calc_loss(self, x, real_flag, y=None, y_real_flag=None):
        """trainer.py:114-121; with ``y`` the two-term sum loss(x) + loss(y) in one op."""
        hinge = self.adv_loss == 'hinge'
        sx = -1.0 if real_flag is True else 1.0
        sy = -1.0 if y_real_flag is True else 1.0
        return ops.GanLossFn.apply(hinge, sx, x, sy, y)

    def _local(self, real_videos, real_labels):
        """Accept the loader's global batch (sliced to this rank's shard) or an already-local shard; enforce the
        dtypes the kernels assume (the reference's modules would raise or cast on anything else)."""
        if real_videos.shape[0] == self.batch_size and self.local_batch != self.batch_size:
            real_videos, real_labels = real_videos[self.shard], real_labels[self.shard]
        if real_videos.shape[0] != self.local_batch:
            raise ValueError(f"expected {self.local_batch} clips per rank (global batch {self.batch_size} over "
                             f"{self.world_size} ranks), got {real_videos.shape[0]}")
        if not real_labels.is_cuda:         # host labels: range-check for free (device labels are the caller's contract)
            if real_labels.numel() and (int(real_labels.min()) < 0 or int(real_labels.max()) >= self.n_class):
                raise IndexError(f"class id out of range [0, {self.n_class})")
        real_videos = real_videos.to(self.device, dtype=torch.float32, non_blocking=True)
        real_labels = real_labels.to(self.device, dtype=torch.int64, non_blocking=True)
        return real_videos, real_labels

    # ------------------------------------------------------------------ one step (trainer.py:229-307)
    def train_step(self, real_videos, real_labels):
        """real_videos (B,C,T,H,W) as the loader yields them (global batch or this rank's shard; host or device),
        real_labels (B,) integer class ids."""
        real_videos, real_labels = self._local(real_videos, real_labels)
        real_videos = ops.Permute5Fn.apply(real_videos, (0, 2, 1, 3, 4))       # -> B,T,C,H,W
        for _ in range(self.d_iters):
            real_s = sample_k_frames(real_videos, self.n_frames, self.k_sample)
            z = torch.randn(self.batch_size, self.z_dim)[self.shard].to(self.device)
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
            g_s_loss = self.calc_loss(g_s, True)
            g_t_loss = self.calc_loss(g_t, True)
            g_loss = ops.AddFn.apply(g_s_loss.view(1), g_t_loss.view(1)).view(())
            self.reset_grad()
            g_loss.backward()
        self.g_optimizer.step(self.world_size)
        self._sched_step(self.g_optimizer)
        return {"ds_loss": ds_loss.detach(), "dt_loss": dt_loss.detach(), "g_loss": g_loss.detach(),
                "g_s_loss": g_s_loss.detach(), "g_t_loss": g_t_loss.detach()}

    def check_numerics(self):
        """Reads the library's two device-side counters (synchronises): class ids outside [0, n_class) -- the reference's
        nn.Embedding would have raised -- and forward operands beyond fp16's range, which the fp16 operand planes clamp.
        On the latter the library is switched to bf16 planes (fp32's exponent range, 16-bit operand precision) for every
        later step, loudly."""
        bad = _C.index_errors()
        if bad:
            raise IndexError(f"{bad} class ids outside [0, {self.n_class}) reached the embedding / projection kernels")
        sat = _C.saturation_count()
        if sat:
            import warnings
            _C.set_option("fwd_bf16", 1)
            warnings.warn(f"dvdgan_b200: {sat} groups of forward activations / weights exceeded fp16's range (65504) and "
                          "were clamped in the tensor-core operand planes; switching to bf16 operand planes "
                          "(dvd_set_option('fwd_bf16', 1)) from here on", RuntimeWarning)
        return sat

    # ------------------------------------------------------------------ loop (trainer.py:189-343)
    def epoch2step(self):
        self.epoch = 0
        step_per_epoch = len(self.data_loader)
        self.total_step = self.total_epoch * step_per_epoch
        self.log_step = self.log_epoch * step_per_epoch
        self.sample_step = self.sample_epoch * step_per_epoch
        self.model_save_step = self.model_save_epoch * step_per_epoch

    def train(self):
        data_iter = iter(self.data_loader)
        self.epoch2step()
        # trainer.py:195-197 (the reference leaves fixed_label on the host, which fails on a GPU run; here it moves)
        self.fixed_z = torch.randn(self.test_batch_size * self.n_class, self.z_dim).to(self.device)
        self.fixed_label = torch.tensor([i for i in range(self.n_class) for _ in range(self.test_batch_size)],
                                        dtype=torch.int64).to(self.device)
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
            out = self.train_step(real_videos, real_labels)
            history.append(out)
            if step % self.log_step == 0 and self.rank == 0:
                elapsed = time.time() - start_time
                start_time = time.time()
                log_str = ("Epoch: [%d/%d], Step: [%d/%d], time: %.1fs, ds_loss: %.4f, dt_loss: %.4f, g_s_loss: %.4f, "
                           "g_t_loss: %.4f, g_loss: %.4f, lr: %.2e"
                           % (self.epoch, self.total_epoch, step, self.total_step, elapsed, float(out["ds_loss"]),
                              float(out["dt_loss"]), float(out["g_s_loss"]), float(out["g_t_loss"]),
                              float(out["g_loss"]), self.g_optimizer.lr))
                if self.writer is not None:
                    for k in ("ds_loss", "dt_loss", "g_loss"):
                        self.writer.add_scalar("data/" + k, float(out[k]), step)
                    self.writer.add_text("logs", log_str, step)
                print(log_str)
            if step % self.log_step == 0 or step == self.total_step:
                self.check_numerics()
            if step % self.sample_step == 0 and self.rank == 0:
                self.sample(step)
            if step % self.model_save_step == 0 and self.rank == 0:
                self.save_models(step)
        return history

    # ------------------------------------------------------------------ sampling (trainer.py:322-334)
    @torch.no_grad()
    def sample(self, step, save=True):
        """G in eval mode (running BatchNorm statistics; the spectral norms still advance u, v -- Q3) on the fixed
        noise / one label per class; returns the de-normalised clips (n_class * test_batch_size, T, 3, H, W) in [0, 1]
        and writes one image grid per clip (frames side by side) like the reference's save_image call."""
        if not hasattr(self, "fixed_z"):
            self.fixed_z = torch.randn(self.test_batch_size * self.n_class, self.z_dim).to(self.device)
            self.fixed_label = torch.tensor([i for i in range(self.n_class) for _ in range(self.test_batch_size)],
                                            dtype=torch.int64).to(self.device)
        was_training = self.G.training
        self.G.eval()
        outs = []
        chunk = max(1, self.local_batch)
        for i in range(0, self.fixed_z.shape[0], chunk):            # bounded memory for n_class = 101 / 600
            fake = self.G(self.fixed_z[i:i + chunk], self.fixed_label[i:i + chunk])
            outs.append(denorm(fake.detach().clone()))
        self.G.train(was_training)
        videos = torch.cat(outs, 0)
        if save:
            os.makedirs(self.sample_path, exist_ok=True)
            try:
                from torchvision.utils import make_grid, save_image
            except Exception:           # no torchvision: keep the tensors
                save_image = make_grid = None
            for i in range(self.n_class):
                for j in range(self.test_batch_size):
                    clip = videos[i * self.test_batch_size + j]
                    name = "Class_%d_No.%d_Step_%d" % (i, j, step)
                    if self.writer is not None and make_grid is not None:
                        self.writer.add_image(name.replace("_Step", "/Step"), make_grid(clip.cpu()), step)
                    elif save_image is not None:
                        save_image(clip.cpu(), os.path.join(self.sample_path, name + ".png"))
                    else:
                        torch.save(clip.cpu(), os.path.join(self.sample_path, name + ".pt"))
        return videos

    def build_tensorboard(self):
        """trainer.py:368-373; tensorboardX is optional (absent in this image): without it scalars go to stdout only."""
        try:
            from tensorboardX import SummaryWriter
        except Exception:
            try:
                from torch.utils.tensorboard import SummaryWriter
            except Exception:
                self.writer = None
                return
        self.writer = SummaryWriter(log_dir=self.log_path)

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
        unexpected = set(sd) - set(own)
        if unexpected:
            raise KeyError(f"unexpected keys in {path}: {sorted(unexpected)[:5]}")
        for k, v in sd.items():
            own[k].copy_(v)          # in place: parameters stay views of the flat arena
        missing = set(own) - set(sd)
        if missing:
            raise KeyError(f"missing keys in {path}: {sorted(missing)[:5]}")

    def load_pretrained_model(self):
        for net, tag in ((self.G, "G"), (self.D_s, "Ds"), (self.D_t, "Dt")):
            self._load(net, os.path.join(self.model_save_path, '{}_{}.pth'.format(self.pretrained_model, tag)))
