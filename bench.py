#!/usr/bin/env python
"""Benchmark of the DVD-GAN G + Ds + Dt training step (BASELINE.json metric: clips/sec, 48f x 64x64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5]

One rank per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE; without it `--gpus N` > 1 re-launches itself under
torchrun).  Weak scaling by default: every rank trains on `--batch` clips (global batch = batch x N); with
`--global-batch G` the G clips are divided over the ranks (strong scaling, the reference's DataParallel split).
Gradients are summed over NCCL.  Rank 0 prints ONE JSON line.  `--config` selects a BASELINE.json configuration
(frame size / clip length / classes / per-GPU batch); the default is configs[1], the one the metric is quoted on.

  value     clips/s with the step's real clips already resident in HBM (device-timed, max over ranks)
  e2e       clips/s through Trainer.train_step with the clips in pinned HOST memory: H2D copy of the
            clips/labels and D2H read of the three losses inside the timed region
  roofline  the dominant kernel family (implicit-GEMM conv forward/dgrad: ConvGRU gates + all 3x3/5x5/3x3x3
            convs), timed live with CUDA events around every launch in a second region of the same K steps (the
            region `value` comes from carries no instrumentation)
  cpu_baseline  the CPU oracle (a port of the reference's PyTorch path, oracle/dvdgan_oracle.py) timed on the
            host cores on a bounded sample (1 clip) of the same workload
`--impl reference` times that CPU path alone (rank 0), K steps after W warm-ups, one clip per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "clips/s"
# algorithmic work (2*MAC; 3*G_fwd + 9*(Ds_fwd + Dt_fwd)) and compulsory HBM traffic under ideal
# fusion, per clip and step, for the BASELINE.json configurations; other shapes report no step-level roofline.
CONFIGS = {
    # name: (frames, latent_dim, classes, batch per GPU, k, step TFLOP per clip, step HBM GB per clip, GPUs quoted)
    2: dict(frames=48, latent_dim=4, classes=101, batch=64, k_sample=8, tflop=8.187, hbm_gb=4.42, gpus=1),
    3: dict(frames=48, latent_dim=8, classes=101, batch=32, k_sample=8, tflop=32.76, hbm_gb=15.40, gpus=8),
    4: dict(frames=12, latent_dim=16, classes=600, batch=8, k_sample=8, tflop=34.73, hbm_gb=16.5, gpus=1),
    5: dict(frames=128, latent_dim=4, classes=101, batch=16, k_sample=8, tflop=21.60, hbm_gb=20.8, gpus=4),
}
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # written by profiles/summarize_ncu.py


def metric_name(a):
    side = 16 * a.latent_dim
    return f"clips/sec ({a.frames}f x {side}x{side}) G+Ds+Dt step"


def workload_name(a, batch_per_gpu, world):
    side = 16 * a.latent_dim
    cfg = next((k for k, c in CONFIGS.items() if (c["frames"], c["latent_dim"], c["classes"]) ==
                (a.frames, a.latent_dim, a.classes)), None)
    tag = f"BASELINE.json configs[{cfg - 1}]" if cfg else "custom shape"
    return (f"{tag}: {a.frames}f {side}x{side}, {a.classes} classes, batch={batch_per_gpu}/GPU x {world} GPU, ch={a.ch}, "
            f"k={a.k_sample}, hinge, Adam(5e-5,(0,0.9)), full G+Ds+Dt step (3 optimizer steps"
            + (", NCCL grad all-reduce)" if world > 1 else ")"))


def gru_policy(lean):
    if lean is None or lean is False:
        return "h + gates + r*h kept"
    if lean is True:
        return "h only, gates recomputed"
    return ("h only + recomputed gates for %d of the 12 ConvGRU layers (Cx,Ch,H,W,k = %s), full state for the rest"
            % (len(lean), sorted(lean)))


def step_work(a):
    """(TFLOP, HBM GB) per clip and step from SURVEY 8(d), or (None, None) off the BASELINE shapes."""
    for c in CONFIGS.values():
        if (c["frames"], c["latent_dim"], c["classes"], c["k_sample"]) == (a.frames, a.latent_dim, a.classes, a.k_sample) \
                and a.ch == 32:
            return c["tflop"], c["hbm_gb"]
    return None, None


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set
    full` capture (profiles/summarize_ncu.py writes the file next to the CSV export it was read from)."""
    try:
        t = json.load(open(NCU_TRAFFIC_FILE))
        return t["conv_tma_fwd"]["dram_bytes_per_launch"], t["conv_tma_fwd"]["source"]
    except Exception:
        return None, "no committed ncu capture found (profiles/ncu_traffic.json)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (1-based): 2 = 48f 64x64 B=64 (default, the metric's), 3 = 48f "
                         "128x128 B=32/GPU (quoted on 8 GPUs), 4 = 12f 256x256 600 classes, 5 = 128f 64x64 (4 GPUs)")
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default: the configuration's)")
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: this many clips in total, divided over the ranks")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--k-sample", type=int, default=None)
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--ch", type=int, default=32)
    ap.add_argument("--latent-dim", type=int, default=None, help="frame side = 16 * latent_dim")
    ap.add_argument("--gru-lean", default="auto", choices=["auto", "on", "off"],
                    help="ConvGRU BPTT keeps h only and recomputes the gates (auto: when the full state would not fit)")
    ap.add_argument("--shard-optimizer", action="store_true", help="reduce-scatter -> sharded Adam -> all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-port", action="store_true", help="reference arm: time the oracle port even if baseline/_ref exists")
    ap.add_argument("--cpu-frames", type=int, default=None, help="debug: shrink the CPU sample")
    ap.add_argument("--cpu-clips", type=int, default=None,
                    help="reference arm: clips per CPU step (default 2, or 1 when more than 12 steps are asked for)")
    ap.add_argument("--prof-dump", default=None, help="write the per-shape launch table of the timed steps here")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only (ncu launch list): skip the e2e leg")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    for k in ("frames", "latent_dim", "classes", "k_sample", "batch"):
        if getattr(a, k) is None:
            setattr(a, k, c[k])
    return a


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, \
            "fallback"


def make_cfg(a, batch):
    return argparse.Namespace(
        adv_loss="hinge", z_dim=120, g_chn=a.ch, ds_chn=a.ch, dt_chn=a.ch, n_frames=a.frames,
        k_sample=a.k_sample, n_class=a.classes, batch_size=batch, d_iters=1, g_lr=5e-5, d_lr=5e-5, beta1=0.0,
        beta2=0.9, lr_schr="const", lr_decay=0.9999, total_epoch=1, log_epoch=10 ** 9, test_batch_size=1,
        pretrained_model=None, version="bench", model_save_path="/tmp/dvd_bench", latent_dim=a.latent_dim,
        gru_lean={"auto": "auto", "on": True, "off": False}[a.gru_lean], shard_optimizer=a.shard_optimizer)


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_step_time(a, steps, warmup, frames=None):
    """Time the oracle's full G+Ds+Dt step (reference trainer.py:229-307 restated) on the host cores."""
    import torch
    from oracle import dvdgan_oracle as O
    from dvdgan_b200.Module.Generator import Generator
    from dvdgan_b200.Module.Discriminators import SpatialDiscriminator, TemporalDiscriminator
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = frames or a.frames
    B = 1
    torch.manual_seed(0)
    # random-init weights of the benchmarked architecture (constructors only; the oracle does the math)
    nets = (Generator(120, a.latent_dim, a.classes, a.ch, T), SpatialDiscriminator(a.ch, a.classes),
            TemporalDiscriminator(a.ch, a.classes))
    sds = [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in nets]
    tr = O.OracleTrainer(*sds, n_frames=T, k_sample=min(a.k_sample, T), n_class=a.classes, batch_size=B,
                         g_chn=a.ch, adv_loss="hinge", latent_dim=a.latent_dim)
    side = 16 * a.latent_dim
    clip = torch.rand(B, 3, T, side, side) * 2 - 1
    lab = torch.randint(0, a.classes, (B,))
    for _ in range(warmup):
        tr.step(clip, lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(clip, lab)
    dt = (time.perf_counter() - t0) / steps
    sample = f"{steps} full G+Ds+Dt step(s) of {B} clip x {T}f x {side}x{side}, k={min(a.k_sample, T)}, " \
             f"{a.classes} classes, ch={a.ch}, fp32, {warmup} warm-up"
    return B / dt, cores, sample, dt, B


REF_DIR = os.path.join(ROOT, "baseline", "_ref")          # git-ignored copy of /root/reference; travels with the snapshot


def ref_cpu_step_time(a, steps, warmup, frames=None):
    """Time the UNMODIFIED reference's own Trainer.train() (trainer.py:189-343) on the host cores: baseline/_ref on
    sys.path, the three shims of SURVEY 8c (tensorboardX stub, .cuda() no-ops, a synthetic loader), B = 2 clips per step
    (the per-clip CPU cost is flat in B).  Only for 64x64 clips: the reference trainer cannot build a
    Generator with another latent_dim (trainer.py:349)."""
    import types
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *aa, **kk: None})
    sys.modules["tensorboardX"] = tbx
    sys.path.insert(0, REF_DIR)
    torch.nn.Module.cuda = lambda self, *aa, **kk: self
    torch.Tensor.cuda = lambda self, *aa, **kk: self
    import trainer as ref_trainer
    T = frames or a.frames
    # clips per CPU step: 2 while K + 1 such steps (~13 s each on 16 cores) end within ~3 minutes, else 1 --
    # the per-clip CPU cost is flat in B, so the clips/s figure does not depend on the choice
    B = a.cpu_clips or (2 if steps + warmup <= 12 else 1)
    n = steps + warmup
    torch.manual_seed(0)
    clips = [torch.rand(B, 3, T, 64, 64) * 2 - 1 for _ in range(n)]
    labels = [torch.randint(0, a.classes, (B,)) for _ in range(n)]

    class Loader:
        def __len__(self):
            return n

        def __iter__(self):
            return iter(zip(clips, labels))
    cfg = argparse.Namespace(
        model="dvd-gan", adv_loss="hinge", imsize=64, g_num=5, z_dim=120, g_chn=a.ch, ds_chn=a.ch, dt_chn=a.ch,
        n_frames=T, g_conv_dim=64, d_conv_dim=64, lr_schr="const", lambda_gp=10, total_epoch=1, d_iters=1, g_iters=1,
        batch_size=B, num_workers=0, g_lr=5e-5, d_lr=5e-5, lr_decay=0.9999, beta1=0.0, beta2=0.9,
        pretrained_model=None, n_class=a.classes, k_sample=min(a.k_sample, T), dataset="synthetic",
        use_tensorboard=False, test_batch_size=1, image_path="", log_path="/tmp/dvd_ref/log",
        model_save_path="/tmp/dvd_ref/m", sample_path="/tmp/dvd_ref/s", log_epoch=10 ** 6, sample_epoch=10 ** 6,
        model_save_epoch=10 ** 6, version="bench", gpus="", parallel=False)
    real_stdout = os.dup(1)          # the reference prints banners: keep stdout for the one JSON line
    os.dup2(2, 1)
    try:
        tr = ref_trainer.Trainer(Loader(), cfg)
        marks = []
        bw = torch.Tensor.backward

        def backward(self, *aa, **kk):          # the third backward of a step (g_loss) closes it
            r = bw(self, *aa, **kk)
            marks.append(time.perf_counter())
            return r
        torch.Tensor.backward = backward
        t_start = time.perf_counter()
        try:
            tr.train()
        finally:
            torch.Tensor.backward = bw
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    ends = [t_start] + marks[2::3]          # step boundaries (the optimizer step of G falls into the next interval)
    dt = (ends[-1] - ends[warmup]) / steps
    sample = f"{steps} full G+Ds+Dt step(s) of the unmodified reference Trainer.train() (baseline/_ref), {B} clip{'s' if B > 1 else ''} x {T}f x " \
             f"64x64, k={min(a.k_sample, T)}, {a.classes} classes, ch={a.ch}, fp32, {warmup} warm-up"
    return B / dt, cores, sample, dt, B


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # warm-ups are full steps too (~10-25 s each on 8-16 cores); cap them so the run ends within minutes
    kind = "reference" if (os.path.isdir(REF_DIR) and a.latent_dim == 4 and not a.force_port) else "port"
    if kind == "reference":
        v, cores, sample, dt, clips_per_step = ref_cpu_step_time(a, a.steps, min(a.warmup, 1), a.cpu_frames)
    else:
        v, cores, sample, dt, clips_per_step = cpu_step_time(a, a.steps, min(a.warmup, 1), a.cpu_frames)
    line = {
        "impl": "reference", "metric": metric_name(a), "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, a.batch, max(a.gpus, 1)) + "; CPU sample = "
                               + ("1 clip" if clips_per_step == 1 else f"{clips_per_step} clips") + " per step",
                   "timing": "host wall clock"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (t0 <= ts <= t1):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import ctypes
    import torch
    import torch.distributed as dist
    from dvdgan_b200 import _C
    from dvdgan_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # exactly ONE line on stdout: libraries (NCCL prints its version there) write to fd 1 too, so park the real stdout
    # and point fd 1 at stderr until the JSON line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _C.lib()
    pk, pk_kind = peaks()
    T = a.frames
    if a.global_batch:
        if a.global_batch % world:
            raise SystemExit(f"--global-batch {a.global_batch} is not divisible by {world} ranks")
        B = a.global_batch // world
    else:
        B = a.batch
    torch.manual_seed(1234)              # one CPU generator stream for every rank (Trainer broadcasts rank 0's anyway)
    tr = Trainer(None, make_cfg(a, B * world))        # config.batch_size is the GLOBAL batch
    tr.G.train(); tr.D_s.train(); tr.D_t.train()
    n_host = 2
    side = 16 * a.latent_dim
    gen = torch.Generator().manual_seed(4321 + rank)          # each rank's own shard of "real" clips
    host_clips = [(torch.rand(B, 3, T, side, side, generator=gen) * 2 - 1).pin_memory() for _ in range(n_host)]
    host_labels = [torch.randint(0, a.classes, (B,), generator=gen).pin_memory() for _ in range(n_host)]
    dev_clips = [c.to(dev) for c in host_clips]
    dev_labels = [l.to(dev) for l in host_labels]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / 1e3, w0, w1

    def step_resident(i):
        tr.train_step(dev_clips[i % n_host], dev_labels[i % n_host])

    losses = []

    def step_e2e(i):
        c = host_clips[i % n_host].to(dev, non_blocking=True)
        l = host_labels[i % n_host].to(dev, non_blocking=True)
        out = tr.train_step(c, l)
        losses.append([float(out[k]) for k in ("ds_loss", "dt_loss", "g_loss")])   # D2H read of the result

    for i in range(a.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.dvd_launch_count()
    sec, w0, w1 = timed(step_resident, a.steps)               # the headline region: no instrumentation
    launches = lib.dvd_launch_count() - n0
    # Second region of the same K steps with CUDA events around every GEMM launch (the roofline numbers); with
    # --prof-dump also around every operand split and helper call (each pair of event records costs ~2 us of stream
    # time).  The ConvGRU time loops run layer by layer as ONE chain here: with the default layer wavefront / batch
    # chains on several streams the kernels of one chain share the SMs with another chain's, and an event bracket
    # would time the sharing, not the kernel.
    from dvdgan_b200 import ops
    chains, wave = _C.get_option("gru_streams"), dict(ops.GRU_WAVEFRONT)
    _C.set_option("gru_streams", 1)
    ops.GRU_WAVEFRONT["enabled"] = 0
    lib.dvd_prof_enable(0xF if a.prof_dump else 1)
    sec_prof = timed(step_resident, a.steps)[0]
    lib.dvd_prof_enable(0)
    _C.set_option("gru_streams", chains)
    ops.GRU_WAVEFRONT.update(wave)
    if a.prof_dump and rank == 0:          # per-shape table of the GEMM / operand-prep launches of the instrumented steps
        os.makedirs(os.path.dirname(os.path.abspath(a.prof_dump)), exist_ok=True)
        lib.dvd_prof_dump(a.prof_dump.encode())
    prof = {}
    for cat, name in ((0, "conv_fwd_dgrad"), (1, "conv_wgrad"), (2, "operand_prep"), (3, "helpers")):
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.dvd_prof_read(cat, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
        prof[name] = (ms.value, fl.value, n.value)
    clocks = sampler.summary(w0, w1) if sampler else None
    sec_e2e = float("nan") if a.no_e2e else timed(step_e2e, a.steps)[0]
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    scratch_hw, _ = _C.scratch_bytes()          # operand planes live in the cudaMallocAsync pool, outside torch's allocator
    bad_ids = _C.index_errors()
    saturated = _C.saturation_count()

    if rank == 0:
        clips = B * world * a.steps
        value = clips / sec
        k_ms, k_fl, k_n = prof["conv_fwd_dgrad"]
        w_ms, w_fl, w_n = prof["conv_wgrad"]
        achieved = k_fl / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        fma_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12 \
            if clocks else None
        tflop_clip, hbm_clip = step_work(a)
        traffic, traffic_of = ncu_traffic()
        line = {
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if a.global_batch else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, B, world),
                       "global_batch": B * world, "l2": "inputs and activations are GBs per step (>> 126 MB L2)",
                       "parallelism": f"dp{world}", "convgru_overlap": {"layer_wavefront": wave, "batch_chains": chains}, "gru_bptt_state": gru_policy(tr.gru_lean), "optimizer": "sharded (RS/Adam/AG)" if a.shard_optimizer and world > 1
                       else "replicated (all-reduce + full Adam)"},
            "e2e": {"value": clips / sec_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": host_clips[0].numel() * 4 + host_labels[0].numel() * 8 + B * 120 * 4 + B * 8,
                    "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "conv_tma_fwd_kernel (tcgen05 implicit-GEMM conv forward + dgrad; operands split into two "
                          "bf16 planes, 3 MMAs per algorithmic MAC, fp32 TMEM accumulators)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "mma_tflops": 3.0 * achieved, "mma_frac": 3.0 * achieved / peak if peak else None,
                "traffic": traffic, "traffic_of": traffic_of,
                "peak_source": f"{pk_kind} bf16_tflops_sustained",
                "launches_per_step": k_n / a.steps, "avg_launch_ms": k_ms / k_n if k_n else None,
                "share_of_step": k_ms * 1e-3 / sec_prof,
                "timed_in": f"a second region of the same {a.steps} steps with CUDA events around every GEMM launch and "
                            f"the ConvGRU time loops layer by layer on one stream (per-launch times exclusive): "
                            f"{sec_prof / a.steps * 1e3:.1f} ms/step there; `value` is the un-instrumented region "
                            f"(ConvGRU layers as a wavefront on one stream each / {chains} batch chains)",
                "fp32_fma_peak_tflops_at_observed_clock": fma_peak,
                "frac_of_fp32_fma": achieved / fma_peak if fma_peak else None,
                "wgrad": {"achieved": w_fl / (w_ms * 1e-3) / 1e12 if w_ms > 0 else 0.0,
                          "share_of_step": w_ms * 1e-3 / sec_prof, "launches_per_step": w_n / a.steps},
                "breakdown_ms_per_step": {k: v[0] / a.steps for k, v in prof.items() if v[2] > 0},
                "operand_prep_gbs": prof["operand_prep"][1] / (prof["operand_prep"][0] * 1e-3) / 1e9
                if prof["operand_prep"][0] > 0 else None,
                "step": {"tflops": tflop_clip * value / world, "hbm_gbs": hbm_clip * value / world,
                         "hbm_frac": hbm_clip * value / world / pk["hbm_gbs"]} if tflop_clip else None,
            },
            "clocks": clocks,
            "peak_mem_gib": mem_gb + scratch_hw / 2 ** 30,
            "peak_mem_detail_gib": {"torch_allocator": mem_gb, "operand_plane_pool_high_water": scratch_hw / 2 ** 30},
            "out_of_range_class_ids": bad_ids, "fp16_saturated_operand_groups": saturated,
            "last_losses": losses[-1] if losses else None,
        }
        if world == 1 and not a.no_cpu_baseline:
            # the reference arm, bounded to one step, in its own process (it patches .cuda() and puts baseline/_ref on
            # sys.path): the unmodified reference when baseline/_ref is there, the oracle port otherwise
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--config", str(a.config), "--frames", str(a.frames), "--classes", str(a.classes),
                   "--k-sample", str(a.k_sample), "--ch", str(a.ch), "--latent-dim", str(a.latent_dim)]
            if a.cpu_frames:
                cmd += ["--cpu-frames", str(a.cpu_frames)]
            r = subprocess.run(cmd, capture_output=True, text=True, env={k: v for k, v in os.environ.items()
                                                                       if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            try:
                line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception:
                line["cpu_baseline"] = {"error": (r.stderr or r.stdout)[-300:]}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def relaunch_under_torchrun(a):
    """`python bench.py --gpus N` without torchrun: start the N ranks ourselves (one node, 127.0.0.1 rendezvous)."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    raise SystemExit(subprocess.call(cmd))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        world_env = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
            relaunch_under_torchrun(args)
        if args.gpus != world_env:
            raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world_env}")
        run_b200(args)
