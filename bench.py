#!/usr/bin/env python
"""Benchmark of the DVD-GAN G + Ds + Dt training step (BASELINE.json metric: clips/sec, 48f x 64x64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One rank per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); each rank trains on its own shard of
`--batch` clips (weak scaling), gradients are all-reduced over NCCL.  Rank 0 prints ONE JSON line.

  value     clips/s with the step's real clips already resident in HBM (device-timed, max over ranks)
  e2e       clips/s through Trainer.train_step with the clips in pinned HOST memory: H2D copy of the
            clips/labels and D2H read of the three losses inside the timed region
  roofline  the dominant kernel family (implicit-GEMM conv forward/dgrad: ConvGRU gates + all 3x3/5x5/3x3x3
            convs), timed live with CUDA events around every launch during the timed steps
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

METRIC = "clips/sec (48f x 64x64) G+Ds+Dt step"
UNIT = "clips/s"
STEP_TFLOP_PER_CLIP = 8.187     # 3*G_fwd + 9*(Ds_fwd + Dt_fwd), 2*MAC
STEP_HBM_GB_PER_CLIP = 4.42     # SURVEY.md 8(d), compulsory traffic under ideal fusion, fwd+bwd
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed `ncu --set full`
# capture (profiles/): the per-timestep h-half update|reset GEMM of the 32x32 ConvGRU stage, B = 64
NCU_TRAFFIC_BYTES = 173.7e6
NCU_TRAFFIC_OF = ("conv_tma_fwd_kernel<256,0,2,1,8,1> (persistent CTA pairs), M=65536 Cin=256 Cout=512 5x5: dram 81.7 MB read + "
                  "92.0 MB write per launch vs 214 MB algorithmic (67 MB operand planes, mostly still in L2 from the split "
                  "kernel, + 13 MB weight planes + 134 MB fp32 output); profiles/r1/README.md")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU (BASELINE.json configs[1]: 64)")
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--k-sample", type=int, default=8)
    ap.add_argument("--classes", type=int, default=101)
    ap.add_argument("--ch", type=int, default=32)
    ap.add_argument("--latent-dim", type=int, default=4, help="frame side = 16 * latent_dim (4: 64x64; 8: configs[2]; "
                                                             "16: configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=None, help="debug: shrink the CPU sample")
    ap.add_argument("--prof-dump", default=None, help="write the per-shape launch table of the timed steps here")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only (ncu launch list): skip the e2e leg")
    return ap.parse_args()


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
        pretrained_model=None, version="bench", model_save_path="/tmp/dvd_bench", latent_dim=a.latent_dim)


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
    sample = f"{steps} full G+Ds+Dt step(s) of {B} clip x {T}f x 64x64, k={min(a.k_sample, T)}, {a.classes} classes, " \
             f"ch={a.ch}, fp32, {warmup} warm-up"
    return B / dt, cores, sample, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # warm-ups are full steps too (~10-25 s each on 8-16 cores); cap them so the run ends within minutes
    v, cores, sample, dt = cpu_step_time(a, a.steps, min(a.warmup, 1), a.cpu_frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config[1]: {a.frames}f 64x64, {a.classes} classes, ch={a.ch}, k={a.k_sample}; "
                               "CPU sample = 1 clip per step", "timing": "host wall clock"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    B, T = a.batch, a.frames
    torch.manual_seed(1234 + rank)
    tr = Trainer(None, make_cfg(a, B))
    tr.G.train(); tr.D_s.train(); tr.D_t.train()
    n_host = 2
    side = 16 * a.latent_dim
    host_clips = [(torch.rand(B, 3, T, side, side) * 2 - 1).pin_memory() for _ in range(n_host)]
    host_labels = [torch.randint(0, a.classes, (B,)).pin_memory() for _ in range(n_host)]
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
    # CUDA events around every GEMM launch (the roofline numbers); with --prof-dump also around every operand split and
    # helper call (each pair of event records costs ~2 us of stream time, so the default run skips those)
    lib.dvd_prof_enable(0xF if a.prof_dump else 1)
    sec, w0, w1 = timed(step_resident, a.steps)
    lib.dvd_prof_enable(0)
    if a.prof_dump and rank == 0:          # per-shape table of the GEMM / operand-prep launches of the timed steps
        os.makedirs(os.path.dirname(os.path.abspath(a.prof_dump)), exist_ok=True)
        lib.dvd_prof_dump(a.prof_dump.encode())
    launches = lib.dvd_launch_count() - n0
    prof = {}
    for cat, name in ((0, "conv_fwd_dgrad"), (1, "conv_wgrad"), (2, "operand_prep"), (3, "helpers")):
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.dvd_prof_read(cat, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
        prof[name] = (ms.value, fl.value, n.value)
    clocks = sampler.summary(w0, w1) if sampler else None
    sec_e2e = float("nan") if a.no_e2e else timed(step_e2e, a.steps)[0]
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    if rank == 0:
        clips = B * world * a.steps
        value = clips / sec
        k_ms, k_fl, k_n = prof["conv_fwd_dgrad"]
        w_ms, w_fl, w_n = prof["conv_wgrad"]
        achieved = k_fl / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        fma_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12 \
            if clocks else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config[1]: {T}f {side}x{side}, {a.classes} classes, batch={B}/GPU, ch={a.ch}, "
                                   f"k={a.k_sample}, hinge, Adam(5e-5,(0,0.9)), full G+Ds+Dt step "
                                   "(3 optimizer steps, NCCL grad all-reduce if N>1)",
                       "global_batch": B * world, "l2": "inputs and activations are GBs per step (>> 126 MB L2)",
                       "parallelism": f"dp{world}"},
            "e2e": {"value": clips / sec_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": host_clips[0].numel() * 4 + host_labels[0].numel() * 8 + B * 120 * 4 + B * 8,
                    "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "conv_tma_fwd_kernel (tcgen05 implicit-GEMM conv forward + dgrad; operands split into two "
                          "16-bit planes, 3 MMAs per algorithmic MAC, fp32 TMEM accumulators)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "mma_tflops": 3.0 * achieved, "mma_frac": 3.0 * achieved / peak if peak else None,
                "traffic": NCU_TRAFFIC_BYTES, "traffic_of": NCU_TRAFFIC_OF,
                "peak_source": f"{pk_kind} bf16_tflops_sustained",
                "launches_per_step": k_n / a.steps, "avg_launch_ms": k_ms / k_n if k_n else None,
                "share_of_step": k_ms * 1e-3 / sec,
                "fp32_fma_peak_tflops_at_observed_clock": fma_peak,
                "frac_of_fp32_fma": achieved / fma_peak if fma_peak else None,
                "wgrad": {"achieved": w_fl / (w_ms * 1e-3) / 1e12 if w_ms > 0 else 0.0,
                          "share_of_step": w_ms * 1e-3 / sec, "launches_per_step": w_n / a.steps},
                "breakdown_ms_per_step": {k: v[0] / a.steps for k, v in prof.items() if v[2] > 0},
                "operand_prep_gbs": prof["operand_prep"][1] / (prof["operand_prep"][0] * 1e-3) / 1e9
                if prof["operand_prep"][0] > 0 else None,
                "step": {"tflops": STEP_TFLOP_PER_CLIP * value / world, "hbm_gbs": STEP_HBM_GB_PER_CLIP * value / world,
                         "hbm_frac": STEP_HBM_GB_PER_CLIP * value / world / pk["hbm_gbs"]},
            },
            "clocks": clocks,
            "peak_mem_gib": mem_gb,
            "last_losses": losses[-1] if losses else None,
        }
        if world == 1 and not a.no_cpu_baseline:
            v, cores, sample, _ = cpu_step_time(a, 1, 0, a.cpu_frames)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
