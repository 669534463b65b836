"""Informational only: the UNMODIFIED reference (baseline/_ref, a git-ignored copy of
/root/reference that travels with gpurun) on the same B200 under PyTorch eager + cuDNN, through its own Trainer.train().
Not the contractual baseline (that is the reference's CPU path, bench.py --impl reference); it gives the clips/s of this
repo a GPU context.   usage: python profiles/ref_eager_b200.py [--batch B] [--steps K]
Prints one JSON line per setting (cuDNN/cuBLAS TF32 on = the PyTorch default for convolutions, and off = strict fp32)."""
import argparse
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--frames", type=int, default=48)
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print(json.dumps({"impl": "reference-eager-gpu", "unavailable": "baseline/_ref not present"}))
        return
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
    sys.modules["tensorboardX"] = tbx
    sys.path.insert(0, REF)
    import trainer as ref_trainer

    B, T = a.batch, a.frames
    n_steps = a.steps + 1          # one warm-up step (cudnn.benchmark autotuning happens there)

    class Loader:
        def __len__(self):
            return n_steps

        def __iter__(self):
            g = torch.Generator().manual_seed(0)
            for _ in range(n_steps):
                yield torch.rand(B, 3, T, 64, 64, generator=g) * 2 - 1, torch.randint(0, 101, (B,), generator=g)

    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True          # main.py:26
        cfg = argparse.Namespace(
            model="dvd-gan", adv_loss="hinge", imsize=64, g_num=5, z_dim=120, g_chn=32, ds_chn=32, dt_chn=32,
            n_frames=T, g_conv_dim=64, d_conv_dim=64, lr_schr="const", lambda_gp=10, total_epoch=1, d_iters=1,
            g_iters=1, batch_size=B, num_workers=0, g_lr=5e-5, d_lr=5e-5, lr_decay=0.9999, beta1=0.0, beta2=0.9,
            pretrained_model=None, n_class=101, k_sample=8, dataset="synthetic", use_tensorboard=False,
            test_batch_size=1, image_path="", log_path="/tmp/ref_eager/log", model_save_path="/tmp/ref_eager/m",
            sample_path="/tmp/ref_eager/s", log_epoch=10 ** 6, sample_epoch=10 ** 6, model_save_epoch=10 ** 6,
            version="eager", gpus=["0"], parallel=False)
        torch.manual_seed(0)
        try:
            tr = ref_trainer.Trainer(Loader(), cfg)
            marks = []
            bw = torch.Tensor.backward

            def backward(self, *aa, **kk):          # the G backward closes a step: time stamps after each third call
                r = bw(self, *aa, **kk)
                marks.append(None)
                if len(marks) % 3 == 0:
                    torch.cuda.synchronize()
                    marks[-1] = time.perf_counter()
                return r
            torch.Tensor.backward = backward
            try:
                tr.train()
            finally:
                torch.Tensor.backward = bw
            torch.cuda.synchronize()
            ts = [m for m in marks if m is not None]
            dt = (ts[-1] - ts[0]) / (len(ts) - 1)          # steps after the warm-up (optimizer step of G included in the next)
            print(json.dumps({"impl": "reference-eager-gpu", "tf32": tf32, "batch": B, "frames": T, "steps": len(ts) - 1,
                              "ms_per_step": dt * 1e3, "clips_per_s": B / dt,
                              "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
                              "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)
            del tr
        except Exception as e:          # informational: report and go on
            print(json.dumps({"impl": "reference-eager-gpu", "tf32": tf32, "batch": B, "error": repr(e)[:300]}), flush=True)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
