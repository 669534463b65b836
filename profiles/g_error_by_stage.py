"""Where does forward error come from?  Generator (ch=32, T=48, 101 classes, B=1) per-stage rel-L2 error of the CUDA
path against the fp64 CPU oracle, next to the fp32 CPU oracle's own error (the reference's noise floor).
usage: python profiles/g_error_by_stage.py [seed] [option=value,... [option=value,...] ...]
Each extra argument is one option set (dvd_set_option) to run the CUDA Generator under, e.g. "oneacc=0" "oneacc=1"."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dvdgan_oracle as O  # noqa: E402
from dvdgan_b200.Module.Generator import Generator  # noqa: E402

from dvdgan_b200 import _C  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
option_sets = sys.argv[2:] or [""]
torch.manual_seed(seed)
T = 48
G = Generator(in_dim=120, latent_dim=4, n_class=101, ch=32, n_frames=T)
sd = {k: v.clone() for k, v in G.state_dict().items()}
z = torch.randn(1, 120)
cls = torch.randint(0, 101, (1,))
cache = f"/tmp/g_ref_{seed}.pt"
if os.path.exists(cache):
    t64, t32 = torch.load(cache)
else:
    t64, t32 = {}, {}
    with torch.no_grad():
        sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        t64["out"] = O.generator_forward(sd64, z.double(), cls, T, 32, 4, taps=t64)
        sd32 = {k: v.clone() for k, v in sd.items()}
        t32["out"] = O.generator_forward(sd32, z, cls, T, 32, 4, taps=t32)
    torch.save((t64, t32), cache)
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
for opts in option_sets:
    for kv in filter(None, opts.split(",")):
        k, _, v = kv.partition("=")
        _C.set_option(k, int(v))
    G.load_state_dict(sd)          # spectral-norm u / v advance on every forward: restart from the same state
    G.to(dev)
    taps = {}
    with torch.no_grad():
        taps["out"] = G(z.to(dev), cls.to(dev), taps=taps)
    print("== options:", opts or "(defaults)", "seed", seed)
    print(f"{'tap':10s} {'cuda vs fp64':>14s} {'cpu-fp32 vs fp64':>18s} {'cuda vs cpu-fp32':>18s}")
    for k in list(t64.keys()):
        print(f"{k:10s} {rel(taps[k], t64[k]):14.3e} {rel(t32[k], t64[k]):18.3e} {rel(taps[k], t32[k]):18.3e}")
    print("max |cuda - cpu-fp32| of the output:", float((taps["out"].cpu() - t32["out"]).abs().max()))
sat = float((t64['out'].abs() > 0.999).float().mean())
print("saturated outputs:", sat)
