"""Attribute the engine-vs-reference error of a CelebA trajectory to its rounding sources (CPU only).

    python scripts/attribute_error.py [--steps 40] [--mini]

Runs the fp32 oracle (bit-identical to the reference) and the engine-numerics emulation (oracle/emulate.py) with
all roundings on, then with each one switched off in turn, on the same weights / x_T / noise as the golden
trajectory, and prints max-abs error relative to max|x_0|.  Analysis tool: nothing here is a product path."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ddpm as od, emulate as em, sampler as osmp, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--mini", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()
torch.set_num_threads(os.cpu_count())
cfg = od.MINI_CFG if args.mini else od.CELEBA_CFG
sd = synth.synth_state_dict(od.ddpm_param_shapes(cfg, 1), 1234, "torch_default")
S = cfg["image_size"]
x = synth.synth_noise((1, 3, S, S), 1234)
betas = osmp.make_betas()
seq, seq_next = osmp.make_sequences(999, args.steps)
g = torch.Generator().manual_seed(4321)
noises = {i: torch.randn(x.shape, generator=g) for i in seq}


def run(fwd):
    return osmp.run_trajectory(fwd, x, betas=betas, seq=seq, seq_next=seq_next, t_edit=500, t_addnoise=200, index=0,
                               hs_coeff=(1.0, 1.0), noises=noises)


t0 = time.time()
ref = run(lambda *a, **k: od.ddpm_forward(sd, cfg, *a, **k))
print(f"reference: max|x_0| = {ref.abs().max():.2f}  ({time.time() - t0:.0f}s)", flush=True)
variants = {"engine (all roundings)": em.ALL}
for k in em.ALL:
    variants[f"without {k}"] = {**em.ALL, k: False}
variants["operands only (w16+in16+x16: TF32-class)"] = {**em.NONE, "w16": True, "in16": True, "x16": True}
variants["storage only (store16)"] = {**em.NONE, "store16": True}
for name, fl in variants.items():
    if args.only and args.only not in name:
        continue
    t0 = time.time()
    out = run(lambda *a, **k: em.ddpm_forward(sd, cfg, *a, flags=fl, **k))
    err = (out - ref).abs().max().item()
    print(f"{name:45s} max-abs {err:.4f}  rel {err / ref.abs().max().item():.2e}  ({time.time() - t0:.0f}s)", flush=True)
