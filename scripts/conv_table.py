"""Per-launch table of one edit-step UNet evaluation: kind, shape, ms, TFLOP/s (CUDA events, eager)."""
import argparse, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_model
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="ddpm_celeba_b16"); ap.add_argument("--batch", type=int)
a = ap.parse_args()
family, key, batch, _, ckpt, _ = WORKLOADS[a.workload]; batch = a.batch or batch
m, _ = build_model(family, key, torch.device("cuda:0"), ckpt)
P = m.engine.plan(batch); P.x.normal_(); P.t.fill_(999.0); P.set_coeffs((1.0, 1.0)); P.run_temb()
seq = P.launches(True)
prof = P.profile(edit=True, reps=5)
agg = collections.OrderedDict()
for L, (kind, ms, fl, nb) in zip(seq, prof):
    if kind != "conv": continue
    d = getattr(L, "desc", "?")
    e = agg.setdefault(d, [0, 0.0, 0.0]); e[0] += 1; e[1] += ms; e[2] += fl
tot = sum(e[1] for e in agg.values())
for d, (n, ms, fl) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{d:70s} x{n:3d} {ms:8.3f} ms {ms/tot*100:5.1f}%  {fl/ms/1e9:7.1f} TF/s")
print("conv total ms", tot)
print("--- individual launches of 128->128 @256x256 single-segment fused convs")
for L, (kind, ms, fl, nb) in zip(seq, prof):
    if kind == "conv" and getattr(L, "desc", "") in ("3x3*:128 -> 128 @256x256", "3x3*:128 -> 128 @128x128", "3x3:128 -> 128 @256x256"):
        print(L.desc, f"{ms*1000:.1f} us", f"{fl/ms/1e9:.0f} TF/s", "has_res" if getattr(L, "has_res", None) else "")
print("--- launch order (kind | desc | algorithmic GFLOP), for joining with an ncu launch list")
for L in seq:
    print(f"LAUNCH|{L.kind}|{getattr(L, 'desc', '')}|{L.flops / 1e9:.3f}")
