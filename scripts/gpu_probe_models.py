"""Exploratory GPU run: engine vs golden fixtures / oracle, prints error metrics (not a test)."""
import json
import os
import sys
import time
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asyrp_official_b200 import modules  # noqa: E402
from asyrp_official_b200.schedule import Schedule, make_sequences  # noqa: E402
from asyrp_official_b200.utils.diffusion_utils import denoising_step  # noqa: E402
from oracle import adm as oa, ddpm as od, sampler as osmp, synth  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")
report = {}


def err(name, out, ref):
    out, ref = out.double().cpu(), torch.as_tensor(ref).double()
    e = (out - ref).abs().max().item()
    m = ref.abs().max().item()
    report[name] = dict(max_abs=e, ref_absmax=m, rel=e / max(m, 1e-30))
    print(f"{name:40s} max_abs={e:.3e} ref_absmax={m:.3e} rel={e / max(m, 1e-30):.3e}", flush=True)


def ddpm_model(cfg, n_delta=1, style="jittered"):
    c = NS(model=NS(**{**cfg, "dropout": 0.0, "resamp_with_conv": True}), data=NS(image_size=cfg["image_size"]))
    m = modules.DDPM(c)
    m.setattr_layers(n_delta)
    m.load_state_dict(synth.synth_state_dict(od.ddpm_param_shapes(cfg, n_delta), seed=1234, style=style))
    return m.to(dev)


def adm_model(hp, n_delta=1, style="jittered"):
    m = modules._create_adm(hp)
    m.setattr_layers(n_delta)
    m.load_state_dict(synth.synth_state_dict(oa.adm_param_shapes(hp, n_delta), seed=1234, style=style))
    return m.to(dev)


def mini(family):
    gold = np.load(os.path.join(G, f"{family}_mini.npz"))
    m = ddpm_model(od.MINI_CFG) if family == "ddpm" else adm_model(oa.MINI_HP)
    S = 32
    x = synth.synth_noise((2, 3, S, S), seed=1234)
    cases = {"plain": dict(t=999.0), "edit": dict(t=600.0, index=0, t_edit=500, hs_coeff=(1.0, 0.7)),
             "pass": dict(t=300.0, index=0, t_edit=500, hs_coeff=(1.0, 0.7))}
    for name, kw in cases.items():
        kw = dict(kw)
        t = torch.ones(2) * kw.pop("t")
        r = m(x.to(dev), t.to(dev), **kw)
        for key, a in zip(("et", "et_mod", "delta_h", "middle_h"), r):
            if a is not None:
                err(f"{family}_mini/{name}/{key}", a, gold[f"{name}_{key}"])
    # trajectory through the drop-in denoising_step
    betas = osmp.make_betas()
    seq, seq_next = make_sequences(999, 10)
    g = torch.Generator().manual_seed(4321)
    noises = {i: torch.randn(x.shape, generator=g) for i in seq}
    xx = x.to(dev)
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t, tn = torch.ones(2) * i, torch.ones(2) * j
        xx, x0_t, _, _ = denoising_step(xx, t, tn, models=m, b=betas, eta=1.0 if i < 300 else 0.0,
                                        learn_sigma=(family == "adm"), index=0, t_edit=500, hs_coeff=(1.0, 1.0),
                                        noise=noises[i])
    err(f"{family}_mini/traj10_step_api", xx, gold["traj_x0"])
    # same trajectory as one CUDA graph
    sch = Schedule(betas, seq, seq_next, t_edit=500, t_addnoise=300, hs_coeff=(1.0, 1.0))
    nz = torch.stack([noises[s.t] for s in sch.steps if s.c1 != 0.0]).to(dev)
    for use_graph in (False, True):
        x0 = m.engine.sample(x.to(dev), sch, noise=nz, use_graph=use_graph)
        err(f"{family}_mini/traj10_sample_graph{int(use_graph)}", x0, gold["traj_x0"])
    x0b = m.engine.sample(x.to(dev), sch, noise=nz, use_graph=True)
    report[f"{family}_mini/graph_replay_deterministic"] = bool(torch.equal(x0, x0b))
    print("replay deterministic:", torch.equal(x0, x0b))


def full(name, family, cfg):
    gold = np.load(os.path.join(G, f"{name}_fwd.npz"))
    t0 = time.time()
    m = (ddpm_model(cfg, style="torch_default") if family == "ddpm" else adm_model(cfg, style="torch_default"))
    x = synth.synth_noise((1, 3, 256, 256), seed=1234)
    r = m(x.to(dev), (torch.ones(1) * 999).to(dev), index=0, t_edit=500, hs_coeff=(1.0, 1.0))
    torch.cuda.synchronize()
    print(f"{name}: build+first forward {time.time() - t0:.1f}s; pool {m.engine.plan(1).pool.total / 2**20:.0f} MiB")
    for key, a in zip(("et", "et_mod", "delta_h", "middle_h"), r):
        ref = gold[key]
        a = a[..., ::4, ::4] if a.shape[-1] == 256 else a
        err(f"{name}/fwd/{key}", a, ref)
    return m


if __name__ == "__main__":
    mini("ddpm")
    mini("adm")
    m = full("ddpm_celeba", "ddpm", od.CELEBA_CFG)
    gold = np.load(os.path.join(G, "ddpm_celeba_traj40.npz"))
    betas = osmp.make_betas()
    seq, seq_next = make_sequences(999, 40)
    g = torch.Generator().manual_seed(4321)
    noises = {i: torch.randn(1, 3, 256, 256, generator=g) for i in seq}
    sch = Schedule(betas, seq, seq_next, t_edit=500, t_addnoise=200, hs_coeff=(1.0, 1.0))
    nz = torch.stack([noises[s.t] for s in sch.steps if s.c1 != 0.0]).to(dev)
    x = synth.synth_noise((1, 3, 256, 256), seed=1234).to(dev)
    t0 = time.time()
    x0 = m.engine.sample(x, sch, noise=nz)
    torch.cuda.synchronize()
    t1 = time.time()
    x0 = m.engine.sample(x, sch, noise=nz)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"traj40 B=1: capture+run {t1 - t0:.2f}s, replay {t2 - t1:.3f}s")
    err("ddpm_celeba/traj40/x0", x0, gold["x0_full_f16"].astype(np.float32))
    report["ddpm_celeba/traj40/replay_s_B1"] = t2 - t1
    del m
    torch.cuda.empty_cache()
    full("adm_afhq", "adm", oa.AFHQ_HP)
    torch.cuda.empty_cache()
    full("adm_imagenet", "adm", oa.IMAGENET_HP)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "probe_models.json"), "w"), indent=1)
