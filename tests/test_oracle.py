"""Pin the CPU oracle against golden vectors produced by the REFERENCE'S OWN modules (tests/golden/make_golden.py).

The fixtures were written bit-exactly equal to the oracle on the build machine; here a 2e-5 relative tolerance is
allowed because oneDNN may pick different conv algorithms on a CPU with another ISA / thread count."""
import os

import numpy as np
import pytest
import torch

from oracle import adm as o_adm, ddpm as o_ddpm, sampler as o_smp, synth

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _close(a, ref, what):
    ref = torch.as_tensor(ref)
    err = (a - ref).abs().max().item()
    assert err <= TOL * max(1.0, ref.abs().max().item()), f"{what}: {err}"


def _mini(family):
    if family == "ddpm":
        cfg = o_ddpm.MINI_CFG
        sd = synth.synth_state_dict(o_ddpm.ddpm_param_shapes(cfg, 1), 1234, "jittered")
        return cfg, sd, (lambda *a, **k: o_ddpm.ddpm_forward(sd, cfg, *a, **k)), False
    cfg = o_adm.MINI_HP
    sd = synth.synth_state_dict(o_adm.adm_param_shapes(cfg, 1), 1234, "jittered")
    return cfg, sd, (lambda *a, **k: o_adm.adm_forward(sd, cfg, *a, **k)), True


@pytest.mark.parametrize("family", ["ddpm", "adm"])
def test_mini_forward_matches_reference(family):
    gold = np.load(os.path.join(G, f"{family}_mini.npz"))
    cfg, sd, fwd, _ = _mini(family)
    x = synth.synth_noise((2, 3, cfg["image_size"], cfg["image_size"]), 1234)
    cases = {"plain": dict(t=999.0), "edit": dict(t=600.0, index=0, t_edit=500, hs_coeff=(1.0, 0.7)),
             "pass": dict(t=300.0, index=0, t_edit=500, hs_coeff=(1.0, 0.7))}
    for name, kw in cases.items():
        kw = dict(kw)
        t = torch.ones(2) * kw.pop("t")
        out = fwd(x, t, **kw)
        for key, a in zip(("et", "et_mod", "delta_h", "middle_h"), out):
            if a is None:
                assert f"{name}_{key}" not in gold
            else:
                _close(a, gold[f"{name}_{key}"], f"{family} {name} {key}")
    # below t_edit the two decoder passes are identical (reference: h2 = h, ddpm/diffusion.py:541-542)
    out = fwd(x, torch.ones(2) * 300.0, index=0, t_edit=500)
    assert torch.equal(out[0], out[1])


@pytest.mark.parametrize("family", ["ddpm", "adm"])
def test_mini_trajectory_matches_reference(family):
    """10-step Asyrp edit, t_edit=500, eta=1 below t_addnoise=300 with the reference's randn_like replaced by the
    same pre-drawn noise"""
    gold = np.load(os.path.join(G, f"{family}_mini.npz"))
    cfg, sd, fwd, learn_sigma = _mini(family)
    x = synth.synth_noise((2, 3, cfg["image_size"], cfg["image_size"]), 1234)
    betas = o_smp.make_betas()
    seq, seq_next = o_smp.make_sequences(999, 10)
    g = torch.Generator().manual_seed(int(gold["traj_noise_seed"]))
    noises = {i: torch.randn(x.shape, generator=g) for i in seq}
    rec = []
    xf = o_smp.run_trajectory(fwd, x, betas=betas, seq=seq, seq_next=seq_next, t_edit=500, t_addnoise=300, index=0,
                              hs_coeff=(1.0, 1.0), learn_sigma=learn_sigma, noises=noises, record=rec)
    _close(xf, gold["traj_x0"], f"{family} traj x0")
    _close(torch.stack([r[1] for r in rec]), gold["traj_x0t"], f"{family} traj x0_t")


def test_schedule_and_tables():
    """beta / alpha-bar / sequence construction (utils/diffusion_utils.py:5-9, diffusion_latent.py:41-61,570-574)"""
    b = o_smp.get_beta_schedule(beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    assert b.dtype == np.float64 and b.shape == (1000,) and b[0] == 1e-4 and abs(b[-1] - 0.02) < 1e-15
    seq, nxt = o_smp.make_sequences(999, 40)
    assert seq[:6] == [0, 25, 51, 76, 102, 128] and seq[-1] == 999 and nxt[0] == -1 and nxt[1:] == seq[:-1]
    lv = o_smp.make_logvar(b)
    assert lv.shape == (1000,) and np.isfinite(lv).all()
    t = torch.tensor([0.0, 999.0])
    e = o_smp.extract(torch.from_numpy(b).float(), t, (2, 3, 4, 4))
    assert e.shape == (2, 1, 1, 1) and e[1, 0, 0, 0] == torch.tensor(b[999]).float()


@pytest.mark.parametrize("name,family,cfg", [("ddpm_celeba", "ddpm", o_ddpm.CELEBA_CFG),
                                             ("adm_afhq", "adm", o_adm.AFHQ_HP)])
def test_full_size_forward_matches_reference(name, family, cfg):
    """256x256 Asyrp forward at t=999 (stride-4 subsample + moments of the reference output)"""
    gold = np.load(os.path.join(G, f"{name}_fwd.npz"))
    torch.set_num_threads(os.cpu_count())
    if family == "ddpm":
        sd = synth.synth_state_dict(o_ddpm.ddpm_param_shapes(cfg, 1), 1234, "torch_default")
        out = o_ddpm.ddpm_forward(sd, cfg, synth.synth_noise((1, 3, 256, 256), 1234), torch.ones(1) * 999, index=0,
                                  t_edit=500, hs_coeff=(1.0, 1.0))
    else:
        sd = synth.synth_state_dict(o_adm.adm_param_shapes(cfg, 1), 1234, "torch_default")
        out = o_adm.adm_forward(sd, cfg, synth.synth_noise((1, 3, 256, 256), 1234), torch.ones(1) * 999, index=0,
                                t_edit=500, hs_coeff=(1.0, 1.0))
    for key, a in zip(("et", "et_mod", "delta_h", "middle_h"), out):
        _close(a[..., ::4, ::4] if a.shape[-1] == 256 else a, gold[key], f"{name} {key}")
        assert abs(a.abs().max().item() - float(gold[key + "_absmax"])) <= TOL * 10


@pytest.mark.parametrize("family", ["ddpm", "adm"])
def test_mini_inversion_matches_reference(family):
    """DDIM inversion (t < t_next, index=None) through the oracle's denoising_step vs the reference's"""
    gold = np.load(os.path.join(G, f"{family}_mini.npz"))
    cfg, sd, fwd, learn_sigma = _mini(family)
    betas = o_smp.make_betas()
    seq, seq_next = o_smp.make_sequences(999, 10)
    logv = None if learn_sigma else o_smp.make_logvar(o_smp.get_beta_schedule(beta_start=1e-4, beta_end=0.02,
                                                                             num_diffusion_timesteps=1000))
    x = torch.from_numpy(gold["inv_x0"])
    for i, j in zip(seq_next[1:], seq[1:]):
        x = o_smp.denoising_step(x, torch.ones(2) * i, torch.ones(2) * j, model=fwd, logvars=logv, b=betas, eta=0.0,
                                 learn_sigma=learn_sigma)[0]
    _close(x, gold["inv_xT"], f"{family} inversion")


def test_image_range_fixture_first_steps_match_reference():
    """tests/golden/ddpm_celeba_bounded_t400.npz (the pipeline in the image range, full size; make_golden.py
    bounded_fixture): the oracle's first inversion step from the stored image stays bounded and its first edit step from
    the reference's x_T reproduces the reference's recorded max|x0_t| of that step."""
    gold = np.load(os.path.join(G, "ddpm_celeba_bounded_t400.npz"))
    cfg = o_ddpm.CELEBA_CFG
    sd = synth.synth_state_dict(o_ddpm.ddpm_param_shapes(cfg, 1), 1234, "torch_default")
    blk = torch.load(os.path.join(G, "checkpoint", "smiling_LC_CelebA_HQ_t999_ninv40_ngen40_0.pth"), map_location="cpu",
                     weights_only=True)["0"]
    for k, v in blk.items():
        sd["layer_0." + k] = v
    sd["conv_out.weight"] = sd["conv_out.weight"] * float(gold["gamma"])
    sd["conv_out.bias"] = sd["conv_out.bias"] * float(gold["gamma"])
    fwd = lambda *a, **k: o_ddpm.ddpm_forward(sd, cfg, *a, **k)  # noqa: E731
    seq = gold["seq"].tolist()
    betas = o_smp.make_betas()
    assert float(np.abs(gold["x0_out"]).max()) < 2 and float(np.abs(gold["x_T"]).max()) < 2 and gold["x0t_absmax"].max() < 2
    x_T = torch.from_numpy(gold["x_T"])
    # first reverse step (t = t_0, an edit step: t >= t_edit), deterministic (t >= t_addnoise)
    _, x0_t, _, _ = o_smp.denoising_step(x_T, torch.ones(1) * seq[-1], torch.ones(1) * seq[-2], model=fwd, b=betas, eta=0.0,
                                         index=0, t_edit=int(gold["t_edit"]), hs_coeff=(1.0, 1.0))
    ref = float(gold["x0t_absmax"][0])
    assert abs(x0_t.abs().max().item() - ref) <= TOL * max(1.0, ref), (x0_t.abs().max().item(), ref)
    # first inversion step of the stored image (t = 0 -> seq[1]): x_1 = sqrt(a_1) x0 + O(gamma)
    x0 = torch.from_numpy(gold["x0_in"]).float()
    x1 = o_smp.denoising_step(x0, torch.ones(1) * seq[0], torch.ones(1) * seq[1], model=fwd, b=betas, eta=0.0)[0]
    assert (x1 - x0).abs().max().item() < 0.1 and x1.abs().max().item() < 1.1
