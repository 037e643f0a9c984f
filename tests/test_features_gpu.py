"""GPU tests of the runner features around the hot loop: every flag of the reference's
main.py that reaches run_test / save_image / denoising_step is exercised through `Asyrp.run_test` on a reduced DDPM
configuration (a YAML config in the reference's format) and checked against the CPU oracle on the same weights, latents
and noise.  Sizes are small so that the oracle side takes seconds."""
import os

import numpy as np
import pytest
import torch
import yaml

from asyrp_official_b200 import main as cli, modules, synthetic
from asyrp_official_b200.diffusion_latent import Asyrp
from asyrp_official_b200.schedule import Schedule, make_sequences
from asyrp_official_b200.utils.diffusion_utils import denoising_step
from oracle import ddpm as od, sampler as osmp

pytestmark = pytest.mark.gpu
TOL = 2e-3  # relative to max(1, max|ref|): short trajectories of the reduced UNet (measured 3e-4 .. 9e-4)
CFG = od.MINI_CFG
MID_C = CFG["ch"] * CFG["ch_mult"][-1]


def _check(out, ref, tol, what):
    ref = torch.as_tensor(ref).double()
    e, m = (out.double().cpu() - ref).abs().max().item(), ref.abs().max().item()
    assert e <= tol * max(1.0, m), f"{what}: max-abs err {e:.3e}, max|ref| {m:.3e}, tol {tol}"


def _write_cfg(path):
    cfg = {"data": dict(dataset="CelebA_HQ", category="CelebA_HQ", image_size=CFG["image_size"], channels=3,
                        rescaled=True, num_workers=0),
           "model": dict(type="simple", in_channels=3, out_ch=3, ch=CFG["ch"], ch_mult=list(CFG["ch_mult"]),
                         num_res_blocks=CFG["num_res_blocks"], attn_resolutions=list(CFG["attn_resolutions"]), dropout=0.0,
                         var_type="fixedsmall", ema_rate=0.999, ema=True, resamp_with_conv=True),
           "diffusion": dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000),
           "sampling": dict(batch_size=4, last_only=True)}
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


def _delta_block_ckpt(path, seed):
    from types import SimpleNamespace as NS
    ns = NS(model=NS(**{**CFG, "dropout": 0.0, "resamp_with_conv": True}), data=NS(image_size=CFG["image_size"]))
    m = modules.DDPM(ns)
    m.setattr_layers(1)
    synthetic.randomize_(m, seed=seed, style="jittered")
    blk = {k: v.clone() for k, v in m.layer_0.state_dict().items()}
    torch.save({"0": blk}, path)
    return blk


class Run:
    """one Asyrp.run_test invocation in a scratch directory, with the latents it drew recorded"""

    def __init__(self, tmp_path, monkeypatch, dev, extra, n_step=4, n_test_img=1, n_train_img=0, bs=1, t_edit=500,
                 t_addnoise=0):
        monkeypatch.chdir(tmp_path)
        os.makedirs("checkpoint", exist_ok=True)
        cfg = _write_cfg(os.path.join(str(tmp_path), "mini.yml"))
        argv = ["--run_test", "--config", cfg, "--exp", "./runs/attr", "--edit_attr", "attr",
                "--do_train", "1" if n_train_img else "0", "--do_test", "1" if n_test_img else "0",
                "--n_test_img", str(n_test_img), "--n_train_img", str(n_train_img), "--bs_train", str(bs),
                "--t_0", "999", "--n_inv_step", str(n_step), "--n_train_step", str(n_step), "--n_test_step", str(n_step),
                "--load_random_noise", "--user_defined_t_edit", str(t_edit), "--user_defined_t_addnoise",
                str(t_addnoise), "--synthetic_weights", "--seed", "1234"] + extra
        self.args, self.config = cli.parse_args_and_config(argv)
        self.runner = Asyrp(self.args, self.config, device=dev)
        self.drawn = {}
        orig = self.runner.random_noise_pairs

        def rec(*a, **k):
            self.drawn.update(orig(*a, **k))
            return self.drawn

        self.runner.random_noise_pairs = rec
        self.n_step = n_step

    def go(self):
        self.results = self.runner.run_test()
        return self.results

    def oracle_sd(self, blocks=()):
        from types import SimpleNamespace as NS
        ns = NS(model=NS(**{**CFG, "dropout": 0.0, "resamp_with_conv": True}), data=NS(image_size=CFG["image_size"]))
        m = modules.DDPM(ns)
        m.setattr_layers(len(blocks))
        synthetic.randomize_(m, seed=1234)
        for i, b in enumerate(blocks):
            getattr(m, f"layer_{i}").load_state_dict(b)
        return {k: v.clone() for k, v in m.state_dict().items()}

    def oracle(self, sd, x_T, **kw):
        seq, seq_next = make_sequences(999, self.n_step)
        kw.setdefault("t_edit", 500)
        return osmp.run_trajectory(lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k), x_T, betas=osmp.make_betas(),
                                   seq=seq, seq_next=seq_next, **kw)


def test_origin_pass_and_delta_interpolation(cuda_device, tmp_path, monkeypatch):
    """--save_x_origin + --delta_interpolation: grid rows = [origin DDIM] + one edit per interpolated coefficient
    (diffusion_latent.py:471-493, 741-755); every row against the oracle"""
    r = Run(tmp_path, monkeypatch, cuda_device, ["--train_delta_block", "--get_h_num", "1", "--manual_checkpoint_name",
                                                 "a_0.pth", "--save_x_origin", "--delta_interpolation", "--num_delta", "3",
                                                 "--min_delta", "0.0", "--max_delta", "1.5", "--hs_coeff_delta_h", "0.8"])
    blk = _delta_block_ckpt("checkpoint/a_0.pth", 11)
    rows = r.go()[("test", 0)]
    assert len(rows) == 4
    sd, x_T = r.oracle_sd([blk]), r.drawn["test"][0][2]
    _check(rows[0], r.oracle(sd, x_T, index=None, hs_coeff=(1.0,)), TOL, "origin pass")
    for row, v in zip(rows[1:], np.linspace(0.0, 1.5, 3)):
        _check(row, r.oracle(sd, x_T, index=0, hs_coeff=(1.0, v * 0.8)), TOL, f"interpolation {v}")
    assert os.path.exists(os.path.join(r.args.test_image_folder, f"test_0_0_ngen{r.n_step}.png"))


def test_multiple_attr_through_run_test(cuda_device, tmp_path, monkeypatch):
    """--multiple_attr 'a b' --multiple_hs_coeff '1.0 0.5': two DeltaBlocks, hs_coeff = (c0, c_k/sqrt(K)*scale) (:629-659)"""
    r = Run(tmp_path, monkeypatch, cuda_device, ["--train_delta_block", "--get_h_num", "2", "--manual_checkpoint_name",
                                                 "attribute_0.pth", "--multiple_attr", "a b", "--multiple_hs_coeff",
                                                 "1.0 0.5", "--hs_coeff_origin_h", "0.9"])
    b0, b1 = _delta_block_ckpt("checkpoint/a_0.pth", 11), _delta_block_ckpt("checkpoint/b_0.pth", 12)
    (row,) = r.go()[("test", 0)]
    sd, x_T = r.oracle_sd([b0, b1]), r.drawn["test"][0][2]
    s = 1.0 / 2 ** 0.5
    _check(row, r.oracle(sd, x_T, index=1, hs_coeff=(0.9, s * 1.0, s * 0.5)), TOL, "two attributes")


def _explicit_oracle(r, sd, x_T, dh_of_t, c0, t_edit=400, index=0):
    """the reference's edit loop with delta_h = dict[t] for t >= t_edit (diffusion_latent.py:507-520)"""
    seq, seq_next = make_sequences(999, r.n_step)
    betas = osmp.make_betas()
    x = x_T.clone()
    for i, j in zip(reversed(seq), reversed(seq_next)):
        x, _, _, _ = osmp.denoising_step(x, torch.ones(1) * i, torch.ones(1) * j,
                                         model=lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k), b=betas, index=index,
                                         t_edit=t_edit, hs_coeff=(c0, 1.0), delta_h=dh_of_t.get(i) if i >= t_edit else None)
    return x


def test_raw_delta_h_checkpoint_with_train_to_test_remap(cuda_device, tmp_path, monkeypatch):
    """--train_delta_h: {str(t): Δh} trained on n_train_step=3 steps, applied on n_test_step=6 (:678-690, 699-718):
    seq_train >= 400 = [499, 999]; seq_test_edit = [599, 799, 999] -> {599: Δh_499, 799: Δh_999, 999: Δh_999}"""
    r = Run(tmp_path, monkeypatch, cuda_device, ["--train_delta_h", "--manual_checkpoint_name", "raw_0.pth",
                                                 "--hs_coeff_origin_h", "0.6"], n_step=6, t_edit=400)
    r.args.n_train_step = 3
    g = torch.Generator().manual_seed(9)
    dh = {t: torch.randn(MID_C, 8, 8, generator=g) for t in (499, 999)}
    torch.save({str(t): v for t, v in dh.items()}, "checkpoint/raw_0.pth")
    (row,) = r.go()[("test", 0)]
    sd, x_T = r.oracle_sd([]), r.drawn["test"][0][2]
    expect = {599: dh[499][None], 799: dh[999][None], 999: dh[999][None]}
    _check(row, _explicit_oracle(r, sd, x_T, expect, 0.6), TOL, "raw Δh checkpoint, remapped")


def test_mean_delta_h_extraction_and_reuse(cuda_device, tmp_path, monkeypatch):
    """--num_mean_of_delta_hs 2 (:616-627, 757, 811-832): the DeltaBlock outputs of the first two training latents are
    averaged per edit timestep (+ key 0 = mean over timesteps), written to checkpoint_latent/<exp>_<n_test>_<N>.pth,
    and used as explicit Δh from the third image on; a second run loads the cached dict"""
    extra = ["--train_delta_block", "--get_h_num", "1", "--manual_checkpoint_name", "a_0.pth", "--num_mean_of_delta_hs", "2"]
    r = Run(tmp_path, monkeypatch, cuda_device, extra, n_train_img=3, n_test_img=0)
    blk = _delta_block_ckpt("checkpoint/a_0.pth", 11)
    res = r.go()
    exp_id = os.path.split(r.args.exp)[-1]
    path = f"checkpoint_latent/{exp_id}_{r.n_step}_2.pth"
    assert os.path.isfile(path)
    saved = torch.load(path, map_location="cpu", weights_only=True)
    sd = r.oracle_sd([blk])
    seq, seq_next = make_sequences(999, r.n_step)
    edit_ts = [t for t in seq if t >= 500]
    assert sorted(k for k in saved if saved[k] is not None) == sorted([0] + edit_ts)
    # oracle: per-step delta_h of the DeltaBlock along the first two trajectories
    acc = {t: 0.0 for t in edit_ts}
    betas = osmp.make_betas()
    for n in range(2):
        x = r.drawn["train"][n][2].clone()
        for i, j in zip(reversed(seq), reversed(seq_next)):
            x, _, d, _ = osmp.denoising_step(x, torch.ones(1) * i, torch.ones(1) * j,
                                             model=lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k), b=betas, index=0,
                                             t_edit=500, hs_coeff=(1.0, 1.0))
            if i >= 500:
                acc[i] = acc[i] + d / 2
    for t in edit_ts:
        _check(saved[t], acc[t], 4e-3, f"mean Δh at t={t}")
    _check(saved[0], sum(acc.values()) / len(acc), 4e-3, "global mean Δh")
    # third image: explicit-Δh branch with the mean (the engine's own dict, as the reference would use its own)
    x3 = r.drawn["train"][2][2]
    (row3,) = res[("train", 2)]
    _check(row3, _explicit_oracle(r, sd, x3, {t: saved[t] for t in edit_ts}, 1.0, t_edit=500), TOL, "edit with mean Δh")
    # second run: the cached dict is found and replaces the DeltaBlock path for every image
    r2 = Run(tmp_path, monkeypatch, cuda_device, extra, n_train_img=1, n_test_img=0)
    (row,) = r2.go()[("train", 0)]
    _check(row, _explicit_oracle(r2, sd, r2.drawn["train"][0][2], {t: saved[t] for t in edit_ts}, 1.0, t_edit=500), TOL,
           "edit with cached mean Δh")


def test_save_process_grids_and_target_image_id(cuda_device, tmp_path, monkeypatch):
    """--save_process_origin / --save_process_delta_h write one [x_t ; x0_t] grid per step (:485-491,523-527);
    --target_image_id restricts the run to the listed images (:768-783)"""
    r = Run(tmp_path, monkeypatch, cuda_device, ["--train_delta_block", "--get_h_num", "1", "--manual_checkpoint_name",
                                                 "a_0.pth", "--save_x_origin", "--save_process_origin",
                                                 "--save_process_delta_h", "--target_image_id", "1"], n_test_img=3)
    blk = _delta_block_ckpt("checkpoint/a_0.pth", 11)
    res = r.go()
    assert list(res) == [("test", 1)]
    seq, _ = make_sequences(999, r.n_step)
    folder = os.path.join(r.args.test_image_folder, "test_1_0")
    for t in seq:
        for kind in ("origin", "delta_h"):
            p = os.path.join(folder, f"{kind}_{t}.png")
            assert os.path.exists(p) and os.path.getsize(p) > 200, p
    sd = r.oracle_sd([blk])
    _check(res[("test", 1)][1], r.oracle(sd, r.drawn["test"][1][2], index=0, hs_coeff=(1.0, 1.0)), TOL, "target image 1")


def test_saved_random_noise_cache(cuda_device, tmp_path, monkeypatch):
    """--saved_random_noise (:1099-1167): latents + their plain generations are cached in the reference's file name and
    list-of-triples format and reused"""
    extra = ["--train_delta_block", "--get_h_num", "1", "--manual_checkpoint_name", "a_0.pth", "--saved_random_noise"]
    r = Run(tmp_path, monkeypatch, cuda_device, extra, n_test_img=2)
    _delta_block_ckpt("checkpoint/a_0.pth", 11)
    res1 = r.go()
    p = f"precomputed/CelebA_HQ_test_random_noise_nim2_ninv{r.n_step}_pairs.pth"
    assert os.path.exists(p)
    pairs = torch.load(p, map_location="cpu", weights_only=True)
    assert len(pairs) == 2 and all(len(t) == 3 and t[2].shape == (1, 3, 32, 32) for t in pairs)
    sd = r.oracle_sd([_delta_block_ckpt("checkpoint/a_0.pth", 11)])
    _check(pairs[0][0], r.oracle(sd, pairs[0][2], index=None, hs_coeff=(1.0,)), TOL, "cached generation")
    r2 = Run(tmp_path, monkeypatch, cuda_device, extra, n_test_img=2)
    res2 = r2.go()
    assert torch.equal(res1[("test", 1)][0], res2[("test", 1)][0])


def test_precompute_pairs_cache_roundtrip(cuda_device, tmp_path, monkeypatch):
    """real-image path (:951-1084): images -> DDIM inversion -> [x0, x_rec, x_T] cache file; the second call loads it"""
    from PIL import Image
    monkeypatch.chdir(tmp_path)
    os.makedirs("imgs", exist_ok=True)
    rng = np.random.RandomState(0)
    for i in range(2):
        Image.fromarray(rng.randint(0, 255, (32, 32, 3), dtype=np.uint8)).save(f"imgs/{i}.png")
    cfg = _write_cfg(os.path.join(str(tmp_path), "mini.yml"))
    argv = ["--run_test", "--config", cfg, "--exp", "./runs/x", "--n_test_img", "2", "--n_train_img", "2", "--bs_train", "2",
            "--n_inv_step", "4", "--custom_train_dataset_dir", "imgs", "--custom_test_dataset_dir", "imgs",
            "--user_defined_t_edit", "500", "--user_defined_t_addnoise", "0", "--synthetic_weights"]
    args, config = cli.parse_args_and_config(argv)
    runner = Asyrp(args, config, device=cuda_device)
    model = runner.load_pretrained_model().to(cuda_device)
    out = runner.precompute_pairs(model)
    p = "precomputed/CelebA_HQ_test_t999_nim2_ninv4_pairs.pth"
    assert os.path.exists(p) and not [f for f in os.listdir("precomputed") if ".tmp." in f]
    again = runner.precompute_pairs(model)
    for a, b in zip(out["test"], again["test"]):
        assert all(torch.equal(u, v) for u, v in zip(a, b)) and a[2].shape == (1, 3, 32, 32)
    # x_T against the oracle's inversion loop
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    seq, seq_next = make_sequences(999, 4)
    x = out["test"][0][0].clone()
    for i, j in zip(seq_next[1:], seq[1:]):
        x = osmp.denoising_step(x, torch.ones(1) * i, torch.ones(1) * j, model=lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k),
                                b=osmp.make_betas(), eta=0.0)[0]
    _check(out["test"][0][2], x, 1e-2, "inverted latent")


def _mini(dev, n_delta=1):
    from types import SimpleNamespace as NS
    ns = NS(model=NS(**{**CFG, "dropout": 0.0, "resamp_with_conv": True}), data=NS(image_size=CFG["image_size"]))
    m = modules.DDPM(ns)
    m.setattr_layers(n_delta)
    return synthetic.randomize_(m, 1234, "jittered").to(dev)


def test_ignore_timestep_graph_equals_step_loop(cuda_device):
    """--ignore_timesteps reaches the DeltaBlock inside the captured trajectory (diffusion_latent.py:515)"""
    m = _mini(cuda_device)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 32, 32, generator=g)
    betas = osmp.make_betas()
    seq, seq_next = make_sequences(999, 5)
    sch = Schedule(betas, seq, seq_next, t_edit=500, t_addnoise=0, hs_coeff=(1.0, 1.2), ignore_timestep=True)
    a = m.engine.sample(x.to(cuda_device), sch)
    xx = x.to(cuda_device)
    xo = x.clone()
    for i, j in zip(reversed(seq), reversed(seq_next)):
        xx = denoising_step(xx, torch.ones(2) * i, torch.ones(2) * j, models=m, b=betas, index=0, t_edit=500,
                            hs_coeff=(1.0, 1.2), ignore_timestep=True)[0]
        xo = osmp.denoising_step(xo, torch.ones(2) * i, torch.ones(2) * j, model=lambda *a_, **k: od.ddpm_forward(sd, CFG, *a_, **k),
                                 b=betas, index=0, t_edit=500, hs_coeff=(1.0, 1.2), ignore_timestep=True)[0]
    assert torch.equal(a, xx), "graph with ignore_timestep != step loop"
    _check(a, xo, TOL, "ignore_timestep trajectory")
    b = m.engine.sample(x.to(cuda_device), Schedule(betas, seq, seq_next, t_edit=500, t_addnoise=0, hs_coeff=(1.0, 1.2)))
    assert not torch.equal(a, b), "the timestep projection must matter"


def test_one_graph_serves_every_coefficient_tuple(cuda_device):
    """hs_coeff lives in device memory: changing it replays the same captured graph (no new capture), and the graph
    cache is bounded"""
    m = _mini(cuda_device)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 32, 32, generator=g)
    betas = osmp.make_betas()
    seq, seq_next = make_sequences(999, 4)
    eng = m.engine
    for c in (0.0, 0.5, 1.5):
        out = eng.sample(x.to(cuda_device), Schedule(betas, seq, seq_next, t_edit=500, hs_coeff=(1.0, c)))
        ref = osmp.run_trajectory(lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k), x, betas=betas, seq=seq,
                                  seq_next=seq_next, t_edit=500, index=0, hs_coeff=(1.0, c))
        _check(out, ref, TOL, f"coefficient {c}")
        assert len(eng.graphs) == 1
    for n in range(2, 9):
        s2, s2n = make_sequences(999, n)
        eng.sample(x.to(cuda_device), Schedule(betas, s2, s2n, t_edit=500, hs_coeff=(1.0, 1.0)))
    assert len(eng.graphs) <= eng.MAX_GRAPHS


@pytest.mark.parametrize("mode", ["dt_lambda", "ddpm"])
def test_sample_type_and_dt_lambda_in_graph(cuda_device, mode):
    """--sample_type ddpm and --dt_lambda are honoured by the captured trajectory
    (utils/diffusion_utils.py:74-82, 99-100; save_image forwards both, diffusion_latent.py:509,518)"""
    m = _mini(cuda_device)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 32, 32, generator=g)
    betas = osmp.make_betas()
    logvar = osmp.make_logvar(osmp.get_beta_schedule(beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    seq, seq_next = make_sequences(999, 5)
    kw = dict(sample_type="ddpm", logvars=logvar) if mode == "ddpm" else dict(dt_lambda=0.8)
    sch = Schedule(betas, seq, seq_next, t_edit=500, t_addnoise=0, hs_coeff=(1.0, 1.0), **kw)
    nz = torch.randn(sch.n_stochastic, 2, 3, 32, 32, generator=g) if sch.n_stochastic else None
    out = m.engine.sample(x.to(cuda_device), sch, noise=None if nz is None else nz.to(cuda_device))
    xo, zi = x.clone(), 0
    for i, j in zip(reversed(seq), reversed(seq_next)):
        z = None
        if mode == "ddpm":
            z, zi = nz[zi], zi + 1
        xo = osmp.denoising_step(xo, torch.ones(2) * i, torch.ones(2) * j, model=lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k),
                                 logvars=logvar, b=betas, index=0, t_edit=500, hs_coeff=(1.0, 1.0), noise=z,
                                 sampling_type="ddpm" if mode == "ddpm" else "ddim",
                                 dt_lambda=0.8 if mode == "dt_lambda" else 1)[0]
    _check(out, xo, TOL, mode)
    plain = m.engine.sample(x.to(cuda_device), Schedule(betas, seq, seq_next, t_edit=500, hs_coeff=(1.0, 1.0)))
    assert not torch.equal(out, plain)


def test_short_horizon_inversion_is_tight(cuda_device):
    """DDIM inversion + reconstruction over t_0 = 300 (--t_0 is a flag of the reference): here the maps stay O(1)
    (1/sqrt(alpha-bar_300) = 1.6, not 160), so the engine-vs-oracle error is bounded ABSOLUTELY"""
    import argparse
    from asyrp_official_b200.configs import load_config
    m = _mini(cuda_device)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(77)
    x0 = torch.tanh(torch.randn(2, 3, 32, 32, generator=g))
    r = Asyrp(argparse.Namespace(user_defined_t_edit=150, user_defined_t_addnoise=0, t_0=300, n_inv_step=10,
                                 sample_type="ddim"), load_config("celeba"), device=cuda_device)
    x_T, x_rec = r.invert_batch(m, x0)
    seq = [int(s + 1e-6) for s in list(np.linspace(0, 1, 10) * 300)]
    seq_next = [-1] + seq[:-1]
    fwd = lambda *a, **k: od.ddpm_forward(sd, CFG, *a, **k)  # noqa: E731
    betas = osmp.make_betas()
    x = x0.clone()
    for i, j in zip(seq_next[1:], seq[1:]):
        x = osmp.denoising_step(x, torch.ones(2) * i, torch.ones(2) * j, model=fwd, b=betas, eta=0.0)[0]
    e_T = (x_T - x).abs().max().item()
    xr = x.clone()
    for i, j in zip(reversed(seq), reversed(seq_next)):
        xr = osmp.denoising_step(xr, torch.ones(2) * i, torch.ones(2) * j, model=fwd, b=betas, eta=0.0)[0]
    e_r = (x_rec - xr).abs().max().item()
    assert x.abs().max() < 20 and xr.abs().max() < 20
    assert e_T <= 5e-3 and e_r <= 5e-3, f"inversion {e_T:.2e} (max|x_T| {x.abs().max():.2f}), reconstruction {e_r:.2e}"
