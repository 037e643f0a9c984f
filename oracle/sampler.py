"""CPU oracle for the Asyrp reverse process (sampler step + trajectory loop) — TEST INFRASTRUCTURE ONLY.

Restates `utils/diffusion_utils.py` (get_beta_schedule :5-9, extract :12-20, denoising_step :24-109) and the hot
loop of `Asyrp.save_image` (diffusion_latent.py:499-520) plus the schedule construction of `Asyrp.__init__`
(:41-61) and `run_test` (:570-574, :626, :659).
"""
import numpy as np
import torch


def get_beta_schedule(*, beta_start, beta_end, num_diffusion_timesteps):
    """linear beta schedule in float64  (utils/diffusion_utils.py:5-9)"""
    return np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)


def make_betas(beta_start=0.0001, beta_end=0.02, n=1000):
    """fp32 device betas as Asyrp.__init__ builds them  (diffusion_latent.py:41-46; configs/celeba.yml:27-31)"""
    return torch.from_numpy(get_beta_schedule(beta_start=beta_start, beta_end=beta_end,
                                              num_diffusion_timesteps=n)).float()


def make_logvar(betas64, var_type="fixedsmall"):
    """posterior log-variance table  (diffusion_latent.py:48-61)"""
    alphas = 1.0 - betas64
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post = betas64 * (1.0 - ac_prev) / (1.0 - ac)
    if var_type == "fixedlarge":
        return np.log(np.append(post[1], betas64[1:]))
    return np.log(np.maximum(post, 1e-20))


def extract(a, t, x_shape):
    """gather a[t] as fp32 and reshape to broadcast over x  (utils/diffusion_utils.py:12-20)"""
    out = torch.gather(torch.as_tensor(a, dtype=torch.float), 0, t.long())
    return out.reshape((t.shape[0],) + (1,) * (len(x_shape) - 1))


def make_sequences(t_0=999, n_step=40):
    """seq_test / seq_test_next  (diffusion_latent.py:570-574)"""
    seq = [int(s + 1e-6) for s in list(np.linspace(0, 1, n_step) * t_0)]
    return seq, [-1] + list(seq[:-1])


@torch.no_grad()
def denoising_step(xt, t, t_next, *, model, logvars=None, b, sampling_type="ddim", eta=0.0, learn_sigma=False,
                   index=None, t_edit=0, hs_coeff=(1.0,), delta_h=None, use_mask=False, dt_lambda=1,
                   ignore_timestep=False, dt_end=999, noise=None):
    """One reverse step  (utils/diffusion_utils.py:24-109).  `model` is a callable with the reference forward
    signature; `noise` replaces torch.randn_like(xt) (:79, :97) so that CPU and GPU runs share the draw."""
    et, et_mod, delta_h, middle_h = model(xt, t, index=index, t_edit=t_edit, hs_coeff=hs_coeff, delta_h=delta_h,
                                          ignore_timestep=ignore_timestep, use_mask=use_mask)
    if learn_sigma:  # :47-51
        et, logvar = torch.split(et, et.shape[1] // 2, dim=1)
        if index is not None:
            et_mod, _ = torch.split(et_mod, et_mod.shape[1] // 2, dim=1)
    else:
        logvar = extract(logvars, t, xt.shape) if logvars is not None else None
    bt = extract(b, t, xt.shape)
    at = extract((1.0 - b).cumprod(dim=0), t, xt.shape)  # fp32 cumprod, :67
    if t_next.sum() == -t_next.shape[0]:  # :68-69
        at_next = torch.ones_like(at)
    else:
        at_next = extract((1.0 - b).cumprod(dim=0), t_next, xt.shape)
    x0_t = None
    if sampling_type == "ddpm":  # :74-82
        weight = bt / torch.sqrt(1 - at)
        mean = 1 / torch.sqrt(1.0 - bt) * (xt - weight * et)
        z = noise if noise is not None else torch.randn_like(xt)
        mask = (1 - (t == 0).float()).reshape((xt.shape[0],) + (1,) * (xt.dim() - 1))
        xt_next = (mean + mask * torch.exp(0.5 * logvar) * z).float()
    else:  # 'ddim', :84-97
        e_for_x0 = et_mod if index is not None else et
        x0_t = (xt - e_for_x0 * (1 - at).sqrt()) / at.sqrt()
        if eta == 0:
            xt_next = at_next.sqrt() * x0_t + (1 - at_next).sqrt() * et
        else:
            c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
            c2 = ((1 - at_next) - c1 ** 2).sqrt()
            z = noise if noise is not None else torch.randn_like(xt)
            xt_next = at_next.sqrt() * x0_t + c2 * et + c1 * z
    if dt_lambda != 1 and t[0] >= dt_end:  # :99-100
        xt_next = at_next.sqrt() * x0_t + (1 - at_next).sqrt() * et * dt_lambda
    return xt_next, x0_t, delta_h, middle_h


@torch.no_grad()
def run_trajectory(model, x_T, *, betas, seq, seq_next, t_edit, t_addnoise=0, index=0, hs_coeff=(1.0, 1.0),
                   learn_sigma=False, noises=None, record=None, logvars=None):
    """The edit loop of Asyrp.save_image  (diffusion_latent.py:499-520): x = x_T; for (i, j) in reversed(seq, seq_next):
    eta = 1.0 when t < t_addnoise else 0.0 (:513); returns the final x.  `noises`: dict t -> pre-drawn N(0,1) tensor for
    the stochastic steps.  `record`: optional list receiving (t, x0_t) per step."""
    x = x_T.clone()
    bs = x.shape[0]
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = torch.ones(bs) * i
        t_next = torch.ones(bs) * j
        eta = 1.0 if i < t_addnoise else 0.0
        x, x0_t, _, _ = denoising_step(x, t, t_next, model=model, logvars=logvars, b=betas, eta=eta,
                                       learn_sigma=learn_sigma, index=index, t_edit=t_edit, hs_coeff=hs_coeff,
                                       noise=None if noises is None else noises.get(i))
        if record is not None:
            record.append((i, x0_t))
    return x
