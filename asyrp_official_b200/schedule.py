"""Host-side schedule of one Asyrp reverse trajectory.

Everything the reference decides per step with device->host synchronisations — `t[0] >= t_edit`
(models/ddpm/diffusion.py:510), `t[0] < t_addnoise` (diffusion_latent.py:513), `t_next.sum() == -bs`
(utils/diffusion_utils.py:68) and the alpha-bar lookups (:66-71) — is an integer/float decision made here, once,
before the trajectory graph is captured.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class Step:
    t: int
    t_next: int
    edit: bool     # t >= t_edit: DeltaBlock injection + second decoder pass
    at: float      # alpha-bar_t       (fp32 value)
    an: float      # alpha-bar_t_next  (1.0 when t_next == -1)
    c1: float      # coefficient of the injected noise (0 for eta = 0)
    c2: float      # coefficient of e_t: sqrt(1-an) for eta = 0
    kind: str = "ddim"     # 'ddpm': ancestral step (utils/diffusion_utils.py:74-82), uses bt / logvar / mask below
    bt: float = 0.0
    logvar: float = 0.0    # fixed-variance table entry (ignored when the UNet predicts the variance)
    mask: float = 1.0      # 0 at t == 0: no noise on the last ancestral step

    @property
    def stochastic(self):
        return self.kind == "ddpm" or self.c1 != 0.0


def make_sequences(t_0=999, n_step=40):
    """seq_test, seq_test_next  (diffusion_latent.py:570-574)"""
    seq = [int(s + 1e-6) for s in list(np.linspace(0, 1, n_step) * t_0)]
    return seq, [-1] + list(seq[:-1])


class Schedule:
    def __init__(self, betas, seq, seq_next, t_edit, t_addnoise=0, hs_coeff=(1.0, 1.0), edit=True, pairs=None,
                 sample_type="ddim", dt_lambda=1.0, dt_end=999, ignore_timestep=False, logvars=None):
        """betas: fp32 tensor (Asyrp.betas).  Coefficients are evaluated with the same fp32 torch expressions as
        utils/diffusion_utils.py:66-100 (cumprod in fp32 on the host) so that they are bit-identical to the CPU oracle.

        sample_type / dt_lambda / dt_end / ignore_timestep: what save_image forwards to denoising_step on every step
        (diffusion_latent.py:507-520; it never passes dt_end, i.e. 999).  logvars: Asyrp.logvar, needed by 'ddpm'
        sampling with a fixed-variance model."""
        if sample_type not in ("ddim", "ddpm"):
            raise ValueError(f"unknown sample_type {sample_type!r}")
        b = torch.as_tensor(betas, dtype=torch.float32).cpu()
        ac = (1.0 - b).cumprod(dim=0)
        lv = None if logvars is None else torch.as_tensor(logvars, dtype=torch.float32)
        steps: List[Step] = []
        for i, j in (pairs if pairs is not None else zip(reversed(seq), reversed(seq_next))):
            at = ac[i]
            an = torch.ones_like(at) if j == -1 else ac[j]
            eta = 1.0 if i < t_addnoise else 0.0
            if eta == 0.0:
                c1 = torch.zeros_like(at)
                c2 = (1 - an).sqrt()
            else:
                c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
                c2 = ((1 - an) - c1 ** 2).sqrt()
            if sample_type == "ddpm":
                if dt_lambda != 1 and i >= dt_end:
                    raise ValueError("sample_type='ddpm' with dt_lambda != 1 at t >= dt_end is undefined in the "
                                     "reference (x0_t is not computed on that branch, utils/diffusion_utils.py:99)")
                steps.append(Step(int(i), int(j), bool(edit and i >= t_edit), float(at), float(an), 0.0, 0.0, "ddpm",
                                  float(b[i]), 0.0 if lv is None else float(lv[i]), 0.0 if i == 0 else 1.0))
                continue
            if dt_lambda != 1 and i >= dt_end:  # :99-100 overrides whatever the eta branch produced
                c1 = torch.zeros_like(at)
                c2 = (1 - an).sqrt() * dt_lambda
            steps.append(Step(int(i), int(j), bool(edit and i >= t_edit), float(at), float(an), float(c1), float(c2)))
        self.steps = steps
        self.hs_coeff: Tuple[float, ...] = tuple(float(c) for c in hs_coeff)
        self.t_edit, self.t_addnoise = t_edit, t_addnoise
        self.ignore_timestep = bool(ignore_timestep)

    @classmethod
    def inversion(cls, betas, seq, seq_next):
        """DDIM inversion x_0 -> x_T: the deterministic step applied with t < t_next, plain UNet (index=None)
        (precompute_pairs, diffusion_latent.py:1032-1044: zip(seq_inv_next[1:], seq_inv[1:]))"""
        return cls(betas, None, None, t_edit=10 ** 9, t_addnoise=0, hs_coeff=(1.0,), edit=False,
                   pairs=list(zip(seq_next[1:], seq[1:])))

    def key(self):
        """identifies the captured graph: the DeltaBlock coefficients are device-side parameters, not part of it"""
        return (tuple(self.steps), self.ignore_timestep, len(self.hs_coeff))

    @property
    def n_stochastic(self):
        return sum(1 for s in self.steps if s.stochastic)

    @property
    def n_edit(self):
        return sum(1 for s in self.steps if s.edit)
