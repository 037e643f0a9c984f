"""Deterministic synthetic weights for parity tests and benchmarks — TEST INFRASTRUCTURE ONLY.

There is no network, hence no pretrained UNet checkpoints; parity is established on seeded random
weights loaded identically into the reference, the oracle and the engine.  Every tensor is drawn from its own
generator seeded by crc32(name) ^ seed, so the values do not depend on module construction order or on which
subset of parameters is requested.

style "torch_default": the distributions torch.nn.Conv2d / Linear / GroupNorm use at construction
(U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases, norm weight 1 / bias 0) — but WITHOUT the
zero_module() zeroing of the ADM family (improved_ddpm/unet.py:252-254,336,657), which would make the network
output identically 0 and the parity test vacuous.
style "jittered": additionally perturbs norm scales/offsets and uses N(0, 1/fan_in) weights, so that affine
parameters are exercised.
"""
import math
import zlib

import torch


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(shapes, seed=1234, style="torch_default"):
    sd = {}
    fan_in_of = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) >= 2:
            fan_in_of[name[:-7]] = int(torch.tensor(shp[1:]).prod().item())
    for name, shp in shapes.items():
        g = _gen(name, seed)
        base = name.rsplit(".", 1)[0]
        if len(shp) >= 2:  # conv / linear weight
            fan = fan_in_of[base]
            if style == "torch_default":
                bound = 1.0 / math.sqrt(fan)
                t = (torch.rand(shp, generator=g) * 2 - 1) * bound
            else:
                t = torch.randn(shp, generator=g) / math.sqrt(fan)
        elif base in fan_in_of:  # bias of a conv / linear
            bound = 1.0 / math.sqrt(fan_in_of[base])
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif name.endswith(".weight"):  # norm scale
            t = torch.ones(shp) if style == "torch_default" else 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:  # norm offset
            t = torch.zeros(shp) if style == "torch_default" else 0.1 * torch.randn(shp, generator=g)
        sd[name] = t.float().contiguous()
    return sd


def synth_noise(shape, seed=1234):
    """x_T ~ N(0, 1) from a seeded CPU generator (the reference seeds with 1234, main.py:145)"""
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(shape, generator=g)


def synth_image(shape, seed=77):
    """a smooth synthetic 'photograph' in [-1, 1]: three octaves of seeded noise, bilinearly upsampled, through tanh
    (stand-in for the dataset images precompute_pairs inverts, diffusion_latent.py:905-933; there are no datasets here)"""
    import torch.nn.functional as F
    g = torch.Generator()
    g.manual_seed(seed)
    n, c, h, w = shape
    img = torch.zeros(shape)
    for div, amp in ((32, 1.0), (8, 0.5), (2, 0.15)):
        z = torch.randn(n, c, max(2, h // div), max(2, w // div), generator=g)
        img = img + amp * F.interpolate(z, size=(h, w), mode="bilinear", align_corners=True)
    return torch.tanh(img)
