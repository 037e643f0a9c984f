"""Seeded synthetic weights for benchmarks and smoke tests (no network => no pretrained checkpoints).

Every parameter is drawn from its own generator seeded by crc32(name) ^ seed, with the distributions torch's
Conv2d / Linear / GroupNorm constructors use — but never zeroed (the ADM family's zero_module() layers would make
a random-init network output exactly 0)."""
import math
import zlib

import torch


def randomize_(model, seed=1234, style="torch_default"):
    """in-place, deterministic, construction-order independent re-initialisation of all parameters"""
    sd = model.state_dict()
    fan = {k[:-7]: math.prod(v.shape[1:]) for k, v in sd.items() if k.endswith(".weight") and v.dim() >= 2}
    new = {}
    for name, p in sd.items():
        g = torch.Generator()
        g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        base = name.rsplit(".", 1)[0]
        shp = tuple(p.shape)
        if p.dim() >= 2:
            if style == "torch_default":
                t = (torch.rand(shp, generator=g) * 2 - 1) * (1.0 / math.sqrt(fan[base]))
            else:
                t = torch.randn(shp, generator=g) / math.sqrt(fan[base])
        elif base in fan:
            t = (torch.rand(shp, generator=g) * 2 - 1) * (1.0 / math.sqrt(fan[base]))
        elif name.endswith(".weight"):
            t = torch.ones(shp) if style == "torch_default" else 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = torch.zeros(shp) if style == "torch_default" else 0.1 * torch.randn(shp, generator=g)
        new[name] = t.float()
    model.load_state_dict(new)
    return model
