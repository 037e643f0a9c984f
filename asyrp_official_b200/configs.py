"""Config surface of the reference (configs/*.yml -> Namespace, main.py:231-233,311-319).

`load_config(x)` accepts a path to a YAML file in the reference's format, or the name of one of the reference's
shipped configs (celeba.yml, afhq.yml, ...): those differ only in data.dataset / data.category, so they are
generated here instead of being copied."""
import argparse
import os

import yaml

# name -> (dataset, category)   (reference configs/*.yml:1-4)
DATASETS = {
    "celeba": ("CelebA_HQ", "CelebA_HQ"),
    "celeba_dialog": ("CelebA_HQ_Dialog", "CelebA_HQ_Dialog"),
    "celeba_p2": ("CelebA_HQ_P2", "CelebA_HQ_P2"),
    "afhq": ("AFHQ", "AFHQ"),
    "ffhq": ("FFHQ", "FFHQ"),
    "metface": ("MetFACE", "MetFACE"),
    "church": ("LSUN", "church_outdoor"),
    "bedroom": ("LSUN", "bedroom"),
    "imagenet": ("IMAGENET", "IMAGENET"),
    "custom": ("CUSTOM", "CUSTOM"),
}


def builtin_config(name):
    dataset, category = DATASETS[name]
    return {
        "data": dict(dataset=dataset, category=category, image_size=256, channels=3, logit_transform=False,
                     uniform_dequantization=False, gaussian_dequantization=False, random_flip=True, rescaled=True,
                     num_workers=0),
        "model": dict(type="simple", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2,
                      attn_resolutions=[16], dropout=0.0, var_type="fixedsmall", ema_rate=0.999, ema=True,
                      resamp_with_conv=True),
        "diffusion": dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000),
        "sampling": dict(batch_size=4, last_only=True),
    }


def dict2namespace(config):
    ns = argparse.Namespace()
    for key, value in config.items():
        setattr(ns, key, dict2namespace(value) if isinstance(value, dict) else value)
    return ns


def load_config(path_or_name):
    if os.path.exists(path_or_name):
        with open(path_or_name) as f:
            return dict2namespace(yaml.safe_load(f))
    for cand in (path_or_name, os.path.join("configs", path_or_name)):
        if os.path.exists(cand):
            with open(cand) as f:
                return dict2namespace(yaml.safe_load(f))
    name = os.path.basename(path_or_name)
    name = name[:-4] if name.endswith(".yml") else name
    if name in DATASETS:
        return dict2namespace(builtin_config(name))
    raise FileNotFoundError(f"config {path_or_name!r}: not a file and not one of {sorted(DATASETS)}")
