"""CPU oracle for the DDPM++ UNet family (CelebA-HQ / LSUN) — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A functional, plain-PyTorch fp32 restatement of `models/ddpm/diffusion.py` of the reference, operating on
a state dict with the reference's parameter names.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  Pinned against the reference's own modules by
tests/golden/make_golden.py (fixtures under tests/golden/) and tests/test_oracle.py.

Reference citations are `models/ddpm/diffusion.py:<line>` unless noted.
"""
import math

import torch
import torch.nn.functional as F


def swish(x):
    """x * sigmoid(x)  (:63-65)"""
    return x * torch.sigmoid(x)


def group_norm(sd, prefix, x, eps=1e-6):
    """GroupNorm(32 groups, eps=1e-6, affine)  (Normalize, :68-69)"""
    return F.group_norm(x, 32, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def conv(sd, prefix, x, stride=1, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], stride=stride, padding=padding)


def timestep_embedding(t, dim):
    """[sin | cos] table with frequencies exp(-i*ln(1e4)/(half-1))  (get_timestep_embedding, :42-60)"""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -step)
    ang = t.float()[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def temb_mlp(sd, t, ch):
    """temb.dense[0] -> swish -> temb.dense[1]  (:477-480, get_temb :464-470)"""
    e = timestep_embedding(t, ch)
    e = F.linear(e, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    return F.linear(swish(e), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])


def resnet_block(sd, p, x, temb):
    """norm1-swish-conv1, + temb_proj(swish(temb)), norm2-swish-(dropout p=0)-conv2, 1x1 shortcut when the
    channel count changes, residual add  (ResnetBlock.forward, :151-170)"""
    h = conv(sd, p + ".conv1", swish(group_norm(sd, p + ".norm1", x)), padding=1)
    h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    h = conv(sd, p + ".conv2", swish(group_norm(sd, p + ".norm2", h)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(sd, p + ".nin_shortcut", x)
    elif (p + ".conv_shortcut.weight") in sd:
        x = conv(sd, p + ".conv_shortcut", x, padding=1)
    return x + h


def attn_block(sd, p, x):
    """single-head attention over H*W positions, logits scaled by C^-0.5 after q k^T  (AttnBlock.forward, :200-225)"""
    hn = group_norm(sd, p + ".norm", x)
    q, k, v = conv(sd, p + ".q", hn), conv(sd, p + ".k", hn), conv(sd, p + ".v", hn)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    o = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + conv(sd, p + ".proj_out", o)


def downsample(sd, p, x):
    """pad right/bottom by one, 3x3 stride-2 conv  (Downsample.forward with_conv, :103-108)"""
    return conv(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), stride=2)


def upsample(sd, p, x):
    """nearest x2 then 3x3 conv  (Upsample.forward with_conv, :83-88)"""
    return conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"), padding=1)


def delta_block(sd, p, x, temb):
    """conv1x1 -> (+temb_proj(swish(temb))) -> norm2 -> swish -> conv1x1  (DeltaBlock.forward, :251-263)"""
    h = conv(sd, p + ".conv1", x)
    if temb is not None:
        h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    return conv(sd, p + ".conv2", swish(group_norm(sd, p + ".norm2", h)))


def slerp(t, v0, v1):
    """spherical interpolation on flattened per-sample vectors  (slerp, :6-40)"""
    shp = v0.shape
    n0 = v0 / torch.norm(v0.reshape(shp[0], -1), dim=1)[:, None, None, None]
    n1 = v1 / torch.norm(v1.reshape(shp[0], -1), dim=1)[:, None, None, None]
    dot = torch.sum(n0.reshape(shp[0], -1) * n1.reshape(shp[0], -1), dim=1)
    th0 = torch.acos(dot)
    tht = th0 * t
    s0 = (torch.sin(th0 - tht) / torch.sin(th0))[:, None, None, None]
    s1 = (torch.sin(tht) / torch.sin(th0))[:, None, None, None]
    return s0 * v0 + s1 * v1


def _decoder(sd, cfg, h, hs, temb):
    """up path + norm_out/swish/conv_out, reading the skip stack from the top without consuming it
    (:544-559 peeks with hs_index, :564-578 pops; both visit the same tensors in the same order)"""
    nres = len(cfg["ch_mult"])
    idx = -1
    for lvl in reversed(range(nres)):
        for blk in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, f"up.{lvl}.block.{blk}", torch.cat([h, hs[idx]], dim=1), temb)
            idx -= 1
            if f"up.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"up.{lvl}.attn.{blk}", h)
        if lvl != 0:
            h = upsample(sd, f"up.{lvl}.upsample", h)
    return conv(sd, "conv_out", swish(group_norm(sd, "norm_out", h)), padding=1)


@torch.no_grad()
def ddpm_forward(sd, cfg, x, t, index=None, t_edit=400, hs_coeff=(1.0, 1.0), delta_h=None,
                 ignore_timestep=False, use_mask=False):
    """DDPM.forward  (:473-580).  cfg: dict(ch, ch_mult, num_res_blocks, image_size).

    Returns (et, et_modified | None, delta_h | None, middle_h) exactly like the reference."""
    assert x.shape[2] == x.shape[3] == cfg["image_size"]
    temb = temb_mlp(sd, t, cfg["ch"])
    nres = len(cfg["ch_mult"])
    hs = [conv(sd, "conv_in", x, padding=1)]
    for lvl in range(nres):
        for blk in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, f"down.{lvl}.block.{blk}", hs[-1], temb)
            if f"down.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"down.{lvl}.attn.{blk}", h)
            hs.append(h)
        if lvl != nres - 1:
            hs.append(downsample(sd, f"down.{lvl}.downsample", hs[-1]))
    h = resnet_block(sd, "mid.block_1", hs[-1], temb)
    h = attn_block(sd, "mid.attn_1", h)
    h = resnet_block(sd, "mid.block_2", h, temb)
    middle_h = h

    et_mod = None
    if index is not None:
        if t[0] >= t_edit:  # :510
            if delta_h is None:  # Asyrp: h2 = c0*h + sum_i c_{i+1} * layer_i(h, temb)   (:512-516)
                h2 = h * hs_coeff[0]
                for i in range(index + 1):
                    delta_h = delta_block(sd, f"layer_{i}", h, None if ignore_timestep else temb)
                    h2 = h2 + delta_h * hs_coeff[i + 1]
            elif use_mask:  # (:519-527)
                mask = torch.zeros_like(h)
                mask[:, :, 4:-1, 3:5] = 1.0
                h2 = slerp(1 - hs_coeff[0], h * mask, delta_h * mask) + (1 - mask) * h
            else:  # explicit delta_h, norm-matched slerp  (:529-539)
                hn = torch.norm(h.reshape(h.shape[0], -1), dim=1)[:, None, None, None]
                dn = torch.norm(delta_h.reshape(h.shape[0], -1), dim=1)[:, None, None, None]
                h2 = slerp(1.0 - hs_coeff[0], h, hn * delta_h / dn)
        else:
            h2 = h  # :541-542
        et_mod = _decoder(sd, cfg, h2, hs, temb)
    et = _decoder(sd, cfg, h, hs, temb)
    return et, et_mod, delta_h, middle_h


# ---------------------------------------------------------------------------------------------------
# Parameter inventory (names and shapes of DDPM.state_dict(), :327-430 and setattr_layers :433-444)
# ---------------------------------------------------------------------------------------------------
def ddpm_param_shapes(cfg, n_delta_blocks=0):
    ch, mult, nrb = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"]
    in_ch, out_ch = cfg.get("in_channels", 3), cfg.get("out_ch", 3)
    res, attn_res = cfg["image_size"], cfg["attn_resolutions"]
    tch = 4 * ch
    shapes = {}

    def lin(p, i, o):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i), (o,)

    def cv(p, i, o, k):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i, k, k), (o,)

    def gn(p, c):
        shapes[p + ".weight"], shapes[p + ".bias"] = (c,), (c,)

    def resblock(p, i, o):
        gn(p + ".norm1", i); cv(p + ".conv1", i, o, 3); lin(p + ".temb_proj", tch, o)
        gn(p + ".norm2", o); cv(p + ".conv2", o, o, 3)
        if i != o:
            cv(p + ".nin_shortcut", i, o, 1)

    def attn(p, c):
        gn(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            cv(p + "." + n, c, c, 1)

    lin("temb.dense.0", ch, tch); lin("temb.dense.1", tch, tch)
    cv("conv_in", in_ch, ch, 3)
    in_mult = (1,) + mult
    cur = res
    block_in = ch
    for lvl in range(len(mult)):
        block_in, block_out = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(nrb):
            resblock(f"down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
            if cur in attn_res:
                attn(f"down.{lvl}.attn.{b}", block_in)
        if lvl != len(mult) - 1:
            cv(f"down.{lvl}.downsample.conv", block_in, block_in, 3)
            cur //= 2
    resblock("mid.block_1", block_in, block_in); attn("mid.attn_1", block_in); resblock("mid.block_2", block_in, block_in)
    mid_ch = block_in
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        skip_in = ch * mult[lvl]
        for b in range(nrb + 1):
            if b == nrb:
                skip_in = ch * in_mult[lvl]
            resblock(f"up.{lvl}.block.{b}", block_in + skip_in, block_out)
            block_in = block_out
            if cur in attn_res:
                attn(f"up.{lvl}.attn.{b}", block_in)
        if lvl != 0:
            cv(f"up.{lvl}.upsample.conv", block_in, block_in, 3)
            cur *= 2
    gn("norm_out", block_in); cv("conv_out", block_in, out_ch, 3)
    for i in range(n_delta_blocks):
        p = f"layer_{i}"
        cv(p + ".conv1", mid_ch, mid_ch, 1); lin(p + ".temb_proj", tch, mid_ch)
        gn(p + ".norm2", mid_ch); cv(p + ".conv2", mid_ch, mid_ch, 1)
    return shapes


# configs/celeba.yml:13-25 (identical model block in every configs/*.yml)
CELEBA_CFG = dict(ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=[16],
                  in_channels=3, image_size=256)
# reduced configuration for fast tests (same code path, three resolution levels)
MINI_CFG = dict(ch=64, out_ch=3, ch_mult=(1, 2, 4), num_res_blocks=1, attn_resolutions=[16], in_channels=3,
                image_size=32)
