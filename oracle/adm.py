"""CPU oracle for the iDDPM / ADM UNet family (AFHQ, FFHQ, MetFaces, CelebA-HQ-P2, ImageNet-256) — TEST
INFRASTRUCTURE, NOT A PRODUCT PATH.

Functional fp32 restatement of `models/improved_ddpm/unet.py` (textually the same network as
`models/guided_diffusion/unet.py`), operating on a state dict with the reference's parameter names.
Citations are `models/improved_ddpm/unet.py:<line>` unless noted (nn.py = models/improved_ddpm/nn.py).
"""
import math

import torch
import torch.nn.functional as F

from .ddpm import slerp

# hyper-parameter dictionaries: improved_ddpm/script_util.py:5-42, guided_diffusion/script_util.py:10-46
AFHQ_HP = dict(image_size=256, model_channels=128, num_res_blocks=1, attention_resolutions=(16,),
               channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64, out_channels=6, in_channels=3)
IMAGENET_HP = dict(image_size=256, model_channels=256, num_res_blocks=2, attention_resolutions=(32, 16, 8),
                   channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64, out_channels=6, in_channels=3)
# reduced configuration for fast tests
MINI_HP = dict(image_size=32, model_channels=64, num_res_blocks=1, attention_resolutions=(16, 8),
               channel_mult=(1, 2, 4), num_head_channels=64, out_channels=6, in_channels=3)


def timestep_embedding(t, dim, max_period=10000):
    """[cos | sin] with frequencies exp(-ln(max_period)*i/half)  (nn.py:103-121)"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def gn32(sd, p, x):
    """GroupNorm32(32, C): fp32 group norm, eps 1e-5  (nn.py:17-19,93-100)"""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], 1e-5).type(x.dtype)


def _conv(sd, p, x, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def res_block(sd, p, x, emb, up=False, down=False):
    """ResBlock._forward with use_scale_shift_norm=True  (:278-298).
    in_layers = GN, SiLU, conv3x3; up/down resample h and x between SiLU and the conv (:279-284);
    emb_layers = SiLU, Linear(->2C); out = GN(h)*(1+scale)+shift, SiLU, dropout(p=0), conv3x3; skip 1x1 if C changes."""
    h = F.silu(gn32(sd, p + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = _conv(sd, p + ".in_layers.2", h, padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[:, :, None, None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = gn32(sd, p + ".out_layers.0", h) * (1 + scale) + shift
    h = _conv(sd, p + ".out_layers.3", F.silu(h), padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def attention_block(sd, p, x, head_ch=64):
    """AttentionBlock._forward + QKVAttentionLegacy  (:341-347, :379-396): qkv 1x1 conv over flattened positions,
    channels grouped [head][q|k|v][ch]; q and k scaled by ch^-1/4; softmax in fp32; zero-init proj_out; residual."""
    b, c, *sp = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(gn32(sd, p + ".norm", xf), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    heads = c // head_ch
    q, k, v = qkv.reshape(b * heads, head_ch * 3, -1).split(head_ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(head_ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, xf.shape[-1])
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, *sp)


def delta_block(sd, p, x, emb):
    """ADM DeltaBlock (use_scale_shift_norm=False): GN, SiLU, conv1x1, (+Linear(SiLU(emb))), GN, SiLU, conv1x1 (:837-853)"""
    h = _conv(sd, p + ".in_layers.2", F.silu(gn32(sd, p + ".in_layers.0", x)))
    if emb is not None:
        h = h + F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[:, :, None, None]
    return _conv(sd, p + ".out_layers.3", F.silu(gn32(sd, p + ".out_layers.0", h)))


def adm_layout(hp):
    """Block structure of UNetModel.__init__ (:513-658) for resblock_updown=True, use_scale_shift_norm=True,
    num_head_channels=64: returns (input_blocks, middle, output_blocks) as lists of per-block layer lists; a layer is
    ('conv_in', cin, cout) | ('res', cin, cout, 'none'|'up'|'down') | ('attn', c)."""
    mc, mult, nrb = hp["model_channels"], hp["channel_mult"], hp["num_res_blocks"]
    attn_ds = tuple(hp["image_size"] // r for r in hp["attention_resolutions"])
    ch = int(mult[0] * mc)
    inputs = [[("conv_in", hp["in_channels"], ch)]]
    chans = [ch]
    ds = 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, int(m * mc), "none")]
            ch = int(m * mc)
            if ds in attn_ds:
                layers.append(("attn", ch))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("res", ch, ch, "down")])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch, "none"), ("attn", ch), ("res", ch, ch, "none")]
    outputs = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * m), "none")]
            ch = int(mc * m)
            if ds in attn_ds:
                layers.append(("attn", ch))
            if level and i == nrb:
                layers.append(("res", ch, ch, "up"))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def _run_layers(sd, prefix, layers, h, emb, head_ch):
    for j, layer in enumerate(layers):
        p = f"{prefix}.{j}"
        if layer[0] == "conv_in":
            h = _conv(sd, p, h, padding=1)
        elif layer[0] == "res":
            h = res_block(sd, p, h, emb, up=layer[3] == "up", down=layer[3] == "down")
        else:
            h = attention_block(sd, p, h, head_ch)
    return h


@torch.no_grad()
def adm_forward(sd, hp, x, t, y=None, index=None, t_edit=400, hs_coeff=(1.0, 1.0), delta_h=None,
                ignore_timestep=False, use_mask=False):
    """UNetModel.forward  (:676-752); class conditioning is disabled in the reference (:685-688)."""
    inputs, middle, outputs = adm_layout(hp)
    head_ch = hp["num_head_channels"]
    emb = timestep_embedding(t, hp["model_channels"])
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x
    for i, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb, head_ch)
        hs.append(h)
    h = _run_layers(sd, "middle_block", middle, h, emb, head_ch)
    middle_h = h

    def decoder(hh):
        idx = -1
        for i, layers in enumerate(outputs):
            hh = _run_layers(sd, f"output_blocks.{i}", layers, torch.cat([hh, hs[idx]], dim=1), emb, head_ch)
            idx -= 1
        return _conv(sd, "out.2", F.silu(gn32(sd, "out.0", hh)), padding=1)

    out_mod = None
    if index is not None:
        if t[0] >= t_edit:  # :699
            if delta_h is None:  # :701-706
                h2 = h * hs_coeff[0]
                for i in range(index + 1):
                    delta_h = delta_block(sd, f"layer_{i}", h, None if ignore_timestep else emb)
                    h2 = h2 + delta_h * hs_coeff[i + 1]
            elif use_mask:  # :709-717
                mask = torch.zeros_like(h)
                mask[:, :, 4:-1, 3:5] = 1.0
                h2 = slerp(1 - hs_coeff[0], h * mask, delta_h * mask) + (1 - mask) * h
            else:  # :720-730
                hn = torch.norm(h.reshape(h.shape[0], -1), dim=1)[:, None, None, None]
                dn = torch.norm(delta_h.reshape(h.shape[0], -1), dim=1)[:, None, None, None]
                h2 = slerp(1.0 - hs_coeff[0], h, hn * delta_h / dn)
        else:
            h2 = h
        out_mod = decoder(h2)
    out = decoder(h)
    return out, out_mod, delta_h, middle_h


def adm_param_shapes(hp, n_delta_blocks=0):
    """Names and shapes of UNetModel.state_dict() (+ layer_i DeltaBlocks, setattr_layers :756-773).
    The unused class-embedding table of the ImageNet checkpoint (label_emb, :519-520) is not included."""
    inputs, middle, outputs = adm_layout(hp)
    ted = 4 * hp["model_channels"]
    shapes = {}

    def cv(p, i, o, k):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i, k, k), (o,)

    def gn(p, c):
        shapes[p + ".weight"], shapes[p + ".bias"] = (c,), (c,)

    def lin(p, i, o):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i), (o,)

    lin("time_embed.0", hp["model_channels"], ted); lin("time_embed.2", ted, ted)

    def add(prefix, layers):
        for j, layer in enumerate(layers):
            p = f"{prefix}.{j}"
            if layer[0] == "conv_in":
                cv(p, layer[1], layer[2], 3)
            elif layer[0] == "res":
                _, ci, co, _ = layer
                gn(p + ".in_layers.0", ci); cv(p + ".in_layers.2", ci, co, 3)
                lin(p + ".emb_layers.1", ted, 2 * co)
                gn(p + ".out_layers.0", co); cv(p + ".out_layers.3", co, co, 3)
                if ci != co:
                    cv(p + ".skip_connection", ci, co, 1)
            else:
                c = layer[1]
                gn(p + ".norm", c)
                shapes[p + ".qkv.weight"], shapes[p + ".qkv.bias"] = (3 * c, c, 1), (3 * c,)
                shapes[p + ".proj_out.weight"], shapes[p + ".proj_out.bias"] = (c, c, 1), (c,)

    for i, layers in enumerate(inputs):
        add(f"input_blocks.{i}", layers)
    add("middle_block", middle)
    for i, layers in enumerate(outputs):
        add(f"output_blocks.{i}", layers)
    c_out = int(hp["channel_mult"][0] * hp["model_channels"])
    gn("out.0", c_out); cv("out.2", c_out, hp["out_channels"], 3)
    mid = middle[0][1]
    for i in range(n_delta_blocks):
        p = f"layer_{i}"
        gn(p + ".in_layers.0", mid); cv(p + ".in_layers.2", mid, mid, 1); lin(p + ".emb_layers.1", ted, mid)
        gn(p + ".out_layers.0", mid); cv(p + ".out_layers.3", mid, mid, 1)
    return shapes
