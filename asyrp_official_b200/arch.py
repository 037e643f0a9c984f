"""Architecture descriptors of the three UNet families of the Asyrp path.

One neutral description — a list of blocks per stage — drives both the parameter inventory of the drop-in
nn.Module mirrors (modules.py: same state-dict keys as the reference) and the kernel plan (engine.py).

Families:
  'ddpm'  DDPM++ UNet, CelebA-HQ / LSUN      reference models/ddpm/diffusion.py:327-430
  'adm'   iDDPM / ADM UNet (AFHQ, FFHQ, MetFaces, CelebA-HQ-P2, ImageNet)
          reference models/improved_ddpm/unet.py:469-658 == models/guided_diffusion/unet.py:468-657
"""
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class Res:
    """residual block; cin is the (possibly concatenated) input width; resample in {'none','up','down'} (ADM only)"""
    name: str
    cin: int
    cout: int
    resample: str = "none"
    split: Tuple[int, ...] = ()   # channel counts of the concatenated inputs (decoder: (h, skip)); () = single input


@dataclass
class Attn:
    name: str
    c: int


@dataclass
class Resample:
    """DDPM Downsample (pad (0,1,0,1) + 3x3 s2, :96-108) / Upsample (nearest x2 + 3x3, :77-88)"""
    name: str
    c: int
    kind: str  # 'down' | 'up'


@dataclass
class Arch:
    family: str
    image_size: int
    in_ch: int
    out_ch: int
    base_ch: int                  # ch / model_channels
    temb_ch: int
    conv_in: str
    # encoder: list of stages; each stage is the list of layers producing ONE skip tensor
    enc: List[list] = field(default_factory=list)
    mid: list = field(default_factory=list)
    # decoder: list of stages; each consumes one skip tensor (concat) at its first layer
    dec: List[list] = field(default_factory=list)
    norm_out: str = ""
    conv_out: str = ""
    mid_ch: int = 0
    gn_eps: float = 1e-6
    head_ch: int = 0              # 0: single head over all channels (DDPM)
    learn_sigma: bool = False
    temb_names: Tuple[str, str] = ("", "")


def ddpm_arch(ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=(16,),
              in_channels=3, image_size=256, **_ignored) -> Arch:
    mult = tuple(ch_mult)
    a = Arch("ddpm", image_size, in_channels, out_ch, ch, 4 * ch, "conv_in", gn_eps=1e-6, head_ch=0,
             learn_sigma=False, temb_names=("temb.dense.0", "temb.dense.1"))
    in_mult = (1,) + mult
    cur = image_size
    a.enc.append([])  # conv_in alone produces hs[0]
    block_in = ch
    for lvl in range(len(mult)):
        block_in, block_out = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(num_res_blocks):
            stage = [Res(f"down.{lvl}.block.{b}", block_in, block_out)]
            block_in = block_out
            if cur in attn_resolutions:
                stage.append(Attn(f"down.{lvl}.attn.{b}", block_in))
            a.enc.append(stage)
        if lvl != len(mult) - 1:
            a.enc.append([Resample(f"down.{lvl}.downsample", block_in, "down")])
            cur //= 2
    a.mid = [Res("mid.block_1", block_in, block_in), Attn("mid.attn_1", block_in), Res("mid.block_2", block_in, block_in)]
    a.mid_ch = block_in
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        skip_in = ch * mult[lvl]
        for b in range(num_res_blocks + 1):
            if b == num_res_blocks:
                skip_in = ch * in_mult[lvl]
            stage = [Res(f"up.{lvl}.block.{b}", block_in + skip_in, block_out, split=(block_in, skip_in))]
            block_in = block_out
            if cur in attn_resolutions:
                stage.append(Attn(f"up.{lvl}.attn.{b}", block_in))
            if b == num_res_blocks and lvl != 0:
                stage.append(Resample(f"up.{lvl}.upsample", block_in, "up"))
                cur *= 2
            a.dec.append(stage)
    a.norm_out, a.conv_out = "norm_out", "conv_out"
    return a


def adm_arch(image_size=256, model_channels=128, num_res_blocks=1, attention_resolutions=(16,),
             channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64, out_channels=6, in_channels=3, **_ignored) -> Arch:
    """resblock_updown=True, use_scale_shift_norm=True, head width 64 — the only variant the reference instantiates
    (improved_ddpm/script_util.py:5-42, guided_diffusion/script_util.py:10-46)."""
    mc, mult = model_channels, tuple(channel_mult)
    attn_ds = tuple(image_size // int(r) for r in attention_resolutions)
    a = Arch("adm", image_size, in_channels, out_channels, mc, 4 * mc, "input_blocks.0.0", gn_eps=1e-5,
             head_ch=num_head_channels, learn_sigma=(out_channels == 2 * in_channels),
             temb_names=("time_embed.0", "time_embed.2"))
    ch = int(mult[0] * mc)
    a.enc.append([])
    chans = [ch]
    ds = 1
    idx = 1
    for level, m in enumerate(mult):
        for _ in range(num_res_blocks):
            stage = [Res(f"input_blocks.{idx}.0", ch, int(m * mc))]
            ch = int(m * mc)
            if ds in attn_ds:
                stage.append(Attn(f"input_blocks.{idx}.1", ch))
            a.enc.append(stage)
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            a.enc.append([Res(f"input_blocks.{idx}.0", ch, ch, "down")])
            chans.append(ch)
            idx += 1
            ds *= 2
    a.mid = [Res("middle_block.0", ch, ch), Attn("middle_block.1", ch), Res("middle_block.2", ch, ch)]
    a.mid_ch = ch
    oidx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            j = 0
            stage = [Res(f"output_blocks.{oidx}.{j}", ch + ich, int(mc * m), split=(ch, ich))]
            ch = int(mc * m)
            j += 1
            if ds in attn_ds:
                stage.append(Attn(f"output_blocks.{oidx}.{j}", ch))
                j += 1
            if level and i == num_res_blocks:
                stage.append(Res(f"output_blocks.{oidx}.{j}", ch, ch, "up"))
                ds //= 2
            a.dec.append(stage)
            oidx += 1
    a.norm_out, a.conv_out = "out.0", "out.2"
    return a


# hard-coded hyper-parameter sets of the reference
AFHQ_HP = dict(image_size=256, model_channels=128, num_res_blocks=1, attention_resolutions=(16,),
               channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64, out_channels=6)    # improved_ddpm/script_util.py:5-22
IMAGENET_HP = dict(image_size=256, model_channels=256, num_res_blocks=2, attention_resolutions=(32, 16, 8),
                   channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64, out_channels=6)  # :25-42
METFACE_HP = dict(AFHQ_HP)       # guided_diffusion/script_util.py:10-27 (same widths; dropout/heads irrelevant at inference)
CELEBA_HQ_P2_HP = dict(AFHQ_HP)  # guided_diffusion/script_util.py:29-46


def param_shapes(a: Arch, n_delta_blocks=0):
    """state-dict names -> shapes, in the reference's naming"""
    shapes = {}

    def cv(p, i, o, k):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i, k, k), (o,)

    def gn(p, c):
        shapes[p + ".weight"], shapes[p + ".bias"] = (c,), (c,)

    def lin(p, i, o):
        shapes[p + ".weight"], shapes[p + ".bias"] = (o, i), (o,)

    lin(a.temb_names[0], a.base_ch, a.temb_ch)
    lin(a.temb_names[1], a.temb_ch, a.temb_ch)
    first_ch = a.enc[1][0].cin
    cv(a.conv_in, a.in_ch, first_ch, 3)

    def add(layer):
        p = layer.name
        if isinstance(layer, Res):
            if a.family == "ddpm":
                gn(p + ".norm1", layer.cin); cv(p + ".conv1", layer.cin, layer.cout, 3)
                lin(p + ".temb_proj", a.temb_ch, layer.cout)
                gn(p + ".norm2", layer.cout); cv(p + ".conv2", layer.cout, layer.cout, 3)
                if layer.cin != layer.cout:
                    cv(p + ".nin_shortcut", layer.cin, layer.cout, 1)
            else:
                gn(p + ".in_layers.0", layer.cin); cv(p + ".in_layers.2", layer.cin, layer.cout, 3)
                lin(p + ".emb_layers.1", a.temb_ch, 2 * layer.cout)
                gn(p + ".out_layers.0", layer.cout); cv(p + ".out_layers.3", layer.cout, layer.cout, 3)
                if layer.cin != layer.cout:
                    cv(p + ".skip_connection", layer.cin, layer.cout, 1)
        elif isinstance(layer, Attn):
            gn(p + ".norm", layer.c)
            if a.family == "ddpm":
                for n in ("q", "k", "v", "proj_out"):
                    cv(p + "." + n, layer.c, layer.c, 1)
            else:
                shapes[p + ".qkv.weight"], shapes[p + ".qkv.bias"] = (3 * layer.c, layer.c, 1), (3 * layer.c,)
                shapes[p + ".proj_out.weight"], shapes[p + ".proj_out.bias"] = (layer.c, layer.c, 1), (layer.c,)
        else:
            cv(p + ".conv", layer.c, layer.c, 3)

    for stage in a.enc:
        for layer in stage:
            add(layer)
    for layer in a.mid:
        add(layer)
    for stage in a.dec:
        for layer in stage:
            add(layer)
    gn(a.norm_out, first_ch)
    cv(a.conv_out, first_ch, a.out_ch, 3)
    shapes.update(delta_block_shapes(a, n_delta_blocks))
    return shapes


def delta_block_shapes(a: Arch, n):
    """layer_i DeltaBlocks (ddpm/diffusion.py:228-249, improved_ddpm/unet.py:776-834); same keys as the shipped
    checkpoint/*.pth files."""
    shapes = {}
    c = a.mid_ch
    for i in range(n):
        p = f"layer_{i}"
        if a.family == "ddpm":
            shapes[p + ".conv1.weight"], shapes[p + ".conv1.bias"] = (c, c, 1, 1), (c,)
            shapes[p + ".temb_proj.weight"], shapes[p + ".temb_proj.bias"] = (c, a.temb_ch), (c,)
            shapes[p + ".norm2.weight"], shapes[p + ".norm2.bias"] = (c,), (c,)
            shapes[p + ".conv2.weight"], shapes[p + ".conv2.bias"] = (c, c, 1, 1), (c,)
        else:
            shapes[p + ".in_layers.0.weight"], shapes[p + ".in_layers.0.bias"] = (c,), (c,)
            shapes[p + ".in_layers.2.weight"], shapes[p + ".in_layers.2.bias"] = (c, c, 1, 1), (c,)
            shapes[p + ".emb_layers.1.weight"], shapes[p + ".emb_layers.1.bias"] = (c, a.temb_ch), (c,)
            shapes[p + ".out_layers.0.weight"], shapes[p + ".out_layers.0.bias"] = (c,), (c,)
            shapes[p + ".out_layers.3.weight"], shapes[p + ".out_layers.3.bias"] = (c, c, 1, 1), (c,)
    return shapes
