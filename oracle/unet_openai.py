"""Oracle: the guided-diffusion ("openai") UNet of imagenet_256.yml as a flat functional PyTorch-fp32 program driven
by a reference-layout ``state_dict``.

Restates /root/reference/guided_diffusion/unet.py for the variant the shipped configs use (use_scale_shift_norm,
resblock_updown, legacy attention order, num_head_channels=64; class_cond adds the label embedding of :478-479, 651-653):
  TimestepEmbedSequential :66-78, Upsample/Downsample (conv-free inside ResBlocks) :81-140, ResBlock._forward :236-256,
  AttentionBlock._forward :299-305, QKVAttentionLegacy :337-354, UNetModel.__init__ :460-617 and .forward :635-664;
  guided_diffusion/nn.py GroupNorm32 :17-19 (32 groups, eps 1e-5), timestep_embedding :103-121;
  guided_diffusion/script_util.py create_model :130-185.
fp32 throughout (the parity mode of SURVEY.md §7; the reference's fp16 torso is a lower-precision variant of the same maths).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class OpenAIUNetConfig:
    image_size: int = 256
    model_channels: int = 256
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 1, 2, 2, 4, 4)
    attention_resolutions: Tuple[int, ...] = (32, 16, 8)     # spatial sizes, as written in the yml
    num_head_channels: int = 64
    out_channels: int = 6                                      # learn_sigma
    in_channels: int = 3
    num_classes: Optional[int] = None                          # class_cond (imagenet_256_cc.yml): nn.Embedding(num_classes, 4*ch)

    @staticmethod
    def imagenet_256():
        return OpenAIUNetConfig()

    @staticmethod
    def tiny():
        # 32x32, 64..128 channels, attention at 16 and 8, one updown level pair: every layer kind, seconds on CPU
        return OpenAIUNetConfig(image_size=32, model_channels=64, num_res_blocks=1, channel_mult=(1, 2, 2),
                                attention_resolutions=(16, 8), num_head_channels=64, out_channels=6)

    @staticmethod
    def tiny_class_cond():
        # the tiny net with the 1000-class label embedding of imagenet_256_cc.yml (unet.py:478-479)
        c = OpenAIUNetConfig.tiny()
        c.num_classes = 1000
        return c

    @property
    def attention_ds(self):
        return tuple(self.image_size // r for r in self.attention_resolutions)


def timestep_embedding(t, dim):
    # nn.py:103-121 — [cos | sin], frequency exponent log(1e4)/half
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(sd, name, x):
    return F.group_norm(x.float(), 32, sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _resblock(sd, p, x, emb, up=False, down=False):
    # unet.py:236-256 with use_scale_shift_norm=True
    h = F.silu(_gn(sd, p + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[..., None, None]
    scale, shift = torch.chunk(emb_out, 2, dim=1)
    h = _gn(sd, p + ".out_layers.0", h) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(sd, p, x, head_ch):
    # unet.py:299-305 + QKVAttentionLegacy :337-354
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + ".norm", xf), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    n_heads = c // head_ch
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", weight, v).reshape(bs, -1, length)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def block_plan(cfg: OpenAIUNetConfig):
    """The module list of UNetModel.__init__ (unet.py:479-611) as data: for each input / output block a list of
    ('conv'|'res'|'res_down'|'res_up'|'attn', cin, cout)."""
    mc = cfg.model_channels
    ch = int(cfg.channel_mult[0] * mc)
    inp = [[("conv", cfg.in_channels, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, int(mult * mc))]
            ch = int(mult * mc)
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inp.append([("res_down", ch, ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("attn", ch, ch), ("res", ch, ch)]
    out = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * mult))]
            ch = int(mc * mult)
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, ch))
            if level and i == cfg.num_res_blocks:
                layers.append(("res_up", ch, ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out, ch


def forward(sd, x, t, cfg: OpenAIUNetConfig, taps=None, y=None):
    inp, mid, out, _ = block_plan(cfg)
    emb = timestep_embedding(t, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    assert (y is not None) == (cfg.num_classes is not None), "must specify y if and only if the model is class-conditional"   # unet.py:644-646
    if cfg.num_classes is not None:
        assert y.shape == (x.shape[0],)
        emb = emb + sd["label_emb.weight"][y.long()]             # unet.py:651-653

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()
        return v

    def run(prefix, layers, h):
        for j, (kind, _cin, _cout) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv":
                h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif kind == "attn":
                h = _attention(sd, p, h, cfg.num_head_channels)
            else:
                h = _resblock(sd, p, h, emb, up=(kind == "res_up"), down=(kind == "res_down"))
        return h

    hs = []
    h = x.float()
    for i, layers in enumerate(inp):
        h = tap(f"in.{i}", run(f"input_blocks.{i}", layers, h))
        hs.append(h)
    h = tap("mid", run("middle_block", mid, h))
    for i, layers in enumerate(out):
        h = tap(f"out.{i}", run(f"output_blocks.{i}", layers, torch.cat([h, hs.pop()], dim=1)))
    h = F.silu(_gn(sd, "out.0", h))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def init_state_dict(cfg: OpenAIUNetConfig, seed=1234, zero_std=0.02):
    """Weights identical to ``torch.manual_seed(seed); create_model(...)`` followed by re-drawing every zero-initialised
    tensor (ResBlock out conv, attention proj_out, final conv: unet.py:210-212,294,616) from N(0, zero_std) with a
    ``torch.Generator().manual_seed(seed + 1)`` in state_dict order — random-init output would otherwise be exactly 0
    and parity vacuous.  Checked against the reference in oracle/gen_golden.py."""
    torch.manual_seed(seed)
    sd = {}
    zero_keys = []
    tdim = cfg.model_channels * 4

    def put(name, mod, zero=False):
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
            if zero:
                zero_keys.append(f"{name}.{k}")

    def res(p, cin, cout):
        put(p + ".in_layers.0", nn.GroupNorm(32, cin))
        put(p + ".in_layers.2", nn.Conv2d(cin, cout, 3, padding=1))
        put(p + ".emb_layers.1", nn.Linear(tdim, 2 * cout))
        put(p + ".out_layers.0", nn.GroupNorm(32, cout))
        put(p + ".out_layers.3", nn.Conv2d(cout, cout, 3, padding=1), zero=True)
        if cin != cout:
            put(p + ".skip_connection", nn.Conv2d(cin, cout, 1))

    def attn(p, c):
        put(p + ".norm", nn.GroupNorm(32, c))
        put(p + ".qkv", nn.Conv1d(c, 3 * c, 1))
        put(p + ".proj_out", nn.Conv1d(c, c, 1), zero=True)

    def build(prefix, layers):
        for j, (kind, cin, cout) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv":
                put(p, nn.Conv2d(cin, cout, 3, padding=1))
            elif kind == "attn":
                attn(p, cin)
            else:
                res(p, cin, cout)

    inp, mid, out, ch = block_plan(cfg)
    put("time_embed.0", nn.Linear(cfg.model_channels, tdim))
    put("time_embed.2", nn.Linear(tdim, tdim))
    if cfg.num_classes is not None:
        put("label_emb", nn.Embedding(cfg.num_classes, tdim))     # created right after time_embed (unet.py:478-479)
    for i, layers in enumerate(inp):
        build(f"input_blocks.{i}", layers)
    build("middle_block", mid)
    for i, layers in enumerate(out):
        build(f"output_blocks.{i}", layers)
    put("out.0", nn.GroupNorm(32, ch))
    put("out.2", nn.Conv2d(int(cfg.channel_mult[0] * cfg.model_channels), cfg.out_channels, 3, padding=1), zero=True)
    g = torch.Generator().manual_seed(seed + 1)
    for k in zero_keys:
        sd[k] = torch.randn(sd[k].shape, generator=g) * zero_std
    return sd
