"""Oracle: the "simple" DDPM UNet (celeba_hq.yml, model.type == "simple") as a flat
functional PyTorch-fp32 program driven by a reference-layout ``state_dict``.

Restates /root/reference/guided_diffusion/models.py:
  get_timestep_embedding :6-24, nonlinearity :27-29, Normalize (GN32, eps 1e-6) :32-33,
  Upsample :36-52, Downsample (pad (0,1,0,1) + 3x3 s2) :55-74, ResnetBlock :77-134,
  AttnBlock :137-189, Model.__init__/forward :192-341.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math
from dataclasses import dataclass, field
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class SimpleUNetConfig:
    ch: int = 128
    out_ch: int = 3
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4, 4)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    in_channels: int = 3
    resolution: int = 256
    num_groups: int = 32
    gn_eps: float = 1e-6

    @staticmethod
    def celeba_hq():
        return SimpleUNetConfig()

    @staticmethod
    def tiny():
        # smallest config that still exercises every layer kind (attention at 16x16, one down/up level,
        # 1x1 shortcut on channel change, concat skip) with channel counts the tensor-core path accepts
        return SimpleUNetConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(16,), resolution=32)


def swish(x):
    return x * torch.sigmoid(x)


def sinusoid_embedding(t, dim):
    # models.py:6-24 — [sin | cos], frequency exponent log(1e4)/(half-1)
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000) / (half - 1)))
    arg = t.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def _gn(sd, name, x, cfg):
    return F.group_norm(x, cfg.num_groups, sd[name + ".weight"], sd[name + ".bias"], cfg.gn_eps)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _resblock(sd, p, x, temb, cfg):
    # models.py:115-134
    h = _conv(sd, p + ".conv1", swish(_gn(sd, p + ".norm1", x, cfg)))
    h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    h = _conv(sd, p + ".conv2", swish(_gn(sd, p + ".norm2", h, cfg)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def _attn(sd, p, x, cfg):
    # models.py:164-189 — single head, softmax over keys, scale C^-0.5
    hn = _gn(sd, p + ".norm", x, cfg)
    q = _conv(sd, p + ".q", hn, padding=0)
    k = _conv(sd, p + ".k", hn, padding=0)
    v = _conv(sd, p + ".v", hn, padding=0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    o = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", o, padding=0)


def forward(sd, x, t, cfg: SimpleUNetConfig, taps=None):
    """eps = Model(x, t).  ``taps``: optional dict filled with intermediate tensors by name."""
    assert x.shape[2] == x.shape[3] == cfg.resolution
    nlev = len(cfg.ch_mult)
    temb = sinusoid_embedding(t, cfg.ch)
    temb = F.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = F.linear(swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()
        return v

    res = cfg.resolution
    hs = [tap("conv_in", _conv(sd, "conv_in", x))]
    for lv in range(nlev):
        for ib in range(cfg.num_res_blocks):
            h = _resblock(sd, f"down.{lv}.block.{ib}", hs[-1], temb, cfg)
            if res in cfg.attn_resolutions:
                h = _attn(sd, f"down.{lv}.attn.{ib}", h, cfg)
            hs.append(tap(f"down.{lv}.{ib}", h))
        if lv != nlev - 1:
            hs.append(tap(f"down.{lv}.ds", _conv(sd, f"down.{lv}.downsample.conv", F.pad(hs[-1], (0, 1, 0, 1)), stride=2, padding=0)))
            res //= 2
    h = hs[-1]
    h = tap("mid.block_1", _resblock(sd, "mid.block_1", h, temb, cfg))
    h = tap("mid.attn_1", _attn(sd, "mid.attn_1", h, cfg))
    h = tap("mid.block_2", _resblock(sd, "mid.block_2", h, temb, cfg))
    for lv in reversed(range(nlev)):
        for ib in range(cfg.num_res_blocks + 1):
            h = _resblock(sd, f"up.{lv}.block.{ib}", torch.cat([h, hs.pop()], dim=1), temb, cfg)
            if res in cfg.attn_resolutions:
                h = _attn(sd, f"up.{lv}.attn.{ib}", h, cfg)
            tap(f"up.{lv}.{ib}", h)
        if lv != 0:
            h = tap(f"up.{lv}.us", _conv(sd, f"up.{lv}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest")))
            res *= 2
    h = swish(_gn(sd, "norm_out", h, cfg))
    return _conv(sd, "conv_out", h)


def init_state_dict(cfg: SimpleUNetConfig, seed=1234):
    """Random-init weights identical to ``torch.manual_seed(seed); Model(config).state_dict()``.

    The reference builds its torch.nn layers in a fixed order (models.py:216-299) and torch's default
    initialisers draw from the global RNG, so constructing the same layer types in the same order
    reproduces the same tensors bit for bit (checked against the reference in oracle/gen_golden.py).
    """
    torch.manual_seed(seed)
    sd = {}

    def put(name, mod):
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()

    def conv(name, cin, cout, k):
        put(name, nn.Conv2d(cin, cout, k, 1, k // 2))

    def gn(name, c):
        put(name, nn.GroupNorm(cfg.num_groups, c, eps=cfg.gn_eps))

    def resblock(p, cin, cout):
        gn(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        put(p + ".temb_proj", nn.Linear(cfg.ch * 4, cout))
        gn(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    def attn(p, c):
        gn(p + ".norm", c)
        for nm in ("q", "k", "v", "proj_out"):
            conv(f"{p}.{nm}", c, c, 1)

    ch, nlev = cfg.ch, len(cfg.ch_mult)
    put("temb.dense.0", nn.Linear(ch, ch * 4))
    put("temb.dense.1", nn.Linear(ch * 4, ch * 4))
    conv("conv_in", cfg.in_channels, ch, 3)
    res = cfg.resolution
    in_mult = (1,) + tuple(cfg.ch_mult)
    block_in = None
    for lv in range(nlev):
        block_in, block_out = ch * in_mult[lv], ch * cfg.ch_mult[lv]
        # the reference appends ResnetBlock then (maybe) AttnBlock per i_block (models.py:241-249)
        for ib in range(cfg.num_res_blocks):
            resblock(f"down.{lv}.block.{ib}", block_in, block_out)
            block_in = block_out
            if res in cfg.attn_resolutions:
                attn(f"down.{lv}.attn.{ib}", block_in)
        if lv != nlev - 1:
            conv(f"down.{lv}.downsample.conv", block_in, block_in, 3)
            res //= 2
    resblock("mid.block_1", block_in, block_in)
    attn("mid.attn_1", block_in)
    resblock("mid.block_2", block_in, block_in)
    for lv in reversed(range(nlev)):
        block_out = ch * cfg.ch_mult[lv]
        skip_in = ch * cfg.ch_mult[lv]
        for ib in range(cfg.num_res_blocks + 1):
            if ib == cfg.num_res_blocks:
                skip_in = ch * in_mult[lv]
            resblock(f"up.{lv}.block.{ib}", block_in + skip_in, block_out)
            block_in = block_out
            if res in cfg.attn_resolutions:
                attn(f"up.{lv}.attn.{ib}", block_in)
        if lv != 0:
            conv(f"up.{lv}.upsample.conv", block_in, block_in, 3)
            res *= 2
    gn("norm_out", block_in)
    conv("conv_out", block_in, cfg.out_ch, 3)
    return sd
