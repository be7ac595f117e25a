"""Synthetic (random-init) weights in the reference checkpoint layout, for benchmarks and smoke runs where the real
CelebA-HQ checkpoint cannot be downloaded.  Parameters are drawn exactly as ``torch.manual_seed(seed); Model(config)``
would draw them (same torch.nn layer types constructed in the order of guided_diffusion/models.py:216-299), so the
result equals the reference model's own ``state_dict()`` for that seed."""
import torch
import torch.nn as nn


def random_state_dict(config, seed=1234):
    m = config.model
    ch, out_ch, mult = int(m.ch), int(m.out_ch), tuple(int(v) for v in m.ch_mult)
    nrb, attn_res = int(m.num_res_blocks), tuple(int(v) for v in m.attn_resolutions)
    res = int(config.data.image_size)
    tch = 4 * ch
    sd = {}
    torch.manual_seed(seed)

    def take(prefix, module):
        for k, v in module.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    def res_block(prefix, cin, cout):
        take(prefix + ".norm1", nn.GroupNorm(32, cin, eps=1e-6))
        take(prefix + ".conv1", nn.Conv2d(cin, cout, 3, padding=1))
        take(prefix + ".temb_proj", nn.Linear(tch, cout))
        take(prefix + ".norm2", nn.GroupNorm(32, cout, eps=1e-6))
        take(prefix + ".conv2", nn.Conv2d(cout, cout, 3, padding=1))
        if cin != cout:
            take(prefix + ".nin_shortcut", nn.Conv2d(cin, cout, 1))

    def attn_block(prefix, c):
        take(prefix + ".norm", nn.GroupNorm(32, c, eps=1e-6))
        for leaf in ("q", "k", "v", "proj_out"):
            take(f"{prefix}.{leaf}", nn.Conv2d(c, c, 1))

    take("temb.dense.0", nn.Linear(ch, tch))
    take("temb.dense.1", nn.Linear(tch, tch))
    take("conv_in", nn.Conv2d(int(m.in_channels), ch, 3, padding=1))
    widths = [ch * v for v in mult]
    prev = [ch] + widths[:-1]
    cur = ch
    for lv, w in enumerate(widths):
        cur = prev[lv]
        for ib in range(nrb):
            res_block(f"down.{lv}.block.{ib}", cur, w)
            cur = w
            if res in attn_res:
                attn_block(f"down.{lv}.attn.{ib}", cur)
        if lv + 1 < len(widths):
            take(f"down.{lv}.downsample.conv", nn.Conv2d(cur, cur, 3, stride=2))
            res //= 2
    res_block("mid.block_1", cur, cur)
    attn_block("mid.attn_1", cur)
    res_block("mid.block_2", cur, cur)
    for lv in range(len(widths) - 1, -1, -1):
        w = widths[lv]
        for ib in range(nrb + 1):
            skip = prev[lv] if ib == nrb else w
            res_block(f"up.{lv}.block.{ib}", cur + skip, w)
            cur = w
            if res in attn_res:
                attn_block(f"up.{lv}.attn.{ib}", cur)
        if lv > 0:
            take(f"up.{lv}.upsample.conv", nn.Conv2d(cur, cur, 3, padding=1))
            res *= 2
    take("norm_out", nn.GroupNorm(32, cur, eps=1e-6))
    take("conv_out", nn.Conv2d(cur, out_ch, 3, padding=1))
    return sd
