"""Throughput of the other BASELINE.json configs (3, 4, 5) on ONE GPU at their per-GPU batch, device-resident inputs,
random-init weights — supplementary to bench.py (whose line is configs[1]).  Prints one JSON object per config.

    python tools/bench_configs.py [3 4 5a 5b]
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddnm_b200.model import Model, create_model                     # noqa: E402
from ddnm_b200 import operators as E                                 # noqa: E402
from ddnm_b200.sampler import sample_device                          # noqa: E402
from ddnm_b200.schedule import time_pairs                            # noqa: E402
from ddnm_b200.weights import random_state_dict, random_state_dict_openai  # noqa: E402

ns = types.SimpleNamespace
dev = torch.device("cuda", 0)


def celeba():
    mcfg = ns(model=ns(type="simple", ch=128, out_ch=3, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[16],
                       dropout=0.0, in_channels=3, resamp_with_conv=True), data=ns(image_size=256),
              diffusion=ns(num_diffusion_timesteps=1000))
    m = Model(mcfg)
    m.load_state_dict(random_state_dict(mcfg, 1234))
    return m


def imagenet():
    m = create_model(image_size=256, num_channels=256, num_res_blocks=2, learn_sigma=True, attention_resolutions="32,16,8",
                     num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True)
    m.load_state_dict(random_state_dict_openai(m, 1234))
    return m


def gauss_kernel():
    sigma = 10
    pdf = lambda z: torch.exp(torch.Tensor([-0.5 * (z / sigma) ** 2]))   # noqa: E731
    k = torch.Tensor([pdf(-2), pdf(-1), pdf(0), pdf(1), pdf(2)])
    return k / k.sum()


def run(name, model, op, B, T, tl, tr, sigma_y, plus, steps=1):
    conf = ns(diffusion=ns(num_diffusion_timesteps=1000), time_travel=ns(T_sampling=T, travel_length=tl, travel_repeat=tr))
    pairs = time_pairs(1000, T, tl, tr)
    g = torch.Generator().manual_seed(1234)
    x_orig = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    y = op.A(x_orig)
    if plus:
        y = y + sigma_y * torch.randn_like(y)
    x_T = torch.randn(B, 3, 256, 256, device=dev)
    noise = torch.empty(len(pairs), B, 3, 256, 256, device=dev)
    for k in range(len(pairs)):
        noise[k].normal_()
    betas = torch.from_numpy(np.linspace(1e-4, 2e-2, 1000, dtype="float64")).float().to(dev)
    fn = lambda: sample_device(x_T, model, betas, 0.85, op, y, sigma_y, plus, conf, noise=noise)   # noqa: E731
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        x0, _ = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    resid = (op.A(x0) - y).abs().max().item() if not plus else None
    evals = sum(1 for i, j in pairs if j < i)
    print(json.dumps(dict(config=name, batch=B, T_sampling=T, travel=[tl, tr], unet_evals=evals, pairs=len(pairs), ms_per_batch=ms,
                          images_per_sec=B * 1e3 / ms, max_abs_Ax0_minus_y=resid, finite=bool(torch.isfinite(x0).all()))), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5a", "5b"]
    if "3" in which or "4" in which:
        m = imagenet()
        if "3" in which:
            run("3: imagenet_256 colorization, DDNM, T=100, 8 images/GPU", m, E.Colorization(256, dev), 8, 100, 1, 1, 0.0, False)
        if "4" in which:
            mask = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "simplified.npz"))["mask_bits"])
            mask = torch.from_numpy(np.unpackbits(mask.numpy())[: 65536].astype(np.int64))
            mr = torch.nonzero(mask == 0).long().reshape(-1) * 3
            op = E.Inpainting(3, 256, torch.cat([mr, mr + 1, mr + 2]), dev)
            run("4: imagenet_256 inpainting (exp/inp_masks), DDNM+ sigma_y=0.05 (0.1 internal), travel 3/3, 8 images/GPU", m, op, 8, 100, 3, 3,
                0.1, True)
        del m
        torch.cuda.empty_cache()
    if "5a" in which or "5b" in which:
        m = celeba()
        if "5a" in which:
            run("5a: celeba_hq deblur_gauss, DDNM, T=250, 16 images/GPU", m, E.Deblurring(gauss_kernel().to(dev), 3, 256, dev), 16, 250, 1, 1, 0.0, False)
        if "5b" in which:
            perm = torch.randperm(256 ** 2, device=dev)
            run("5b: celeba_hq cs_walshhadamard ratio 0.25, DDNM, T=250, 16 images/GPU", m, E.WalshHadamardCS(3, 256, 4, perm, dev), 16, 250, 1, 1, 0.0, False)
