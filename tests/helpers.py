"""Shared builders for the tests: oracle/engine operator pairs from golden artefacts, config namespaces."""
import types

import numpy as np
import torch

from oracle import operators as O


def ns(**k):
    return types.SimpleNamespace(**k)


def model_config(cfg):
    return ns(model=ns(type="simple", ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                       attn_resolutions=list(cfg.attn_resolutions), dropout=0.0, in_channels=cfg.in_channels, resamp_with_conv=True),
              data=ns(image_size=cfg.resolution), diffusion=ns(num_diffusion_timesteps=1000))


def sampler_config(T, tl, tr):
    return ns(diffusion=ns(num_diffusion_timesteps=1000), time_travel=ns(T_sampling=T, travel_length=tl, travel_repeat=tr))


def art(gold_ops, tag, name, key):
    return torch.from_numpy(gold_ops[f"{tag}_{name}_art_{key}"])


def gauss_kernel():
    sigma = 10
    pdf = lambda z: torch.exp(torch.Tensor([-0.5 * (z / sigma) ** 2]))   # noqa: E731
    k = torch.Tensor([pdf(-2), pdf(-1), pdf(0), pdf(1), pdf(2)])
    return k / k.sum()


def oracle_ops(gold_ops, dim=32):
    """Oracle operators at image size ``dim`` built from the artefacts stored in the golden file (dim 32) or by
    repeating the constructor arithmetic (dim 256; perm for WH comes from the file)."""
    tag = f"d{dim}"
    ops = {}
    if dim == 32:
        ops["sr4"] = O.SuperResolution(3, dim, 4, art(gold_ops, tag, "sr4", "U_small"), art(gold_ops, tag, "sr4", "singulars_small"),
                                       art(gold_ops, tag, "sr4", "V_small"))
        ops["color"] = O.Colorization(dim, art(gold_ops, tag, "color", "U_small"), art(gold_ops, tag, "color", "singulars_small"),
                                      art(gold_ops, tag, "color", "V_small"))
        ops["inpaint"] = O.Inpainting(3, dim, art(gold_ops, tag, "inpaint", "mask").numpy())
        ops["deblur"] = O.Deblurring(3, dim, art(gold_ops, tag, "deblur", "U_small"), art(gold_ops, tag, "deblur", "V_small"),
                                     art(gold_ops, tag, "deblur", "singulars"), art(gold_ops, tag, "deblur", "singulars_orig"),
                                     art(gold_ops, tag, "deblur", "perm"))
        ops["bicubic"] = O.SRConv(3, dim, 4, art(gold_ops, tag, "bicubic", "U_small"), art(gold_ops, tag, "bicubic", "singulars_small"),
                                  art(gold_ops, tag, "bicubic", "V_small"))
        ops["deblur2d"] = O.Deblurring2D(3, dim, art(gold_ops, tag, "deblur2d", "U_small1"), art(gold_ops, tag, "deblur2d", "V_small1"),
                                         art(gold_ops, tag, "deblur2d", "U_small2"), art(gold_ops, tag, "deblur2d", "V_small2"),
                                         art(gold_ops, tag, "deblur2d", "singulars"), art(gold_ops, tag, "deblur2d", "perm"))
    else:
        ops["sr4"] = O.SuperResolution.make(3, dim, 4)
        ops["color"] = O.Colorization.make(dim)
        ops["deblur"] = O.Deblurring.make(gauss_kernel(), 3, dim)
        ops["bicubic"] = O.SRConv.make(O.SRConv.bicubic_kernel(4), 3, dim, 4)
    ops["wh"] = O.WalshHadamardCS(3, dim, 4, art(gold_ops, tag, "wh", "perm"))
    ops["denoise"] = O.Denoising(3, dim)
    ops["cs"] = O.CS(3, dim, 0.25, O.hadamard_basis())
    return ops


def engine_op(name, oop, dim, device="cuda"):
    """The ddnm_b200 operator sharing the oracle operator's artefacts."""
    from ddnm_b200 import operators as E
    if name == "sr4":
        return E.SuperResolution(3, dim, 4, device, artefacts=(oop.U_small, oop.singulars_small, oop.V_small))
    if name == "color":
        return E.Colorization(dim, device, artefacts=(oop.U_small, oop.singulars_small, oop.V_small))
    if name == "inpaint":
        m = (~oop.mask_img).reshape(-1)
        mr = torch.nonzero(m).long().reshape(-1) * 3           # diffusion.py:467-470
        return E.Inpainting(3, dim, torch.cat([mr, mr + 1, mr + 2]), device)
    if name == "wh":
        return E.WalshHadamardCS(3, dim, 4, oop.perm, device)
    if name == "deblur":
        return E.Deblurring(None, 3, dim, device, artefacts=(oop.U_small, oop.V_small, oop.S, oop.S_orig, oop.perm))
    if name == "bicubic":
        return E.SRConv(None, 3, dim, device, stride=4, artefacts=(oop.U_small, oop.S_small, oop.V_small))
    if name == "denoise":
        return E.Denoising(3, dim, device)
    if name == "cs":
        return E.CS(3, dim, 0.25, device, artefacts=oop.V_small)
    if name == "deblur2d":
        return E.Deblurring2D(None, None, 3, dim, device, artefacts=(oop.U1, oop.V1, oop.U2, oop.V2, oop.S, oop.perm))
    raise KeyError(name)


LAMBDA_CASES = [(0.9, 0.1, 0.3), (0.99, 0.1, 0.02), (1.0, 0.1, 0.0), (0.5, 0.0, 0.4)]


def assert_close(got, ref, rtol=1e-3, atol=1e-4, what=""):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    if got.shape != ref.shape and got.numel() == ref.numel():
        ref = ref.reshape(got.shape)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol)
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol} atol={atol}; max err {err.max().item():.3e} (ref absmax {ref.abs().max().item():.3e})"


def openai_model_kwargs(cfg):
    """keyword arguments of script_util.create_model for an oracle OpenAIUNetConfig (imagenet_256.yml style)."""
    return dict(image_size=cfg.image_size, num_channels=cfg.model_channels, num_res_blocks=cfg.num_res_blocks,
                channel_mult=",".join(str(c) for c in cfg.channel_mult), learn_sigma=(cfg.out_channels == 6), class_cond=False,
                attention_resolutions=",".join(str(r) for r in cfg.attention_resolutions), num_heads=4,
                num_head_channels=cfg.num_head_channels, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
                resblock_updown=True, use_fp16=True, use_new_attention_order=False)
