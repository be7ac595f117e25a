"""The full-size (256x256, real network shapes) parity cases shared by oracle/gen_golden.py (which runs the unmodified reference
on them and stores its results in tests/golden/fullsize.npz) and the tests (which regenerate the inputs from the same seeds).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch

FULLSIZE_CASES = [
    # key, net, operator, T_sampling, travel_length, travel_repeat, sigma_y (already doubled; 0 = ddnm_diffusion)
    ("cfg1", "celeba", "sr4", 20, 1, 1, 0.0),            # BASELINE configs[0]: celeba sr4 T=20 B=1 sigma_y=0
    ("celeba_wh", "celeba", "wh", 3, 1, 1, 0.0),         # configs[4] operators at full size, short schedule
    ("celeba_deblur", "celeba", "deblur", 3, 1, 1, 0.0),
    ("celeba_deblur_uni", "celeba", "deblur_uni", 2, 1, 1, 0.1),
    ("imagenet_color", "imagenet", "color", 3, 1, 1, 0.0),        # configs[2]
    ("imagenet_inpaint", "imagenet", "inpaint", 4, 2, 2, 0.1),    # configs[3]: real mask.npy, DDNM+ sigma_y 0.05 (x2), time travel
]


def fullsize_inputs(key, n_pairs):
    """(x_orig, x_T, noise tape) of a full-size case, all from seeds (the tests regenerate them; nothing is stored)."""
    seed = 9000 + sum(ord(c) for c in key)
    g = torch.Generator().manual_seed(seed)
    x_orig = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    x_T = torch.randn(1, 3, 256, 256, generator=g)
    tape = [torch.randn(1, 3, 256, 256, generator=g) for _ in range(n_pairs)]
    ynoise = torch.randn(1, 3 * 256 * 256, generator=g)
    return x_orig, x_T, tape, ynoise


def uni_kernel():
    return torch.Tensor([1 / 9] * 9)          # diffusion.py:500-503


def mask_from_bits(bits):
    """(256, 256) 0/1 array of exp/inp_masks/mask.npy from its packed bits in the fixture."""
    return np.unpackbits(bits)[: 256 * 256].reshape(256, 256)


def oracle_op(g, name):
    """Oracle operator of a full-size case built from the artefacts stored in fullsize.npz (``g``)."""
    from . import operators as O
    if name == "sr4":
        return O.SuperResolution.make(3, 256, 4)
    if name == "color":
        return O.Colorization.make(256)
    if name == "inpaint":
        return O.Inpainting(3, 256, mask_from_bits(g["mask_bits"]))
    if name == "wh":
        return O.WalshHadamardCS(3, 256, 4, torch.from_numpy(g["wh_perm"]).long())
    if name in ("deblur", "deblur_uni"):
        a = lambda k: torch.from_numpy(g[f"{name}_art_{k}"])     # noqa: E731
        return O.Deblurring(3, 256, a("U_small"), a("V_small"), a("singulars"), a("singulars_orig"), a("perm").long())
    raise KeyError(name)


def measurement(op, x_orig, ynoise, sigma_y):
    y = op.A(x_orig.reshape(x_orig.shape[0], -1))
    if sigma_y > 0:
        y = y + sigma_y * ynoise[:, : y.shape[1]]
    return y
