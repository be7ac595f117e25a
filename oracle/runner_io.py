"""TEST INFRASTRUCTURE ONLY — CPU restatement of the runner's I/O step either side of the sampling loop.

Follows datasets/__init__.py:195-227 (logit_transform, data_transform, inverse_data_transform),
guided_diffusion/diffusion.py:596-602 (save_image + PSNR) and torchvision.utils.save_image's uint8 quantisation
(``mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(uint8)``).  Pinned to the reference by tests/golden/runner_io.npz
(oracle/gen_golden.py runner_fixtures).  Never imported by the product path.
"""
import torch


def data_transform(X, rescaled, logit, uniform_noise=None, gauss_noise=None):
    # datasets/__init__.py:201-213; the two dequantisation draws are passed in (torch.rand_like / torch.randn_like there)
    if uniform_noise is not None:
        X = X / 256.0 * 255.0 + uniform_noise / 256.0
    if gauss_noise is not None:
        X = X + gauss_noise * 0.01
    if rescaled:
        X = 2 * X - 1.0
    elif logit:
        lam = 1e-6
        image = lam + (1 - 2 * lam) * X
        X = torch.log(image) - torch.log1p(-image)
    return X


def inverse_data_transform(X, rescaled, logit):
    # datasets/__init__.py:216-227
    if logit:
        X = torch.sigmoid(X)
    elif rescaled:
        X = (X + 1.0) / 2.0
    return torch.clamp(X, 0.0, 1.0)


def to_uint8_hwc(img01):
    """(C,H,W) or (B,C,H,W) in [0,1] -> uint8 (..., H, W, C): what tvu.save_image hands to PIL for a single image."""
    q = img01.mul(255).add(0.5).clamp(0, 255)
    return q.movedim(-3, -1).to(torch.uint8)


def psnr(x01, orig01):
    # diffusion.py:600-601, per image
    mse = torch.mean((x01 - orig01) ** 2)
    return 10 * torch.log10(1 / mse)
