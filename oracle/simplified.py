"""Oracle: the "simplified" DDNM+ loop of the reference runner (README quick start).

Restates /root/reference/guided_diffusion/diffusion.py: helpers MeanUpsample :27-31, color2gray :33-36, gray2color :38-42;
operator table :244-290; loop :325-395 (NB sigma_t = sqrt(1 - at_next**2), scalar lambda_t / gamma_t, Eq. 17/19).
Image-space operators on (B, 3, H, W); the reference enforces B == 1 (:308-309) — for B > 1 color2gray's ``repeat`` is
generalised per image.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch

from .schedule import alpha_bar_table, time_pairs


def mean_upsample(x, scale):
    n, c, h, w = x.shape
    out = torch.zeros(n, c, h, scale, w, scale) + x.view(n, c, h, 1, w, 1)
    return out.view(n, c, scale * h, scale * w)


def color2gray(x):
    coef = 1 / 3
    g = x[:, 0, :, :] * coef + x[:, 1, :, :] * coef + x[:, 2, :, :] * coef
    return g[:, None].repeat(1, 3, 1, 1)


def gray2color(x):
    x = x[:, 0, :, :]
    coef = 1 / 3
    base = coef ** 2 + coef ** 2 + coef ** 2
    return torch.stack((x * coef / base, x * coef / base, x * coef / base), 1)


def degradation(deg, scale=1, mask=None, size=256):
    """(A, Ap) of diffusion.py:244-290.  mask: (H, W) tensor of 0/1."""
    pool = torch.nn.AdaptiveAvgPool2d((size // scale, size // scale)) if scale > 1 else None
    if deg == "colorization":
        return color2gray, gray2color
    if deg == "denoising":
        return (lambda z: z), (lambda z: z)
    if deg == "sr_averagepooling":
        return pool, (lambda z: mean_upsample(z, scale))
    if deg == "inpainting":
        return (lambda z: z * mask), (lambda z: z * mask)
    if deg in ("mask_color_sr", "diy"):
        return (lambda z: pool(color2gray(z * mask))), (lambda z: gray2color(mean_upsample(z, scale)) * mask)
    raise NotImplementedError("degradation type not supported")


def simplified_sample(x_T, model, betas, eta, A, Ap, y, sigma_y, noise, num_timesteps=1000, t_sampling=100, travel_length=1,
                      travel_repeat=1):
    abar = alpha_bar_table(betas)
    pairs = time_pairs(num_timesteps, t_sampling, travel_length, travel_repeat)
    n = x_T.shape[0]
    xt, x0_last = x_T, None
    for k, (i, j) in enumerate(pairs):
        at_next = abar[j + 1]
        z = noise[k]
        if j < i:
            at = abar[i + 1]
            sigma_t = (1 - at_next ** 2).sqrt()
            et = model(xt, torch.ones(n) * i)
            if et.size(1) == 6:
                et = et[:, :3]
            x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
            if sigma_t >= at_next * sigma_y:
                lambda_t = 1.
                gamma_t = (sigma_t ** 2 - (at_next * sigma_y) ** 2).sqrt()
            else:
                lambda_t = (sigma_t) / (at_next * sigma_y)
                gamma_t = 0.
            x0_hat = x0_t - lambda_t * Ap(A(x0_t) - y)
            c1 = (1 - at_next).sqrt() * eta
            c2 = (1 - at_next).sqrt() * ((1 - eta ** 2) ** 0.5)
            xt_next = at_next.sqrt() * x0_hat + gamma_t * (c1 * z + c2 * et)
            x0_last = x0_t
        else:
            xt_next = at_next.sqrt() * x0_last + z * (1 - at_next).sqrt()
        xt = xt_next
    return xt, x0_last
