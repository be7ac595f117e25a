"""Oracle: hq_demo's arbitrary-size DDNM ("mask-shift" restoration).

Restates /root/reference/hq_demo/guided_diffusion/gaussian_diffusion.py:
  * the spaced DDPM constants (:155-206 with respace.py:89-107),
  * ``p_mean_variance``'s "DDNM core" (:318-390): x0_t from eps, clipping, Eq. 19 lambda_t / gamma_t from the posterior
    coefficients, Eq. 17 ``x0_t_hat = lambda_t*Apy + x0_t - lambda_t*Ap(A(x0_t))``, the mask-shift overwrite of the window's
    already-restored region from ``x_temp`` (:344-384), the posterior mean of x0_t_hat and ``variance = gamma_t``,
  * ``p_sample`` (:431-493, incl. classifier ``condition_mean`` :414-430), ``_undo`` (:208-217),
  * the window loop of ``p_sample_loop_progressive`` (:578-750): 256x256 windows every 128 pixels over the (H_target, W_target)
    canvas, each window a full DDNM schedule (scheduler.py get_schedule_jump), the iterate carried over from window to window.
Gaussian draws come from an explicit tape in the reference's order (initial x, then one per p_sample / undo call).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math

import numpy as np
import torch


def get_schedule_jump(t_T, n_sample, jump_length, jump_n_sample):
    """scheduler.py:70-148 with the (default) inactive jump2 / jump3 levels dropped."""
    jumps = {j: jump_n_sample - 1 for j in range(0, t_T - jump_length, jump_length)}
    t = t_T
    ts = []
    while t >= 1:
        t = t - 1
        ts.append(t)
        if t + 1 < t_T - 1:
            for _ in range(n_sample - 1):
                t = t + 1
                ts.append(t)
                if t >= 0:
                    t = t - 1
                    ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] = jumps[t] - 1
            for _ in range(jump_length):
                t = t + 1
                ts.append(t)
    ts.append(-1)
    return ts


def space_timesteps(num_timesteps, count):
    """respace.py:24-86 for a single integer section."""
    frac_stride = 1 if count <= 1 else (num_timesteps - 1) / (count - 1)
    cur, taken = 0.0, []
    for _ in range(count):
        taken.append(round(cur))
        cur += frac_stride
    return sorted(set(taken))


class SpacedConstants:
    """float64 tables of the respaced process (gaussian_diffusion.py:165-206), indexed by the RESPACED step."""

    def __init__(self, diffusion_steps=1000, respacing=100):
        scale = 1000 / diffusion_steps
        betas = np.linspace(scale * 0.0001, scale * 0.02, diffusion_steps, dtype=np.float64)     # :80-93 (linear, use_scale)
        ac = np.cumprod(1.0 - betas, axis=0)
        use = set(space_timesteps(diffusion_steps, respacing))
        new_betas, self.timestep_map, last = [], [], 1.0
        for i, a in enumerate(ac):                                                                 # respace.py:97-105
            if i in use:
                new_betas.append(1 - a / last)
                last = a
                self.timestep_map.append(i)
        b = np.array(new_betas, dtype=np.float64)
        self.betas = b
        alphas = 1.0 - b
        acp = np.cumprod(alphas, axis=0)
        prev = np.append(1.0, acp[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / acp)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / acp - 1)
        self.posterior_variance = b * (1.0 - prev) / (1.0 - acp)
        self.posterior_mean_coef1 = b * np.sqrt(prev) / (1.0 - acp)
        self.posterior_mean_coef2 = (1.0 - prev) * np.sqrt(alphas) / (1.0 - acp)
        self.num_timesteps = len(b)


def f32(table, t):
    """_extract_into_tensor (:752-766): the float64 table entry as a float32 scalar tensor."""
    return torch.from_numpy(table)[t].float()


def color2gray(x):      # :54-57
    coef = 1 / 3
    x = x[:, 0, :, :] * coef + x[:, 1, :, :] * coef + x[:, 2, :, :] * coef
    return x.repeat(1, 3, 1, 1)


def gray2color(x):      # :59-63
    x = x[:, 0, :, :]
    coef = 1 / 3
    base = coef ** 2 + coef ** 2 + coef ** 2
    return torch.stack((x * coef / base, x * coef / base, x * coef / base), 1)


def mean_upsample(x, scale):   # :65-69
    n, c, h, w = x.shape
    out = torch.zeros(n, c, h, scale, w, scale) + x.view(n, c, h, 1, w, 1)
    return out.view(n, c, scale * h, scale * w)


def degradation(deg, scale, gt_shape):
    """(A, Ap, A_temp) of :598-649 for the degradations defined at arbitrary output size."""
    pool = torch.nn.AdaptiveAvgPool2d((256 // scale, 256 // scale))
    pool_t = torch.nn.AdaptiveAvgPool2d((gt_shape[2] // scale, gt_shape[3] // scale))
    up = lambda z: mean_upsample(z, scale)      # noqa: E731
    if deg == "sr_averagepooling":
        return pool, up, pool_t
    if deg == "colorization":
        return color2gray, gray2color, color2gray
    if deg == "sr_color":
        return (lambda z: color2gray(pool(z))), (lambda z: up(gray2color(z))), (lambda z: color2gray(pool_t(z)))
    raise NotImplementedError("degradation type not supported")


def shift_overwrite(x0h, x_temp, sh, sw, sh_total, sw_total, H, W):
    """The mask-shift trick (:344-384): the part of the window that earlier windows already restored is copied from the canvas."""
    if sw == 0 and sh == 0:
        return
    if sw == 0 and sh != 0:
        h_l = int(128 * sh)
        h_r = h_l + 128
        if sh == sh_total - 1 and H % 128 != 0:
            h_l = h_l - 128 + H % 128
            x0h[:, :, 0:256 - H % 128, :] = x_temp[:, :, h_l:h_r, 0:256]
        else:
            x0h[:, :, 0:128, :] = x_temp[:, :, h_l:h_r, 0:256]
        return
    w_l = int(128 * sw)
    w_r = w_l + 128
    h_l = int(128 * sh)
    h_r = h_l + 256
    if sw == sw_total - 1 and W % 128 != 0:
        w_l = w_l - 128 + W % 128
        if sh == sh_total - 1 and H % 128 != 0:
            h_l_tmp = h_l - 128 + H % 128
            x0h[:, :, :, 0:256 - W % 128] = x_temp[:, :, h_l_tmp:h_r, w_l:w_r]
        else:
            x0h[:, :, :, 0:256 - W % 128] = x_temp[:, :, h_l:h_r, w_l:w_r]
    else:
        if sh == sh_total - 1 and H % 128 != 0:
            h_l_tmp = h_l - 128 + H % 128
            x0h[:, :, :, 0:128] = x_temp[:, :, h_l_tmp:h_r, w_l:w_r]
        else:
            x0h[:, :, :, 0:128] = x_temp[:, :, h_l:h_r, w_l:w_r]
    if sh != 0:
        h_r = h_l + 128
        w_r = w_l + 256
        if sh == sh_total - 1 and H % 128 != 0:
            h_l = h_l - 128 + H % 128
            x0h[:, :, 0:256 - H % 128, :] = x_temp[:, :, h_l:h_r, w_l:w_r]
        else:
            x0h[:, :, 0:128, :] = x_temp[:, :, h_l:h_r, w_l:w_r]


def window_origin(sh, sw, sh_total, sw_total, H, W):
    """(h_l, w_l) of a window (:683-699) — also where its result is written back (:737-747)."""
    h_l = int(128 * sh)
    if sh == sh_total - 1 and H % 128 != 0:
        h_l = H - 256
    w_l = int(128 * sw)
    if sw == sw_total - 1 and W % 128 != 0:
        w_l = W - 256
    return h_l, w_l


def restore(model, gt, classes, noise, *, deg="sr_averagepooling", scale=4, sigma_y=0.0, resize_y=False, diffusion_steps=1000,
            respacing=100, jump=None, clip_denoised=True, cond_fn=None, trace=None):
    """``model(x, t_original, y) -> (B, 6, 256, 256)``; gt: the degraded input image (B,3,h,w) in [-1,1]; noise: list of
    (B,3,256,256) draws.  Returns the restored canvas (B,3,H_target,W_target)."""
    K = SpacedConstants(diffusion_steps, respacing)
    jump = jump or dict(t_T=respacing, n_sample=1, jump_length=10, jump_n_sample=3)
    noise = list(noise)
    B = gt.shape[0]
    if 256 % scale != 0:
        raise ValueError("Please set a SR scale divisible by 256")
    if resize_y:
        gt = mean_upsample(gt, scale)
    A, Ap, A_temp = degradation(deg, scale, gt.shape)
    Apy_temp = Ap(A_temp(gt))
    H, W = Apy_temp.shape[2], Apy_temp.shape[3]
    if H < 256 or W < 256:
        raise ValueError("Please set a larger SR scale")
    final = torch.zeros_like(Apy_temp)
    sh_total, sw_total = math.ceil(H / 128) - 1, math.ceil(W / 128) - 1
    x = noise.pop(0)                                      # th.randn(*shape) (:574), carried over from window to window
    x0_hat = None
    tmap = torch.tensor(K.timestep_map)
    for sh in range(sh_total):
        for sw in range(sw_total):
            h_l, w_l = window_origin(sh, sw, sh_total, sw_total, H, W)
            Apy = Apy_temp[:, :, h_l:h_l + 256, w_l:w_l + 256]
            times = get_schedule_jump(**jump)
            for t_last, t_cur in zip(times[:-1], times[1:]):
                if t_cur < t_last:
                    t = t_last
                    tt = torch.full((B,), t, dtype=torch.long)
                    out = model(x, tmap[tt], classes)
                    eps = out[:, :3]
                    x0_t = f32(K.sqrt_recip_alphas_cumprod, t) * x - f32(K.sqrt_recipm1_alphas_cumprod, t) * eps      # :404-411
                    if clip_denoised:
                        x0_t = x0_t.clamp(-1, 1)
                    sigma_t = torch.sqrt(f32(K.posterior_variance, t))
                    a_t = f32(K.posterior_mean_coef1, t)
                    if sigma_t >= a_t * sigma_y:                                                                 # Eq. 19 (:330-336)
                        lambda_t = 1
                        gamma_t = f32(K.posterior_variance, t) - (a_t * lambda_t * sigma_y) ** 2
                    else:
                        lambda_t = sigma_t / a_t * sigma_y
                        gamma_t = 0.
                    x0_hat = lambda_t * Apy + x0_t - lambda_t * Ap(A(x0_t))                                      # Eq. 17 (:340)
                    shift_overwrite(x0_hat, final, sh, sw, sh_total, sw_total, H, W)
                    mean = f32(K.posterior_mean_coef1, t) * x0_hat + f32(K.posterior_mean_coef2, t) * x           # :227-232
                    if cond_fn is not None:
                        mean = mean.float() + gamma_t * cond_fn(x, tmap[tt], classes).float()                     # :414-430
                    z = noise.pop(0)
                    nonzero = 0.0 if t == 0 else 1.0
                    x = mean + nonzero * torch.sqrt(torch.ones(1) * gamma_t) * z                                  # :487-488
                    if trace is not None:
                        trace.append(dict(sh=sh, sw=sw, t=t, x0_hat=x0_hat.clone(), x=x.clone()))
                else:
                    t = t_last + 1                                                                               # inpa_inj_time_shift = 1
                    beta = f32(K.betas, t)
                    x = torch.sqrt(1 - beta) * x + torch.sqrt(beta) * noise.pop(0)                                # :211-217
            final[:, :, h_l:h_l + 256, w_l:w_l + 256] = x0_hat
    return final


def count_draws(H, W, jump):
    """number of (B,3,256,256) Gaussian draws `restore` consumes for an (H, W) canvas"""
    times = get_schedule_jump(**jump)
    per_window = len(times) - 1
    return 1 + (math.ceil(H / 128) - 1) * (math.ceil(W / 128) - 1) * per_window
