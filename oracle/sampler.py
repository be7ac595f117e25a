"""Oracle: the DDNM / DDNM+ reverse-diffusion loops with an explicit noise tape.

Restates /root/reference/functions/svd_ddnm.py:19-78 (ddnm_diffusion) and :80-164
(ddnm_plus_diffusion).  The reference draws ``torch.randn_like`` once per time pair; here pair k uses
``noise[k]`` so that the CUDA engine and the oracle (and the reference, via a patched randn_like in
oracle/gen_golden.py) consume identical noise.  Device round trips (.to('cuda') / .to('cpu')) are dropped.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch

from .schedule import alpha_bar_table, time_pairs


CLASS_NUM = 951   # svd_ddnm.py:7 — the reference overwrites the caller's ``classes`` with this label (:49, :110)


def ddnm_sample(x_T, model, betas, eta, op, y, noise, num_timesteps=1000, t_sampling=100,
                travel_length=1, travel_repeat=1, sigma_y=0.0, trace=None, cls_fn=None):
    """Returns (x_0, last x0_pred).  ``model(xt, t)`` -> eps (channels beyond 3 dropped, :54-55).
    sigma_y == 0 -> DDNM (:57-65); sigma_y > 0 (already doubled by the caller, diffusion.py:524) -> DDNM+ (:114-131).
    ``noise``: callable k -> (B,3,H,W) tensor or an indexable tape.
    ``cls_fn`` (classifier guidance, :48-52 / :109-113): the model is then called as ``model(xt, t, classes)`` with
    classes = CLASS_NUM for every row, only channels 0..2 of its output are kept, and
    ``et -= sqrt(1 - at) * cls_fn(x, t, classes)`` where ``x`` is the function's INPUT (the initial noise x_T), not the current
    iterate — exactly what the reference code does."""
    abar = alpha_bar_table(betas)
    pairs = time_pairs(num_timesteps, t_sampling, travel_length, travel_repeat)
    n = x_T.shape[0]
    xt = x_T
    x0_last = None
    get_noise = noise if callable(noise) else (lambda k: noise[k])
    for k, (i, j) in enumerate(pairs):
        at_next = abar[j + 1]
        z = get_noise(k)
        if j < i:
            t = torch.ones(n) * i
            at = abar[i + 1]
            if cls_fn is None:
                et = model(xt, t)
            else:
                classes = torch.ones(n, dtype=torch.long) * CLASS_NUM
                et = model(xt, t, classes)[:, :3]
                et = et - (1 - at).sqrt() * cls_fn(x_T, t, classes)
            if et.size(1) == 6:
                et = et[:, :3]
            x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
            resid = op.A_pinv(op.A(x0_t.reshape(n, -1)) - y.reshape(n, -1))
            if sigma_y == 0.0:
                x0_hat = x0_t - resid.reshape(x0_t.shape)
                c1 = (1 - at_next).sqrt() * eta
                c2 = (1 - at_next).sqrt() * ((1 - eta ** 2) ** 0.5)
                xt_next = at_next.sqrt() * x0_hat + c1 * z + c2 * et
            else:
                sigma_t = (1 - at_next).sqrt()
                a = at_next.sqrt()
                x0_hat = x0_t - op.Lambda(resid.reshape(n, -1), a, sigma_y, sigma_t, eta).reshape(x0_t.shape)
                xt_next = a * x0_hat + op.Lambda_noise(z.reshape(n, -1), a, sigma_y, sigma_t, eta,
                                                       et.reshape(n, -1)).reshape(x0_t.shape)
            x0_last = x0_t
            if trace is not None:
                trace.append(dict(k=k, i=i, j=j, et=et.clone(), x0_t=x0_t.clone(), xt_next=xt_next.clone()))
        else:
            # time-travel back from the UN-projected x0_t (:69-76)
            xt_next = at_next.sqrt() * x0_last + z * (1 - at_next).sqrt()
            if trace is not None:
                trace.append(dict(k=k, i=i, j=j, xt_next=xt_next.clone()))
        xt = xt_next
    return xt, x0_last
