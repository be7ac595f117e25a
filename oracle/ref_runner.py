"""Drive the reference's OWN code (the unmodified copy in oracle/_ref/, see oracle/make_ref.py) on the host cores: the
``--impl reference`` arm and the ``cpu_baseline`` leg of bench.py.  Public API used, exactly as the runner does
(guided_diffusion/diffusion.py:117-140, :481-484, :578-590):  ``Model(config)``, ``SuperResolution(channels, img, ratio, dev)``,
``ddnm_diffusion(x, model, betas, eta, A_funcs, y, config=config)``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import importlib
import os
import sys
import time
import types

import torch

from . import make_ref
from .ref_shim import cpu_shim

ns = types.SimpleNamespace


def _import_ref():
    if not make_ref.available():
        raise RuntimeError("oracle/_ref/ is missing: run `python -m oracle.make_ref` where /root/reference exists")
    if make_ref.DST not in sys.path:
        sys.path.insert(0, make_ref.DST)
    models = importlib.import_module("guided_diffusion.models")
    ops = importlib.import_module("functions.svd_operators")
    ddnm = importlib.import_module("functions.svd_ddnm")
    assert os.path.abspath(models.__file__).startswith(make_ref.DST), "guided_diffusion resolved outside oracle/_ref"
    return models, ops, ddnm


def celeba_config(T_sampling):
    """configs/celeba_hq.yml's model block + the time_travel block the runner adds (main.py)."""
    return ns(model=ns(type="simple", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2,
                       attn_resolutions=[16], dropout=0.0, var_type="fixedsmall", ema_rate=0.999, ema=True, resamp_with_conv=True),
              data=ns(image_size=256, channels=3), diffusion=ns(num_diffusion_timesteps=1000),
              time_travel=ns(T_sampling=T_sampling, travel_length=1, travel_repeat=1))


def time_reference_sr4(batch, T_sampling, threads, repeats=1, seed=1234):
    """Seconds per call of the reference's ddnm_diffusion for celeba_hq 4x sr_averagepooling (sigma_y = 0) on `threads` host
    threads: `batch` images, a `T_sampling`-step schedule (every step = one Model forward + the SVD projection + re-noising, the
    same work at every t).  Returns (seconds per call [list], |A x0 - y| of the last call)."""
    models, ops, ddnm = _import_ref()
    torch.set_num_threads(threads)
    cfg = celeba_config(T_sampling)
    torch.manual_seed(seed)
    model = models.Model(cfg).eval()
    A = ops.SuperResolution(3, 256, 4, "cpu")
    import numpy as np
    betas = torch.from_numpy(np.linspace(1e-4, 0.02, 1000, dtype=np.float64)).float()   # get_beta_schedule("linear"), diffusion.py:61-63,94
    g = torch.Generator().manual_seed(seed)
    x_orig = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1
    y = A.A(x_orig)
    secs, resid = [], None
    for _ in range(repeats):
        x = torch.randn(batch, 3, 256, 256, generator=g)
        with torch.no_grad(), cpu_shim():
            t0 = time.perf_counter()
            xs, _ = ddnm.ddnm_diffusion(x, model, betas, 0.85, A, y, config=cfg)
            secs.append(time.perf_counter() - t0)
        resid = (A.A(xs[0]).reshape(batch, -1) - y.reshape(batch, -1)).abs().max().item()
    return secs, resid
