"""Drop-in for ``functions/svd_ddnm.py``: ``ddnm_diffusion`` (:19-78) and ``ddnm_plus_diffusion`` (:80-164) with the
reference's signatures and return convention (``([x_0.cpu()], [x0_pred.cpu()])``).

The whole loop runs inside libddnm_b200.so on the current CUDA stream without host round trips.  The Gaussian draws
are taken from torch's generator in the reference's order (one ``randn_like`` per time pair, :65/:74) into a tape
before the loop starts, so a run is seed-for-seed comparable with the reference on the same device.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import _EngineModel
from .operators import CS, Deblurring2D, GeneralA, SRConv, _Operator
from .schedule import alpha_bar_table, time_pairs

class_num = 951


def sample_device(x, model, b, eta, A_funcs, y, sigma_y, plus, config, noise=None, cls_fn=None):
    """The loop with device-resident inputs and outputs (no host copies): returns (x_0, x0_pred) CUDA tensors."""
    return _run(x, model, b, eta, A_funcs, y, sigma_y, plus, cls_fn, None, config, noise, to_host=False)


def _run(x, model, b, eta, A_funcs, y, sigma_y, plus, cls_fn, classes, config, noise=None, to_host=True):
    if not isinstance(model, _EngineModel):
        model = getattr(model, "module", model)          # tolerate nn.DataParallel-style wrappers
    if not isinstance(model, _EngineModel) or not isinstance(A_funcs, _Operator):
        raise TypeError("ddnm_b200.sampler needs a ddnm_b200.model denoiser and a ddnm_b200.operators operator")
    if plus and isinstance(A_funcs, (SRConv, Deblurring2D, CS, GeneralA)):
        # these operators define no Lambda / Lambda_noise: the reference fails at its first step with the base class's
        # NotImplementedError (svd_operators.py:93-97)
        raise NotImplementedError()
    with torch.no_grad():
        if not x.is_cuda:
            x = x.to("cuda", non_blocking=True)            # the reference moves xs[-1] to 'cuda' itself (svd_ddnm.py:45)
        n = x.size(0)
        pairs = time_pairs(config.diffusion.num_diffusion_timesteps, config.time_travel.T_sampling,
                           config.time_travel.travel_length, config.time_travel.travel_repeat)
        abar = np.ascontiguousarray(alpha_bar_table(b).numpy())
        ti = np.ascontiguousarray(np.array([p[0] for p in pairs], dtype=np.int32))
        tj = np.ascontiguousarray(np.array([p[1] for p in pairs], dtype=np.int32))
        x = x.float().contiguous()
        if noise is None:
            noise = torch.empty((len(pairs),) + tuple(x.shape), device=x.device, dtype=torch.float32)
            for k in range(len(pairs)):
                noise[k] = torch.randn_like(x)              # same generator consumption as the reference loop
        else:
            assert noise.shape == (len(pairs),) + tuple(x.shape)
            noise = noise.to(x.device).float().contiguous()
        yv = y.reshape(n, -1).to(x.device, non_blocking=True).float().contiguous()
        assert yv.shape[1] == A_funcs.y_dim, f"y has {yv.shape[1]} entries per image, operator expects {A_funcs.y_dim}"
        s = _lib.Schedule()
        s.n_pairs, s.t_i, s.t_j, s.abar = len(pairs), ti.ctypes.data, tj.ctypes.data, abar.ctypes.data
        s.num_timesteps, s.eta, s.sigma_y = int(config.diffusion.num_diffusion_timesteps), float(eta), float(sigma_y)
        s.plus = 1 if plus else 0
        out = torch.empty_like(x)
        x0p = torch.empty_like(x)
        if cls_fn is None:
            _lib.check(_lib.lib().ddnm_sample(model.engine(n), A_funcs._h, C.byref(s), _lib.ptr(x), _lib.ptr(yv), _lib.ptr(noise), n,
                                             _lib.ptr(out), _lib.ptr(x0p), _lib.cur_stream()))
        else:
            _guided(x, model, A_funcs, s, yv, noise, n, cls_fn, out, x0p)
        if not to_host:
            return out, x0p
        return [out.to("cpu")], [x0p.to("cpu")]


CLASS_NUM = 951   # functions/svd_ddnm.py:7


def _guided(x, model, A_funcs, sched, yv, noise, n, cls_fn, out, x0p):
    """Classifier-guided loop (svd_ddnm.py:48-52, :109-113).  As in the reference, the caller's ``classes`` are replaced by
    ``class_num`` for every row, the denoiser is called as ``model(xt, t, classes)``, only channels 0..2 of its output are
    kept, and ``cls_fn`` is evaluated at ``x`` — the function's INPUT, not the current iterate.  ``cls_fn`` (the classifier's
    autograd gradient, diffusion.py:181-189) is the caller's PyTorch callable; everything else runs in libddnm_b200.so."""
    assert model.num_classes is not None, "must specify y if and only if the model is class-conditional"   # unet.py:644-646
    classes = torch.ones(n, dtype=torch.long, device=x.device) * CLASS_NUM
    labels = classes.to(torch.int32)
    grad = torch.empty_like(x)
    failure = []

    def cb(_user, _k, t, _stream):
        try:
            tt = torch.ones(n, device=x.device) * t
            with torch.enable_grad():
                g = cls_fn(x, tt, classes)
            grad.copy_(g.reshape(grad.shape))
            return 0
        except BaseException as e:   # noqa: BLE001 — re-raised after the C call returns
            failure.append(e)
            return 1
    fn = _lib.GuidanceFn(cb)
    rc = _lib.lib().ddnm_sample_guided(model.engine(n), A_funcs._h, C.byref(sched), _lib.ptr(x), _lib.ptr(yv), _lib.ptr(noise), n,
                                       _lib.ptr(labels), _lib.ptr(grad), fn, None, _lib.ptr(out), _lib.ptr(x0p), _lib.cur_stream())
    if failure:
        raise failure[0]
    _lib.check(rc)


def ddnm_diffusion(x, model, b, eta, A_funcs, y, cls_fn=None, classes=None, config=None, noise=None):
    return _run(x, model, b, eta, A_funcs, y, 0.0, False, cls_fn, classes, config, noise)


def ddnm_plus_diffusion(x, model, b, eta, A_funcs, y, sigma_y, cls_fn=None, classes=None, config=None, noise=None):
    return _run(x, model, b, eta, A_funcs, y, sigma_y, True, cls_fn, classes, config, noise)


# ------------------------------------------------------------------------------------------------------------------
# The runner's "simplified" DDNM+ (guided_diffusion/diffusion.py:211-415): the reference inlines this loop in
# Diffusion.simplified_ddnm_plus; here it is a function with the same ingredients.
# ------------------------------------------------------------------------------------------------------------------
class SimplifiedDegradation:
    """A / Ap of diffusion.py:244-290 for ``args.deg`` in {colorization, denoising, sr_averagepooling, inpainting,
    mask_color_sr, diy}; ``mask`` is the (H, W) 0/1 array of exp/inp_masks/mask.npy."""

    def __init__(self, deg, deg_scale=1, mask=None, image_size=256, device="cuda"):
        table = {"colorization": (0, 1, 1), "denoising": (0, 0, 1), "sr_averagepooling": (0, 0, None), "inpainting": (1, 0, 1),
                 "mask_color_sr": (1, 1, None), "diy": (1, 1, None)}
        if deg not in table:
            raise NotImplementedError("degradation type not supported")
        use_mask, use_gray, sc = table[deg]
        self.scale = int(round(deg_scale)) if sc is None else sc
        self.image_size = image_size
        self._mask = None
        if use_mask:
            assert mask is not None, "this degradation needs the inpainting mask"
            self._mask = torch.as_tensor(mask).to(device=device, dtype=torch.float32).reshape(image_size, image_size).contiguous()
        d = _lib.SimpleDeg()
        d.use_mask, d.use_gray, d.scale, d.img_dim, d.channels = use_mask, use_gray, self.scale, image_size, 3
        d.mask = None if self._mask is None else self._mask.data_ptr()
        self._d = d

    def A(self, z):
        z = z.float().contiguous()
        s = self.image_size // self.scale
        y = torch.empty(z.shape[0], 3, s, s, device=z.device, dtype=torch.float32)
        _lib.check(_lib.lib().ddnm_simplified_A(C.byref(self._d), _lib.ptr(z), z.shape[0], _lib.ptr(y), _lib.cur_stream()))
        return y

    def Ap(self, y):
        y = y.float().contiguous()
        x = torch.empty(y.shape[0], 3, self.image_size, self.image_size, device=y.device, dtype=torch.float32)
        _lib.check(_lib.lib().ddnm_simplified_Ap(C.byref(self._d), _lib.ptr(y), y.shape[0], _lib.ptr(x), _lib.cur_stream()))
        return x


def simplified_ddnm_plus(x, model, b, eta, degradation, y, sigma_y, config=None, noise=None):
    """x: x_T (B,3,H,W); y = degradation.A(x_orig); sigma_y already doubled (diffusion.py:292).  Returns
    ``([x_0.cpu()], [x0_pred.cpu()])`` like the SVD samplers."""
    if not isinstance(model, _EngineModel):
        model = getattr(model, "module", model)
    if not isinstance(model, _EngineModel) or not isinstance(degradation, SimplifiedDegradation):
        raise TypeError("simplified_ddnm_plus needs a ddnm_b200.model denoiser and a SimplifiedDegradation")
    with torch.no_grad():
        if not x.is_cuda:
            x = x.to("cuda", non_blocking=True)
        n = x.size(0)
        pairs = time_pairs(config.diffusion.num_diffusion_timesteps, config.time_travel.T_sampling,
                           config.time_travel.travel_length, config.time_travel.travel_repeat)
        abar = np.ascontiguousarray(alpha_bar_table(b).numpy())
        ti = np.ascontiguousarray(np.array([p[0] for p in pairs], dtype=np.int32))
        tj = np.ascontiguousarray(np.array([p[1] for p in pairs], dtype=np.int32))
        x = x.float().contiguous()
        if noise is None:
            noise = torch.empty((len(pairs),) + tuple(x.shape), device=x.device, dtype=torch.float32)
            for k in range(len(pairs)):
                noise[k] = torch.randn_like(x)
        else:
            noise = noise.to(x.device).float().contiguous()
        yv = y.to(x.device, non_blocking=True).float().contiguous()
        s = _lib.Schedule()
        s.n_pairs, s.t_i, s.t_j, s.abar = len(pairs), ti.ctypes.data, tj.ctypes.data, abar.ctypes.data
        s.num_timesteps, s.eta, s.sigma_y, s.plus = int(config.diffusion.num_diffusion_timesteps), float(eta), float(sigma_y), 1
        out, x0p = torch.empty_like(x), torch.empty_like(x)
        _lib.check(_lib.lib().ddnm_sample_simplified(model.engine(n), C.byref(degradation._d), C.byref(s), _lib.ptr(x), _lib.ptr(yv),
                                                    _lib.ptr(noise), n, _lib.ptr(out), _lib.ptr(x0p), _lib.cur_stream()))
        return [out.to("cpu")], [x0p.to("cpu")]
