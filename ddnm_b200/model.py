"""Drop-in for the reference denoiser ``guided_diffusion.models.Model`` (models.py:192-341).

    model = Model(config)                 # same config Namespace the reference takes
    model.load_state_dict(state_dict)     # same keys/layout as the reference checkpoint
    et = model(xt, t)                     # same call as functions/svd_ddnm.py:47

The forward pass runs entirely inside libddnm_b200.so (CUDA graph of hand-written sm_100a kernels); torch is
only used for tensor storage and the current stream.  Inference only (the reference samples under no_grad).
"""
import ctypes as C
import math

import torch

from . import _lib


class Model:
    def __init__(self, config):
        m = config.model
        self.ch, self.out_ch = int(m.ch), int(m.out_ch)
        self.ch_mult = tuple(int(v) for v in m.ch_mult)
        self.num_res_blocks = int(m.num_res_blocks)
        self.attn_resolutions = tuple(int(v) for v in m.attn_resolutions)
        self.in_channels = int(m.in_channels)
        self.resolution = int(config.data.image_size)
        assert getattr(m, "resamp_with_conv", True), "only resamp_with_conv=True (the shipped configs) is built"
        self.config = config
        self._sd = None
        self._engines = {}      # batch -> handle
        self.use_cuda_graph = True
        _lib.lib()              # fail early if the CUDA library is absent

    # --- torch.nn.Module-like surface used by the reference runner (diffusion.py:117-140) ---
    def load_state_dict(self, sd, strict=True):
        self._sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v.detach().float().contiguous()
                    for k, v in sd.items()}
        self._destroy()
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def parameters(self):
        return iter(self._sd.values()) if self._sd else iter(())

    def _cfg(self):
        c = _lib.SimpleCfg()
        c.ch, c.out_ch, c.n_levels = self.ch, self.out_ch, len(self.ch_mult)
        for i, v in enumerate(self.ch_mult):
            c.ch_mult[i] = v
        c.num_res_blocks = self.num_res_blocks
        c.n_attn_res = len(self.attn_resolutions)
        for i, v in enumerate(self.attn_resolutions):
            c.attn_res[i] = v
        c.in_channels, c.resolution, c.groups, c.eps = self.in_channels, self.resolution, 32, 1e-6
        return c

    def engine(self, batch):
        if batch in self._engines:
            return self._engines[batch]
        if self._sd is None:
            raise _lib.DDNMError("Model has no weights: call load_state_dict first")
        L = _lib.lib()
        h = C.c_void_p()
        cfg = self._cfg()
        _lib.check(L.ddnm_unet_simple_create(C.byref(cfg), batch, C.byref(h)))
        # frequency table with the reference's own arithmetic (models.py:15-18)
        half = self.ch // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
        params = dict(self._sd)
        params["__freq"] = freq
        for name, t in params.items():
            t = t.contiguous()
            _lib.check(L.ddnm_unet_set_param(h, name.encode(), _lib.ptr(t), t.numel()))
        _lib.check(L.ddnm_unet_finalize(h))
        _lib.check(L.ddnm_unet_set_graph(h, 1 if self.use_cuda_graph else 0))
        self._engines[batch] = h
        return h

    def __call__(self, x, t):
        return self.forward(x, t)

    def forward(self, x, t):
        assert x.is_cuda and x.dtype == torch.float32, "ddnm_b200.Model runs on CUDA fp32 tensors"
        assert x.shape[2] == x.shape[3] == self.resolution     # models.py:302
        x = x.contiguous()
        t = t.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty(x.shape[0], self.out_ch, self.resolution, self.resolution, device=x.device, dtype=torch.float32)
        h = self.engine(x.shape[0])
        _lib.check(_lib.lib().ddnm_unet_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), _lib.cur_stream()))
        return out

    # --- extras ---
    def read_tap(self, batch, name, shape):
        out = torch.empty(shape, device="cuda", dtype=torch.float32)
        _lib.check(_lib.lib().ddnm_unet_read_tap(self.engine(batch), name.encode(), _lib.ptr(out), out.numel(), _lib.cur_stream()))
        return out

    def info(self, batch):
        ws, nl, fl = C.c_longlong(), C.c_int(), C.c_double()
        _lib.check(_lib.lib().ddnm_unet_info(self.engine(batch), C.byref(ws), C.byref(nl), C.byref(fl)))
        return dict(workspace_bytes=ws.value, launches=nl.value, flops_per_forward=fl.value)

    def profile(self, x, t):
        import json
        out = torch.empty(x.shape[0], self.out_ch, self.resolution, self.resolution, device=x.device, dtype=torch.float32)
        buf = C.create_string_buffer(1 << 20)
        _lib.check(_lib.lib().ddnm_unet_profile(self.engine(x.shape[0]), _lib.ptr(x.contiguous()),
                                               _lib.ptr(t.float().contiguous()), _lib.ptr(out), _lib.cur_stream(), buf, len(buf)))
        return json.loads(buf.value.decode())

    def _destroy(self):
        if self._engines:
            L = _lib.lib()
            for h in self._engines.values():
                L.ddnm_unet_destroy(h)
        self._engines = {}

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass
