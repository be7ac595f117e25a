"""Drop-in degradation operators with the reference's constructor signatures and the ``A_functions`` call contract
(functions/svd_operators.py:9-97): ``A``, ``A_pinv``, ``Lambda``, ``Lambda_noise`` on (B, .) CUDA fp32 tensors.

Constructors repeat the reference's *init-time* arithmetic (tiny SVDs, kernels, sort) with the same torch calls on
the same device, so LAPACK/cuSOLVER-dependent artefacts (``V_small`` complement bases, unstable-sort ``_perm``)
are identical to what the reference would hold; every per-step method runs in libddnm_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

KIND = dict(sr=0, color=1, inpaint=2, wh=3, deblur=4, srconv=5, denoise=6, deblur2d=7, cs=8, general=9)


def _host_f32(t):
    return None if t is None else np.ascontiguousarray(torch.as_tensor(t).detach().float().cpu().numpy())


def _host_i64(t):
    return None if t is None else np.ascontiguousarray(torch.as_tensor(t).detach().long().cpu().numpy())


class _Operator:
    """Shared ctypes plumbing.  Mirrors ``A_functions`` for the sampling path; the spectral accessors the sampler
    never calls (V, Vt, U, Ut, singulars, add_zeros) raise NotImplementedError exactly like the reference's base class."""

    def _create(self, kind, channels, img_dim, ratio=0, v_small=None, u_small=None, singulars=None, singulars_orig=None,
                perm=None, mask=None, v_small2=None, u_small2=None):
        self.channels, self.img_dim = channels, img_dim
        keep = [_host_f32(v_small), _host_f32(u_small), _host_f32(singulars), _host_f32(singulars_orig), _host_i64(perm),
                _host_i64(mask), _host_f32(v_small2), _host_f32(u_small2)]
        d = _lib.OperatorDesc()
        d.kind, d.channels, d.img_dim, d.ratio = KIND[kind], channels, img_dim, ratio
        for name, arr in zip(("v_small", "u_small", "singulars", "singulars_orig", "perm", "mask", "v_small2", "u_small2"), keep):
            setattr(d, name, None if arr is None else arr.ctypes.data)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().ddnm_operator_create(C.byref(d), C.byref(self._h)))
        self.y_dim = _lib.lib().ddnm_operator_y_dim(self._h)
        self.x_dim = img_dim if kind == "general" else channels * img_dim * img_dim

    @staticmethod
    def _prep(vec, dim):
        v = vec.reshape(vec.shape[0], -1)
        assert v.is_cuda, "ddnm_b200 operators run on CUDA tensors"
        assert v.shape[1] == dim, f"expected {dim} elements per row, got {v.shape[1]}"
        return v.float().contiguous()      # always a fresh or read-only view: caller tensors are never mutated

    def A(self, vec):
        x = self._prep(vec, self.x_dim)
        y = torch.empty(x.shape[0], self.y_dim, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().ddnm_operator_A(self._h, _lib.ptr(x), x.shape[0], _lib.ptr(y), _lib.cur_stream()))
        return y

    def A_pinv(self, vec):
        y = self._prep(vec, self.y_dim)
        x = torch.empty(y.shape[0], self.x_dim, device=y.device, dtype=torch.float32)
        _lib.check(_lib.lib().ddnm_operator_A_pinv(self._h, _lib.ptr(y), y.shape[0], _lib.ptr(x), _lib.cur_stream()))
        return x

    def project(self, x0, y):
        """x0 - A_pinv(A(x0) - y) in one pass (svd_ddnm.py:59-61)."""
        x = self._prep(x0, self.x_dim)
        yy = self._prep(y, self.y_dim)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().ddnm_operator_project(self._h, _lib.ptr(x), _lib.ptr(yy), x.shape[0], _lib.ptr(out), _lib.cur_stream()))
        return out.reshape(x0.shape)

    def Lambda(self, vec, a, sigma_y, sigma_t, eta):
        v = self._prep(vec, self.x_dim)
        out = torch.empty_like(v)
        _lib.check(_lib.lib().ddnm_operator_lambda(self._h, _lib.ptr(v), v.shape[0], float(a), float(sigma_y), float(sigma_t),
                                                  float(eta), _lib.ptr(out), _lib.cur_stream()))
        return out

    def Lambda_noise(self, vec, a, sigma_y, sigma_t, eta, epsilon):
        v = self._prep(vec, self.x_dim)
        e = self._prep(epsilon, self.x_dim)
        out = torch.empty_like(v)
        _lib.check(_lib.lib().ddnm_operator_lambda_noise(self._h, _lib.ptr(v), _lib.ptr(e), v.shape[0], float(a), float(sigma_y),
                                                        float(sigma_t), float(eta), _lib.ptr(out), _lib.cur_stream()))
        return out

    def At(self, vec):
        raise NotImplementedError()

    def V(self, vec):
        raise NotImplementedError()

    Vt = U = Ut = add_zeros = V

    def singulars(self):
        raise NotImplementedError()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().ddnm_operator_destroy(self._h)
        except Exception:
            pass


class SuperResolution(_Operator):
    """svd_operators.py:479-623 — ``SuperResolution(channels, img_dim, ratio, device)``."""

    def __init__(self, channels, img_dim, ratio, device, artefacts=None):
        assert img_dim % ratio == 0
        self.ratio, self.y_dim_side = ratio, img_dim // ratio
        if artefacts is None:
            A = torch.Tensor([[1 / ratio ** 2] * ratio ** 2]).to(device)
            self.U_small, self.singulars_small, self.V_small = torch.svd(A, some=False)
        else:
            self.U_small, self.singulars_small, self.V_small = artefacts
        self._create("sr", channels, img_dim, ratio, self.V_small, self.U_small, self.singulars_small)


class Colorization(_Operator):
    """svd_operators.py:627-736 — ``Colorization(img_dim, device)``."""

    def __init__(self, img_dim, device, artefacts=None):
        if artefacts is None:
            A = torch.Tensor([[0.3333, 0.3334, 0.3333]]).to(device)
            self.U_small, self.singulars_small, self.V_small = torch.svd(A, some=False)
        else:
            self.U_small, self.singulars_small, self.V_small = artefacts
        self._create("color", 3, img_dim, 0, self.V_small, self.U_small, self.singulars_small)


class Inpainting(_Operator):
    """svd_operators.py:324-439 — ``Inpainting(channels, img_dim, missing_indices, device)``; ``missing_indices``
    address the (pixel, channel)-interleaved vector (diffusion.py:466-470).  The reference's O(N*missing) python
    loop for ``kept_indices`` (:330) is replaced by a boolean complement (same index set)."""

    def __init__(self, channels, img_dim, missing_indices, device):
        n = channels * img_dim ** 2
        keep = torch.ones(n, dtype=torch.long)
        keep[torch.as_tensor(missing_indices).long().cpu()] = 0
        self.missing_indices = missing_indices
        self._create("inpaint", channels, img_dim, 0, mask=keep)


class WalshHadamardCS(_Operator):
    """svd_operators.py:211-320 — ``WalshHadamardCS(channels, img_dim, ratio, perm, device)``."""

    def __init__(self, channels, img_dim, ratio, perm, device):
        self.ratio, self.perm = ratio, perm
        self._create("wh", channels, img_dim, ratio, perm=perm)


class Deblurring(_Operator):
    """svd_operators.py:934-1091 — ``Deblurring(kernel, channels, img_dim, device, ZERO=3e-2)``."""

    def __init__(self, kernel, channels, img_dim, device, ZERO=3e-2, artefacts=None):
        if artefacts is None:
            # constructor arithmetic of svd_operators.py:944-962 (banded matrix uses taps i-k//2 .. i+k//2-1)
            A_small = torch.zeros(img_dim, img_dim, device=device)
            k = kernel.shape[0]
            for i in range(img_dim):
                for j in range(i - k // 2, i + k // 2):
                    if j < 0 or j >= img_dim:
                        continue
                    A_small[i, j] = kernel[j - i + k // 2]
            U, S, V = torch.svd(A_small, some=False)
            S_orig = S.clone()
            S[S < ZERO] = 0
            big_orig = torch.matmul(S_orig.reshape(img_dim, 1), S_orig.reshape(1, img_dim)).reshape(img_dim ** 2)
            big = torch.matmul(S.reshape(img_dim, 1), S.reshape(1, img_dim)).reshape(img_dim ** 2)
            big, perm = big.sort(descending=True)
            big_orig = big_orig[perm]
            artefacts = (U, V, big, big_orig, perm)
        self.U_small, self.V_small, self._singulars, self._singulars_orig, self._perm = artefacts
        self._create("deblur", channels, img_dim, 1, self.V_small, self.U_small, self._singulars, self._singulars_orig, self._perm)


class SRConv(_Operator):
    """svd_operators.py:851-931 — ``SRConv(kernel, channels, img_dim, device, stride=1)`` (bicubic SR).  Defines no
    Lambda: DDNM+ raises, as in the reference."""

    def __init__(self, kernel, channels, img_dim, device, stride=1, artefacts=None):
        if artefacts is None:
            small = img_dim // stride
            A_small = torch.zeros(small, img_dim, device=device)
            for i in range(stride // 2, img_dim + stride // 2, stride):
                for j in range(i - kernel.shape[0] // 2, i + kernel.shape[0] // 2):
                    je = j
                    if je < 0:
                        je = -je - 1
                    if je >= img_dim:
                        je = (img_dim - 1) - (je - img_dim)
                    A_small[i // stride, je] += kernel[j - i + kernel.shape[0] // 2]
            U, S, V = torch.svd(A_small, some=False)
            S[S < 3e-2] = 0
            artefacts = (U, S, V)
        self.U_small, self.singulars_small, self.V_small = artefacts
        self.ratio = stride
        self._create("srconv", channels, img_dim, stride, self.V_small, self.U_small, self.singulars_small)

    def Lambda(self, *a, **k):
        raise NotImplementedError()

    def Lambda_noise(self, *a, **k):
        raise NotImplementedError()


class Denoising(_Operator):
    """svd_operators.py:442-476 — ``Denoising(channels, img_dim, device)``."""

    def __init__(self, channels, img_dim, device):
        self._create("denoise", channels, img_dim)


def _band(kernel, img_dim, device):
    A_small = torch.zeros(img_dim, img_dim, device=device)
    for i in range(img_dim):
        for j in range(i - kernel.shape[0] // 2, i + kernel.shape[0] // 2):
            if j < 0 or j >= img_dim:
                continue
            A_small[i, j] = kernel[j - i + kernel.shape[0] // 2]
    return A_small


class Deblurring2D(_Operator):
    """svd_operators.py:1094-1166 — ``Deblurring2D(kernel1, kernel2, channels, img_dim, device)`` (deblur_aniso).
    Defines no Lambda: DDNM+ raises, as in the reference."""

    def __init__(self, kernel1, kernel2, channels, img_dim, device, artefacts=None):
        if artefacts is None:
            U1, S1, V1 = torch.svd(_band(kernel1, img_dim, device), some=False)
            U2, S2, V2 = torch.svd(_band(kernel2, img_dim, device), some=False)
            S1[S1 < 3e-2] = 0
            S2[S2 < 3e-2] = 0
            big = torch.matmul(S1.reshape(img_dim, 1), S2.reshape(1, img_dim)).reshape(img_dim ** 2)
            big, perm = big.sort(descending=True)
            artefacts = (U1, V1, U2, V2, big, perm)
        self.U_small1, self.V_small1, self.U_small2, self.V_small2, self._singulars, self._perm = artefacts
        self._create("deblur2d", channels, img_dim, 1, self.V_small1, self.U_small1, self._singulars, None, self._perm,
                     v_small2=self.V_small2, u_small2=self.U_small2)

    def Lambda(self, *a, **k):
        raise NotImplementedError()

    def Lambda_noise(self, *a, **k):
        raise NotImplementedError()


class CS(_Operator):
    """svd_operators.py:101-159 — ``CS(channels, img_dim, ratio, device)`` (cs_blockbased).  The 1024x1024 basis is the V of
    an SVD of ``torch.randn`` drawn from the global RNG, exactly as the reference constructor does.  No Lambda."""

    def __init__(self, channels, img_dim, ratio, device, artefacts=None):
        if artefacts is None:
            A = torch.randn(32 ** 2, 32 ** 2).to(device)
            _, _, V = torch.svd(A, some=False)
        else:
            V = artefacts
        self.V_small = V
        self.cs_size = int(32 * 32 * ratio)
        self._create("cs", channels, img_dim, self.cs_size, self.V_small)

    def Lambda(self, *a, **k):
        raise NotImplementedError()

    def Lambda_noise(self, *a, **k):
        raise NotImplementedError()


class GeneralA(_Operator):
    """svd_operators.py:173-208 — ``GeneralA(A)``: any dense degradation matrix A [m, n] (m <= n) through its full SVD
    (``torch.svd(A, some=False)``, singular values below 1e-3 zeroed, as the reference constructor does).  The
    reference calls it "memory inefficient": V is n x n, so it is only usable for small n.  No Lambda."""

    def __init__(self, A, artefacts=None):
        if artefacts is None:
            U, S, V = torch.svd(A, some=False)
            S = S.clone()
            S[S < 1e-3] = 0
        else:
            U, S, V = artefacts
        m, n = int(U.shape[0]), int(V.shape[0])
        if m > n:
            raise ValueError("GeneralA needs m <= n (the reference's mat_by_vec shapes only work for wide A)")
        self._U, self._singulars, self._V = U, S, V
        self._create("general", 1, n, m, V, U, S)

    def Lambda(self, *a, **k):
        raise NotImplementedError()

    def Lambda_noise(self, *a, **k):
        raise NotImplementedError()
