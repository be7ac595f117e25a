"""Oracle: image-space closed forms of the reference's SVD degradation operators.

Each class restates one operator of /root/reference/functions/svd_operators.py through the public
contract of ``A_functions`` (:9-97): ``A``, ``A_pinv``, ``Lambda``, ``Lambda_noise`` on (B, .) fp32
tensors.  The reference factors every operator as U S V^T and shuffles data between "spectral"
orderings; here the same linear maps are written directly on the (B, C, H, W) image, keeping the
reference's quirks.  Artefacts that depend on LAPACK / RNG / unstable sort
(``V_small`` bases, ``perm``) are INPUTS — produced once by the reference constructors (or by the
``make_*`` helpers below, which repeat the constructor arithmetic) and shared with the CUDA engine.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Coefficient rules shared by every Lambda / Lambda_noise (e.g. svd_operators.py:568-604).
# ``a`` and ``sigma_t`` are 0-dim tensors or floats, ``sigma_y`` and ``eta`` python floats.
# ----------------------------------------------------------------------------------------------
def lambda_coeff(singulars, a, sigma_y, sigma_t, eta):
    s = singulars
    inv = 1.0 / s
    inv[s == 0] = 0.0
    lam = torch.ones_like(s)
    if a != 0 and sigma_y != 0:
        ci = (sigma_t < a * sigma_y * inv) * 1.0
        lam = lam * (-ci + 1.0) + ci * (s * sigma_t * (1 - eta ** 2) ** 0.5 / a / sigma_y)
    return lam


def noise_coeff(singulars, a, sigma_y, sigma_t, eta):
    s = singulars
    inv = 1.0 / s
    inv[s == 0] = 0.0
    d1 = torch.ones_like(s) * sigma_t * eta
    d2 = torch.ones_like(s) * sigma_t * (1 - eta ** 2) ** 0.5
    if a != 0 and sigma_y != 0:
        ci = (sigma_t < a * sigma_y * inv) * 1.0
        d1 = d1 * (-ci + 1.0) + ci * sigma_t * eta
        d2 = d2 * (-ci + 1.0)
        ci = (sigma_t > a * sigma_y * inv) * 1.0
        d1 = d1 * (-ci + 1.0) + torch.sqrt(ci * (sigma_t ** 2 - a ** 2 * sigma_y ** 2 * inv ** 2))
        d2 = d2 * (-ci + 1.0)
        ci = (s == 0) * 1.0
        d1 = d1 * (-ci + 1.0) + ci * sigma_t * eta
        d2 = d2 * (-ci + 1.0) + ci * sigma_t * (1 - eta ** 2) ** 0.5
    return d1, d2


def _pad_singulars(s, n):
    out = torch.zeros(n, dtype=s.dtype)
    out[: s.numel()] = s
    return out


class OracleOperator:
    channels = 3
    img_dim = 256

    def _img(self, v):
        return v.reshape(v.shape[0], self.channels, self.img_dim, self.img_dim)

    def project(self, x0, y):
        """x0 - A^+(A x0 - y): the null-space projection of svd_ddnm.py:59-61."""
        b = x0.shape[0]
        return x0 - self.A_pinv(self.A(x0.reshape(b, -1)) - y.reshape(b, -1)).reshape(x0.shape)


# ----------------------------------------------------------------------------------------------
class SuperResolution(OracleOperator):
    """svd_operators.py:479-623.  A = r x r block mean, A^+ = replicate.  Lambda acts in the r^2-dim
    per-patch basis ``V_small`` (LAPACK-dependent complement; column 0 = +-1/r)."""

    def __init__(self, channels, img_dim, ratio, U_small, singulars_small, V_small):
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.y_dim = img_dim // ratio
        self.U_small, self.singulars_small, self.V_small = U_small, singulars_small, V_small

    @staticmethod
    def make(channels, img_dim, ratio):
        # constructor arithmetic of svd_operators.py:486-488
        A = torch.Tensor([[1 / ratio ** 2] * ratio ** 2])
        U, S, V = torch.svd(A, some=False)
        return SuperResolution(channels, img_dim, ratio, U, S, V)

    def _patches(self, v):   # (B, C, y, y, r*r), patch entries row-major (unfold order, :510-512)
        r, yd = self.ratio, self.y_dim
        x = self._img(v).reshape(-1, self.channels, yd, r, yd, r).permute(0, 1, 2, 4, 3, 5)
        return x.reshape(-1, self.channels, yd, yd, r * r)

    def _unpatch(self, p):   # inverse of _patches -> (B, C*H*W)   (:504-507)
        r, yd = self.ratio, self.y_dim
        x = p.reshape(-1, self.channels, yd, yd, r, r).permute(0, 1, 2, 4, 3, 5)
        return x.reshape(p.shape[0], -1)

    def A(self, v):
        coeff = self._patches(v) @ self.V_small[:, 0]                 # Vt, first spectral row
        return (self.U_small[0, 0] * (self.singulars_small[0] * coeff)).reshape(v.shape[0], -1)

    def A_pinv(self, y):
        yd = self.y_dim
        c = (self.U_small[0, 0] * y.reshape(-1, self.channels, yd, yd)) * (1.0 / self.singulars_small[0])
        p = c[..., None] * self.V_small[:, 0]                          # V applied to (c, 0, ..., 0)
        return self._unpatch(p)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        lam = lambda_coeff(_pad_singulars(self.singulars_small, self.ratio ** 2), a, sigma_y, sigma_t, eta)
        spec = self._patches(v) @ self.V_small                         # Vt
        return self._unpatch((spec * lam) @ self.V_small.t())          # V

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        d1, d2 = noise_coeff(_pad_singulars(self.singulars_small, self.ratio ** 2), a, sigma_y, sigma_t, eta)
        # NB the reference scales RAW patch pixels (no Vt first) and then applies V (:581-621)
        pv = (self._patches(v) * d1) @ self.V_small.t()
        pe = (self._patches(eps) * d2) @ self.V_small.t()
        return self._unpatch(pv) + self._unpatch(pe)


# ----------------------------------------------------------------------------------------------
class Colorization(OracleOperator):
    """svd_operators.py:627-736.  Per pixel A = [0.3333 0.3334 0.3333], 3x3 basis ``V_small``."""

    def __init__(self, img_dim, U_small, singulars_small, V_small):
        self.channels, self.img_dim = 3, img_dim
        self.U_small, self.singulars_small, self.V_small = U_small, singulars_small, V_small

    @staticmethod
    def make(img_dim):
        A = torch.Tensor([[0.3333, 0.3334, 0.3333]])
        U, S, V = torch.svd(A, some=False)
        return Colorization(img_dim, U, S, V)

    def _needles(self, v):   # (B, HW, 3)
        return v.reshape(v.shape[0], 3, -1).permute(0, 2, 1)

    def _unneedle(self, n):
        return n.permute(0, 2, 1).reshape(n.shape[0], -1)

    def A(self, v):
        c = self._needles(v) @ self.V_small[:, 0]
        return self.U_small[0, 0] * (self.singulars_small[0] * c)

    def A_pinv(self, y):
        c = (self.U_small[0, 0] * y.reshape(y.shape[0], -1)) * (1.0 / self.singulars_small[0])
        return self._unneedle(c[..., None] * self.V_small[:, 0])

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        lam = lambda_coeff(_pad_singulars(self.singulars_small, 3), a, sigma_y, sigma_t, eta)
        return self._unneedle(((self._needles(v) @ self.V_small) * lam) @ self.V_small.t())

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        d1, d2 = noise_coeff(_pad_singulars(self.singulars_small, 3), a, sigma_y, sigma_t, eta)
        return self._unneedle((self._needles(v) * d1) @ self.V_small.t()) + \
            self._unneedle((self._needles(eps) * d2) @ self.V_small.t())


# ----------------------------------------------------------------------------------------------
class Inpainting(OracleOperator):
    """svd_operators.py:324-439 with the index construction of diffusion.py:464-471: ``mask`` is a flat
    (H*W,) array, 0 = missing; indices address the (pixel, channel)-interleaved layout.  Pure gather /
    scatter — must be bit-exact."""

    def __init__(self, channels, img_dim, mask):
        self.channels, self.img_dim = channels, img_dim
        m = torch.as_tensor(np.asarray(mask)).reshape(-1)
        self.keep_px = torch.nonzero(m != 0).reshape(-1)            # ascending pixel ids
        # kept entries of the interleaved vector, ascending: 3*p + c  (== reference kept_indices, :330)
        self.kept = (self.keep_px[:, None] * channels + torch.arange(channels)[None, :]).reshape(-1)
        self.mask_img = (m != 0).reshape(img_dim, img_dim)

    def _interleave(self, v):   # (B, HW*C) pixel-major
        return v.reshape(v.shape[0], self.channels, -1).permute(0, 2, 1).reshape(v.shape[0], -1)

    def _deinterleave(self, t):
        return t.reshape(t.shape[0], -1, self.channels).permute(0, 2, 1).reshape(t.shape[0], -1)

    def A(self, v):
        return self._interleave(v)[:, self.kept]

    def A_pinv(self, y):
        out = torch.zeros(y.shape[0], self.channels * self.img_dim ** 2, dtype=y.dtype)
        out[:, self.kept] = y.reshape(y.shape[0], -1)
        return self._deinterleave(out)

    def _sing_img(self):   # singular value per image element: 1 kept, 0 missing
        return self.mask_img.to(torch.float32)[None, None].expand(1, self.channels, -1, -1).reshape(-1)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        return v.reshape(v.shape[0], -1) * lambda_coeff(self._sing_img().clone(), a, sigma_y, sigma_t, eta)

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        d1, d2 = noise_coeff(self._sing_img().clone(), a, sigma_y, sigma_t, eta)
        b = v.shape[0]
        return v.reshape(b, -1) * d1 + eps.reshape(b, -1) * d2


# ----------------------------------------------------------------------------------------------
def fwht_natural(x):
    """Unnormalised natural-order Walsh-Hadamard butterfly over the last dim (svd_operators.py:212-222)."""
    n = x.shape[-1]
    lead = x.shape[:-1]
    h = 1
    while h < n:
        x = x.reshape(*lead, n // (2 * h), 2, h)
        x = torch.stack([x[..., 0, :] + x[..., 1, :], x[..., 0, :] - x[..., 1, :]], dim=-2)
        h *= 2
    return x.reshape(*lead, n)


class WalshHadamardCS(OracleOperator):
    """svd_operators.py:211-320.  ``perm``: (H*W,) permutation from diffusion.py:458 (global RNG)."""

    def __init__(self, channels, img_dim, ratio, perm):
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.perm = torch.as_tensor(perm).long()
        self.m = channels * img_dim ** 2 // ratio

    def _fwht(self, v):
        return fwht_natural(v.reshape(v.shape[0], self.channels, -1)) / self.img_dim

    def _spec(self, v):        # Vt: transform, permute positions, (pos, chan) interleave
        return self._fwht(v)[:, :, self.perm].permute(0, 2, 1).reshape(v.shape[0], -1)

    def _unspec(self, s):      # V
        t = torch.zeros(s.shape[0], self.channels, self.img_dim ** 2, dtype=s.dtype)
        t[:, :, self.perm] = s.reshape(s.shape[0], -1, self.channels).permute(0, 2, 1)
        return self._fwht(t).reshape(s.shape[0], -1)

    def A(self, v):
        return self._spec(v)[:, : self.m]

    def A_pinv(self, y):
        s = torch.zeros(y.shape[0], self.channels * self.img_dim ** 2, dtype=y.dtype)
        s[:, : self.m] = y.reshape(y.shape[0], -1)
        return self._unspec(s)

    def _sing(self):
        return _pad_singulars(torch.ones(self.m), self.channels * self.img_dim ** 2)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        return self._unspec(self._spec(v) * lambda_coeff(self._sing(), a, sigma_y, sigma_t, eta))

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        d1, d2 = noise_coeff(self._sing(), a, sigma_y, sigma_t, eta)
        b = v.shape[0]

        def raw(z):            # raw pixels taken as spectral coordinates (:289-293)
            return z.reshape(b, self.channels, -1)[:, :, self.perm].permute(0, 2, 1).reshape(b, -1)
        return self._unspec(raw(v) * d1) + self._unspec(raw(eps) * d2)


# ----------------------------------------------------------------------------------------------
class Deblurring(OracleOperator):
    """svd_operators.py:934-1091.  Separable: per channel X -> U_s (D_c o (V_s^T X V_s)) U_s^T with the
    reference's TILED singular table (``singulars()`` = _singulars.repeat(1,3), :1001) against the
    (pos, chan)-interleaved spectral vector: D[c, perm[p]] = S_sorted[(3p+c) mod n^2]."""

    def __init__(self, channels, img_dim, U_small, V_small, singulars_sorted, singulars_orig_sorted, perm):
        self.channels, self.img_dim = channels, img_dim
        self.U_small, self.V_small = U_small, V_small
        self.S, self.S_orig, self.perm = singulars_sorted, singulars_orig_sorted, torch.as_tensor(perm).long()
        n2 = img_dim ** 2
        tiled = self.S.repeat(1, channels).reshape(-1)                  # length C*n2, index 3p+c
        D = torch.zeros(channels, n2)
        D[:, self.perm] = tiled.reshape(n2, channels).t()
        self.D = D.reshape(channels, img_dim, img_dim)
        self.Dinv = torch.where(self.D == 0, torch.zeros_like(self.D), 1.0 / self.D)

    @staticmethod
    def make(kernel, channels, img_dim, ZERO=3e-2):
        # constructor arithmetic of svd_operators.py:944-962 (note: only the first 2*(k//2) taps are used)
        k = kernel.shape[0]
        A_small = torch.zeros(img_dim, img_dim)
        for i in range(img_dim):
            for j in range(i - k // 2, i + k // 2):
                if j < 0 or j >= img_dim:
                    continue
                A_small[i, j] = kernel[j - i + k // 2]
        U, S, V = torch.svd(A_small, some=False)
        S_orig = S.clone()
        S[S < ZERO] = 0
        big_orig = torch.matmul(S_orig.reshape(img_dim, 1), S_orig.reshape(1, img_dim)).reshape(-1)
        big = torch.matmul(S.reshape(img_dim, 1), S.reshape(1, img_dim)).reshape(-1)
        big, perm = big.sort(descending=True)
        return Deblurring(channels, img_dim, U, V, big, big_orig[perm], perm)

    def _sandwich(self, L, x, R):   # L @ X_c @ R for every (b, c)
        return torch.matmul(torch.matmul(L, x), R)

    def A(self, v):
        spec = self._sandwich(self.V_small.t(), self._img(v), self.V_small)
        out = self._sandwich(self.U_small, spec * self.D, self.U_small.t())
        # the reference returns U(...) in (C, H, W) order flattened
        return out.reshape(v.shape[0], -1)

    def A_pinv(self, y):
        spec = self._sandwich(self.U_small.t(), self._img(y), self.U_small)
        return self._sandwich(self.V_small, spec * self.Dinv, self.V_small.t()).reshape(y.shape[0], -1)

    def _table(self, coeff_sorted):  # per-position table (broadcast over channels): T[perm[p]] = coeff[p]
        t = torch.zeros(self.img_dim ** 2)
        t[self.perm] = coeff_sorted
        return t.reshape(self.img_dim, self.img_dim)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        lam = self._table(lambda_coeff(self.S_orig.clone(), a, sigma_y, sigma_t, eta))
        spec = self._sandwich(self.V_small.t(), self._img(v), self.V_small)
        return self._sandwich(self.V_small, spec * lam, self.V_small.t()).reshape(v.shape[0], -1)

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        d1, d2 = noise_coeff(self.S_orig.clone(), a, sigma_y, sigma_t, eta)
        b = v.shape[0]
        ov = self._sandwich(self.V_small, self._img(v) * self._table(d1), self.V_small.t())
        oe = self._sandwich(self.V_small, self._img(eps) * self._table(d2), self.V_small.t())
        return ov.reshape(b, -1) + oe.reshape(b, -1)


# ----------------------------------------------------------------------------------------------
class SRConv(OracleOperator):
    """svd_operators.py:851-931 (bicubic x ratio).  A(X) = M X M^T with M = U_s diag(S) V_s[:, :small]^T;
    no Lambda => DDNM+ unsupported (the base class raises NotImplementedError, :93-97)."""

    def __init__(self, channels, img_dim, ratio, U_small, singulars_small, V_small):
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.small = img_dim // ratio
        self.U_small, self.S_small, self.V_small = U_small, singulars_small, V_small
        self.S2 = torch.outer(singulars_small, singulars_small)         # (small, small)
        self.S2inv = torch.where(self.S2 == 0, torch.zeros_like(self.S2), 1.0 / self.S2)

    @staticmethod
    def bicubic_kernel(factor):
        # diffusion.py:485-497
        def cubic(x, a=-0.5):
            if abs(x) <= 1:
                return (a + 2) * abs(x) ** 3 - (a + 3) * abs(x) ** 2 + 1
            elif 1 < abs(x) and abs(x) < 2:
                return a * abs(x) ** 3 - 5 * a * abs(x) ** 2 + 8 * a * abs(x) - 4 * a
            return 0
        k = np.zeros((factor * 4))
        for i in range(factor * 4):
            x = (1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5)
            k[i] = cubic(x)
        k = k / np.sum(k)
        kernel = torch.from_numpy(k).float()
        return kernel / kernel.sum()

    @staticmethod
    def make(kernel, channels, img_dim, stride, ZERO=3e-2):
        small = img_dim // stride
        A_small = torch.zeros(small, img_dim)
        for i in range(stride // 2, img_dim + stride // 2, stride):
            for j in range(i - kernel.shape[0] // 2, i + kernel.shape[0] // 2):
                je = j
                if je < 0:
                    je = -je - 1
                if je >= img_dim:
                    je = (img_dim - 1) - (je - img_dim)
                A_small[i // stride, je] += kernel[j - i + kernel.shape[0] // 2]
        U, S, V = torch.svd(A_small, some=False)
        S[S < ZERO] = 0
        return SRConv(channels, img_dim, stride, U, S, V)

    def A(self, v):
        Vk = self.V_small[:, : self.small]
        spec = torch.matmul(torch.matmul(Vk.t(), self._img(v)), Vk) * self.S2
        out = torch.matmul(torch.matmul(self.U_small, spec), self.U_small.t())
        return out.reshape(v.shape[0], -1)

    def A_pinv(self, y):
        Vk = self.V_small[:, : self.small]
        yi = y.reshape(y.shape[0], self.channels, self.small, self.small)
        spec = torch.matmul(torch.matmul(self.U_small.t(), yi), self.U_small) * self.S2inv
        return torch.matmul(torch.matmul(Vk, spec), Vk.t()).reshape(y.shape[0], -1)

    def Lambda(self, *args):
        raise NotImplementedError()

    def Lambda_noise(self, *args):
        raise NotImplementedError()


# ----------------------------------------------------------------------------------------------
class Denoising(OracleOperator):
    """svd_operators.py:442-476.  A = identity; Lambda / Lambda_noise are scalar rules (epsilon is ignored)."""

    def __init__(self, channels, img_dim):
        self.channels, self.img_dim = channels, img_dim

    def A(self, v):
        return v.clone().reshape(v.shape[0], -1)

    def A_pinv(self, y):
        return y.clone().reshape(y.shape[0], -1)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        if sigma_t < a * sigma_y:
            factor = (sigma_t * (1 - eta ** 2) ** 0.5 / a / sigma_y).item()
            return v * factor
        return v

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        if sigma_t >= a * sigma_y:
            factor = torch.sqrt(sigma_t ** 2 - a ** 2 * sigma_y ** 2).item()
            return v * factor
        return v * sigma_t * eta


class Deblurring2D(OracleOperator):
    """svd_operators.py:1094-1166 (deblur_aniso): X -> U1 (D_c o (V1^T X V2)) U2^T with the same tiled-singular quirk as
    Deblurring (``singulars()`` = repeat(1, 3), :1162); base-class A_pinv; no Lambda."""

    def __init__(self, channels, img_dim, U1, V1, U2, V2, singulars_sorted, perm):
        self.channels, self.img_dim = channels, img_dim
        self.U1, self.V1, self.U2, self.V2 = U1, V1, U2, V2
        self.S, self.perm = singulars_sorted, torch.as_tensor(perm).long()
        n2 = img_dim ** 2
        tiled = self.S.repeat(1, channels).reshape(-1)
        D = torch.zeros(channels, n2)
        D[:, self.perm] = tiled.reshape(n2, channels).t()
        self.D = D.reshape(channels, img_dim, img_dim)
        self.Dinv = torch.where(self.D == 0, torch.zeros_like(self.D), 1.0 / self.D)

    @staticmethod
    def band(kernel, img_dim):
        k = kernel.shape[0]
        A = torch.zeros(img_dim, img_dim)
        for i in range(img_dim):
            for j in range(i - k // 2, i + k // 2):
                if 0 <= j < img_dim:
                    A[i, j] = kernel[j - i + k // 2]
        return A

    @staticmethod
    def make(kernel1, kernel2, channels, img_dim, ZERO=3e-2):
        U1, S1, V1 = torch.svd(Deblurring2D.band(kernel1, img_dim), some=False)
        U2, S2, V2 = torch.svd(Deblurring2D.band(kernel2, img_dim), some=False)
        S1[S1 < ZERO] = 0
        S2[S2 < ZERO] = 0
        big = torch.matmul(S1.reshape(img_dim, 1), S2.reshape(1, img_dim)).reshape(-1)
        big, perm = big.sort(descending=True)
        return Deblurring2D(channels, img_dim, U1, V1, U2, V2, big, perm)

    def A(self, v):
        spec = torch.matmul(torch.matmul(self.V1.t(), self._img(v)), self.V2) * self.D
        return torch.matmul(torch.matmul(self.U1, spec), self.U2.t()).reshape(v.shape[0], -1)

    def A_pinv(self, y):
        spec = torch.matmul(torch.matmul(self.U1.t(), self._img(y)), self.U2) * self.Dinv
        return torch.matmul(torch.matmul(self.V1, spec), self.V2.t()).reshape(y.shape[0], -1)

    def Lambda(self, *args):
        raise NotImplementedError()

    def Lambda_noise(self, *args):
        raise NotImplementedError()


class CS(OracleOperator):
    """svd_operators.py:101-159 (cs_blockbased): per 32x32 patch keep the first ``cs_size`` coefficients in the orthonormal
    basis ``V_small`` (1024 x 1024; the reference draws it as the V of an SVD of torch.randn); U = I, singulars = 1; no Lambda."""

    def __init__(self, channels, img_dim, ratio, V_small):
        self.channels, self.img_dim, self.V_small = channels, img_dim, V_small
        self.y_dim, self.p = img_dim // 32, 32
        self.cs_size = int(32 * 32 * ratio)

    def _patches(self, v):   # (B, C, y*y, 1024), row-major inside the patch (unfold order, :136-138)
        yd, p = self.y_dim, self.p
        x = self._img(v).reshape(-1, self.channels, yd, p, yd, p).permute(0, 1, 2, 4, 3, 5)
        return x.reshape(-1, self.channels, yd * yd, p * p)

    def A(self, v):
        spec = self._patches(v) @ self.V_small                      # row form of Vt_small @ patch
        return spec[..., : self.cs_size].reshape(v.shape[0], -1)

    def A_pinv(self, y):
        yd, p = self.y_dim, self.p
        c = y.reshape(y.shape[0], self.channels, yd * yd, self.cs_size)
        patches = c @ self.V_small[:, : self.cs_size].t()            # V_small @ (c, 0, ..)
        x = patches.reshape(-1, self.channels, yd, yd, p, p).permute(0, 1, 2, 4, 3, 5)
        return x.reshape(y.shape[0], -1)

    def Lambda(self, *args):
        raise NotImplementedError()

    def Lambda_noise(self, *args):
        raise NotImplementedError()


class GeneralA(OracleOperator):
    """svd_operators.py:173-208: dense A = U diag(s) V^T from ``torch.svd(A, some=False)`` with s < 1e-3 zeroed; batched
    mat-vec products (:174-179); add_zeros pads the m coefficients to n (:203-206).  No Lambda."""

    def __init__(self, U, S, V):
        self._U, self._S, self._V = U, S, V
        self.m, self.n = U.shape[0], V.shape[0]

    def A(self, v):
        v = v.reshape(v.shape[0], -1)
        temp = torch.matmul(self._V.t(), v.reshape(v.shape[0], self.n, 1)).reshape(v.shape[0], self.n)
        return torch.matmul(self._U, (self._S * temp[:, : self.m]).reshape(-1, self.m, 1)).reshape(v.shape[0], self.m)

    def A_pinv(self, y):
        y = y.reshape(y.shape[0], -1)
        temp = torch.matmul(self._U.t(), y.reshape(y.shape[0], self.m, 1)).reshape(y.shape[0], self.m)
        factors = 1.0 / self._S
        factors[self._S == 0] = 0.0
        out = torch.zeros(y.shape[0], self.n)
        out[:, : self.m] = temp * factors
        return torch.matmul(self._V, out.reshape(-1, self.n, 1)).reshape(y.shape[0], self.n)

    def Lambda(self, *args):
        raise NotImplementedError()

    def Lambda_noise(self, *args):
        raise NotImplementedError()


def hadamard_basis(n=1024):
    """A reproducible orthonormal 1024 x 1024 basis (Sylvester Hadamard / sqrt(n), exactly representable) used where tests
    need the SAME ``V_small`` on every machine; the reference draws its own from the global RNG."""
    H = torch.ones(1, 1)
    while H.shape[0] < n:
        H = torch.cat([torch.cat([H, H], 1), torch.cat([H, -H], 1)], 0)
    return H / (n ** 0.5)
