"""CPU: the oracle against the reference results stored in tests/golden/fullsize.npz — BASELINE's configs at their real size
(256x256, celeba `Model` / imagenet `UNetModel` with seeded random weights, the real exp/inp_masks/mask.npy) and the
`deblur_uni` operator (diffusion.py:500-503).  oracle/gen_golden.py::fullsize_fixtures produced the file from the unmodified
reference; these tests pin the oracle to it on machines without /root/reference."""
import numpy as np
import pytest
import torch

from oracle import fullsize as FS
from oracle import operators as O
from oracle import sampler as S
from oracle import schedule as SCH
from oracle import unet_openai as UO
from oracle import unet_simple as U

from helpers import LAMBDA_CASES

_NETS = {}


def oracle_net(kind):
    if kind not in _NETS:
        if kind == "celeba":
            cfg = U.SimpleUNetConfig.celeba_hq()
            sd = U.init_state_dict(cfg, 1234)
            _NETS[kind] = (cfg, sd, lambda a, b: U.forward(sd, a, b, cfg))
        else:
            cfg = UO.OpenAIUNetConfig.imagenet_256()
            sd = UO.init_state_dict(cfg, 1234)
            _NETS[kind] = (cfg, sd, lambda a, b: UO.forward(sd, a, b, cfg))
    return _NETS[kind]


@pytest.fixture(scope="module")
def full(gold):
    return gold["fullsize"]


@pytest.mark.parametrize("case", FS.FULLSIZE_CASES, ids=lambda c: c[0])
def test_oracle_fullsize_sampler_matches_reference(full, case):
    key, kind, opname, T, tl, tr, sy = case
    _, _, fwd = oracle_net(kind)
    op = FS.oracle_op(full, opname)
    npairs = len(SCH.time_pairs(1000, T, tl, tr))
    x_orig, x_T, tape, ynoise = FS.fullsize_inputs(key, npairs)
    y = FS.measurement(op, x_orig, ynoise, sy)
    with torch.no_grad():
        x0, x0p = S.ddnm_sample(x_T, fwd, SCH.linear_betas(), 0.85, op, y, tape, t_sampling=T, travel_length=tl, travel_repeat=tr,
                                sigma_y=sy)
    # cfg1 (20 steps): oracle-vs-reference rounding (the closed-form operator vs the reference's SVD plumbing, ~1e-6 per call) is
    # amplified to 3e-4 by the random-init net; the short schedules agree to rounding
    tol = 2e-3 if T >= 10 else 2e-5
    d0 = np.abs(x0[:, :, ::4, ::4].numpy() - full[key + "_x0_s4"]).max()
    d1 = np.abs(x0p[:, :, ::4, ::4].numpy() - full[key + "_x0pred_s4"]).max()
    assert d0 <= tol and d1 <= tol * max(1.0, np.abs(full[key + "_x0pred_s4"]).max()), (key, d0, d1)
    sums = full[key + "_sums"]
    assert abs(x0.double().sum().item() - sums[0]) <= 1e-3 * sums[1]
    if sy == 0.0:      # alpha-bar = 1 at the last step: x_0 is the projection itself, A x_0 = y (svd_ddnm.py:57-65)
        assert (op.A(x0.reshape(1, -1)) - y).abs().max().item() <= 1e-4


@pytest.mark.parametrize("dim", [32, 256])
def test_oracle_deblur_uni_matches_reference(full, dim):
    B = 2 if dim == 32 else 1
    rng = torch.Generator().manual_seed(4321)
    x = torch.rand(B, 3, dim, dim, generator=rng) * 2 - 1
    v = torch.randn(B, 3 * dim * dim, generator=rng)
    e = torch.randn(B, 3 * dim * dim, generator=rng)
    tag = f"d{dim}_deblur_uni"
    if dim == 32:
        a = lambda k: torch.from_numpy(full[f"{tag}_art_{k}"])     # noqa: E731
        o = O.Deblurring(3, 32, a("U_small"), a("V_small"), a("singulars"), a("singulars_orig"), a("perm"))
    else:
        o = FS.oracle_op(full, "deblur_uni")
    sub = (lambda z: z) if dim == 32 else (lambda z: z.reshape(B, -1)[:, ::61])
    y = o.A(x.reshape(B, -1))
    assert np.abs(sub(y).numpy() - full[f"{tag}_A"]).max() <= 4e-6
    yq = y * 0.9 + 0.05
    assert np.abs(sub(o.A_pinv(yq.clone())).numpy() - full[f"{tag}_Apinv"]).max() <= 4e-6
    assert np.abs(sub(o.project(x, yq)).numpy() - full[f"{tag}_proj"].reshape(sub(o.project(x, yq)).shape)).max() <= 8e-6
    for ci, (a_, sy, st) in enumerate(LAMBDA_CASES):
        at, stt = torch.tensor(a_), torch.tensor(st)
        assert np.abs(sub(o.Lambda(v.clone(), at, sy, stt, 0.85)).numpy() - full[f"{tag}_L{ci}"]).max() <= 8e-6
        assert np.abs(sub(o.Lambda_noise(v.clone(), at, sy, stt, 0.85, e.clone())).numpy() - full[f"{tag}_Ln{ci}"]).max() <= 8e-6


def test_fullsize_mask_is_the_reference_mask(full, gold):
    m = FS.mask_from_bits(full["mask_bits"])
    assert m.shape == (256, 256) and set(np.unique(m)) == {0, 1}
    assert np.array_equal(full["mask_bits"], gold["simplified"]["mask_bits"])      # the same exp/inp_masks/mask.npy
    assert 3 * int(m.sum()) == 145314                                              # M of the real mask
