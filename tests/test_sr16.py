"""16x average-pooling super-resolution with measurement noise (evaluation.sh: `--deg sr_averagepooling --deg_scale 16 --sigma_y 0.2
--add_noise`): a ratio outside the per-thread-patch kernels' 2 / 4 / 8.  CPU: the oracle against the reference's stored outputs;
GPU: the engine's generic-ratio path (patch rows + the 256 x 256 basis as a GEMM) against both, and the simplified runner loop."""
import numpy as np
import pytest
import torch

from oracle import operators as O
from oracle import sampler as S
from oracle import schedule as SCH
from oracle import unet_simple as U

from helpers import LAMBDA_CASES, assert_close, sampler_config

dev = "cuda"


@pytest.fixture(scope="module")
def g16(gold):
    return gold["sr16"]


def _oracle_op(g):
    return O.SuperResolution(3, 32, 16, torch.from_numpy(g["art_U_small"]), torch.from_numpy(g["art_singulars_small"]),
                             torch.from_numpy(g["art_V_small"]))


def _inputs():
    rng = torch.Generator().manual_seed(4321)
    x = torch.rand(2, 3, 32, 32, generator=rng) * 2 - 1
    v = torch.randn(2, 3 * 32 * 32, generator=rng)
    e = torch.randn(2, 3 * 32 * 32, generator=rng)
    return x, v, e


def test_oracle_sr16_matches_reference(g16):
    o = _oracle_op(g16)
    x, v, e = _inputs()
    y = o.A(x.reshape(2, -1))
    assert y.shape == (2, 3 * 2 * 2)
    assert np.abs(y.numpy() - g16["op_A"]).max() <= 4e-6
    yq = y * 0.9 + 0.05
    assert np.abs(o.A_pinv(yq.clone()).numpy() - g16["op_Apinv"]).max() <= 4e-6
    assert np.abs(o.project(x, yq).numpy() - g16["op_proj"]).max() <= 8e-6
    for ci, (a, sy, st) in enumerate(LAMBDA_CASES):
        at, stt = torch.tensor(a), torch.tensor(st)
        assert np.abs(o.Lambda(v.clone(), at, sy, stt, 0.85).numpy() - g16[f"op_L{ci}"]).max() <= 8e-6
        assert np.abs(o.Lambda_noise(v.clone(), at, sy, stt, 0.85, e.clone()).numpy() - g16[f"op_Ln{ci}"]).max() <= 8e-6


def _sampler_inputs(g):
    gen = torch.Generator().manual_seed(int(g["samp_seed"][0]))
    torch.rand(2, 3, 32, 32, generator=gen), torch.randn(2, 3, 32, 32, generator=gen)       # x_orig, x_T (stored)
    tape = [torch.randn(2, 3, 32, 32, generator=gen) for _ in range(6)]
    return torch.from_numpy(g["samp_x_T"]), torch.from_numpy(g["samp_y"]), tape


def test_oracle_sr16_sampler_matches_reference(g16):
    cfg = U.SimpleUNetConfig.tiny()
    sd = U.init_state_dict(cfg, 1234)
    x_T, y, tape = _sampler_inputs(g16)
    with torch.no_grad():
        x0, x0p = S.ddnm_sample(x_T, lambda a, b: U.forward(sd, a, b, cfg), SCH.linear_betas(), 0.85, _oracle_op(g16), y, tape,
                                t_sampling=6, sigma_y=0.4)
    assert np.abs(x0.numpy() - g16["samp_x0"]).max() <= 3e-3
    assert np.abs(x0p.numpy() - g16["samp_x0pred"]).max() <= 3e-3


@pytest.mark.gpu
def test_engine_sr16_operator_vs_oracle_and_reference(g16):
    from ddnm_b200 import operators as E
    o = _oracle_op(g16)
    eop = E.SuperResolution(3, 32, 16, dev, artefacts=(o.U_small, o.singulars_small, o.V_small))
    x, v, e = _inputs()
    xd, vd, ed = x.to(dev), v.to(dev), e.to(dev)
    y = o.A(x.reshape(2, -1))
    yq = y * 0.9 + 0.05
    assert_close(eop.A(xd), g16["op_A"], 1e-4, 1e-5, "sr16 A vs reference")
    assert_close(eop.A_pinv(yq.to(dev)), g16["op_Apinv"], 1e-4, 1e-5, "sr16 A_pinv vs reference")
    assert_close(eop.project(xd, yq.to(dev)).reshape(2, -1), g16["op_proj"].reshape(2, -1), 1e-4, 2e-5, "sr16 project vs reference")
    for ci, (a, sy, st) in enumerate(LAMBDA_CASES):
        at, stt = torch.tensor(a), torch.tensor(st)
        assert_close(eop.Lambda(vd, at, sy, stt, 0.85), g16[f"op_L{ci}"], 1e-4, 2e-5, f"sr16 Lambda{ci} vs reference")
        assert_close(eop.Lambda_noise(vd, at, sy, stt, 0.85, ed), g16[f"op_Ln{ci}"], 1e-4, 2e-5, f"sr16 Lambda_noise{ci} vs reference")
    assert torch.equal(vd.cpu(), v)
    # full size, ratio 16 and an odd one (ratio 32): properties
    big = E.SuperResolution(3, 256, 16, dev)
    xb = torch.rand(2, 3, 256, 256, device=dev) * 2 - 1
    yb = big.A(xb)
    assert yb.shape == (2, 3 * 16 * 16)
    assert_close(yb.reshape(2, 3, 16, 16), torch.nn.functional.avg_pool2d(xb, 16), 1e-4, 1e-5, "sr16@256 A = average pooling")
    assert_close(big.A(big.A_pinv(yb)), yb, 1e-4, 1e-5, "A A^+ y = y")
    p = big.project(torch.randn_like(xb), yb)
    assert_close(big.A(p), yb, 1e-3, 1e-4, "A(project(z, y)) = y")


@pytest.mark.gpu
def test_engine_sr16_ddnm_plus_sampler_vs_reference(g16):
    from ddnm_b200 import operators as E
    from ddnm_b200.sampler import ddnm_plus_diffusion
    from test_gpu_parity import _engine_model
    cfg = U.SimpleUNetConfig.tiny()
    m = _engine_model(cfg)
    o = _oracle_op(g16)
    eop = E.SuperResolution(3, 32, 16, dev, artefacts=(o.U_small, o.singulars_small, o.V_small))
    x_T, y, tape = _sampler_inputs(g16)
    xs, x0s = ddnm_plus_diffusion(x_T.to(dev), m, SCH.linear_betas().to(dev), 0.85, eop, y.to(dev), 0.4, config=sampler_config(6, 1, 1),
                                  noise=torch.stack(tape).to(dev))
    assert_close(xs[0], g16["samp_x0"], 1e-3, 3e-3, "sr16 DDNM+ x_0 vs reference")
    assert_close(x0s[0], g16["samp_x0pred"], 1e-3, 3e-3, "sr16 DDNM+ x0_pred vs reference")


@pytest.mark.gpu
def test_simplified_sr16_vs_reference_runner(gold):
    """README quick-start loop with deg_scale 16, sigma_y 0.2 (diffusion.py:211-415)."""
    from ddnm_b200.sampler import SimplifiedDegradation, simplified_ddnm_plus
    from test_gpu_parity import _engine_model
    from test_oracle_golden import simplified_inputs
    deg, scale, sy, T, tl, tr = "sr_averagepooling", 16, 0.2, 3, 1, 1
    cfg = U.SimpleUNetConfig.celeba_hq()
    m = _engine_model(cfg)
    x_T, x_orig, mask, tape = simplified_inputs(gold["simplified"], T, tl, tr)
    D = SimplifiedDegradation(deg, scale, mask, 256)
    y = D.A(x_orig.to(dev))
    assert_close(y, torch.nn.functional.avg_pool2d(x_orig, 16), 1e-4, 1e-5, "simplified sr16: A")
    xs, _ = simplified_ddnm_plus(x_T.to(dev), m, SCH.linear_betas().to(dev), 0.85, D, y, 2 * sy, config=sampler_config(T, tl, tr),
                                 noise=torch.stack(tape).to(dev))
    img = torch.clamp((xs[0] + 1.0) / 2.0, 0.0, 1.0)
    assert_close(img[:, :, ::4, ::4], gold["simplified_r2"][f"{deg}_s{scale}_sy{sy}_T{T}_l{tl}_r{tr}_img_s4"], 1e-3, 5e-4,
                 "simplified sr16 vs reference runner")
