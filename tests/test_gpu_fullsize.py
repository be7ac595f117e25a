"""GPU (B200): parity at BASELINE's real sizes — 256x256, the celeba `Model` and the imagenet `UNetModel`, the real inpainting mask.

(1) every full-size case of tests/golden/fullsize.npz (results of the UNMODIFIED reference samplers) end to end on the engine
    with the same noise tape, including BASELINE configs[0] (celeba sr4, T=20, B=1);
(2) teacher-forced single steps through ddnm_diffusion / ddnm_plus_diffusion with both networks for sr4, colorization,
    inpainting (real mask.npy), Walsh-Hadamard CS and Gaussian deblurring at north_star's rtol 1e-3 / atol 1e-4 x scale;
(3) Inpainting with exp/inp_masks/mask.npy: A, A_pinv, project bit-exact (torch.equal);
(4) deblur_uni (diffusion.py:500-503) against the reference's stored outputs.
Measured drifts are appended to gpurun_out/fullsize_drift.jsonl."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fullsize as FS
from oracle import operators as O
from oracle import sampler as S
from oracle import schedule as SCH

from helpers import LAMBDA_CASES, assert_close, engine_op, sampler_config
from test_fullsize_oracle import oracle_net

pytestmark = pytest.mark.gpu
dev = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def full(gold):
    return gold["fullsize"]


_ENG = {}


def engine_net(kind):
    if kind not in _ENG:
        from test_gpu_parity import _engine_model, _engine_openai
        cfg, _, _ = oracle_net(kind)
        _ENG[kind] = _engine_model(cfg) if kind == "celeba" else _engine_openai(cfg)
    return _ENG[kind]


def engine_operator(name, oop):
    return engine_op("deblur" if name == "deblur_uni" else name, oop, 256)


def _log(rec):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fullsize_drift.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


@pytest.mark.parametrize("case", FS.FULLSIZE_CASES, ids=lambda c: c[0])
def test_fullsize_sampler_vs_reference(full, case):
    from ddnm_b200.sampler import ddnm_diffusion, ddnm_plus_diffusion
    key, kind, opname, T, tl, tr, sy = case
    m = engine_net(kind)
    oop = FS.oracle_op(full, opname)
    eop = engine_operator(opname, oop)
    npairs = len(SCH.time_pairs(1000, T, tl, tr))
    x_orig, x_T, tape, ynoise = FS.fullsize_inputs(key, npairs)
    y = FS.measurement(oop, x_orig, ynoise, sy)
    conf = sampler_config(T, tl, tr)
    betas = SCH.linear_betas().to(dev)
    noise = torch.stack(tape).to(dev)
    if sy == 0.0:
        xs, x0s = ddnm_diffusion(x_T.to(dev), m, betas, 0.85, eop, y.to(dev), config=conf, noise=noise)
    else:
        xs, x0s = ddnm_plus_diffusion(x_T.to(dev), m, betas, 0.85, eop, y.to(dev), sy, config=conf, noise=noise)
    ref0, ref1 = full[key + "_x0_s4"], full[key + "_x0pred_s4"]
    d0 = float(np.abs(xs[0][:, :, ::4, ::4].numpy() - ref0).max())
    d1 = float(np.abs(x0s[0][:, :, ::4, ::4].numpy() - ref1).max())
    resid = float((oop.A(xs[0].reshape(1, -1)) - y).abs().max()) if sy == 0.0 else None
    _log(dict(test="fullsize_sampler_vs_reference", case=key, T=T, pairs=npairs, drift_x0=d0, drift_x0pred=d1,
              x0_absmax=float(np.abs(ref0).max()), x0pred_absmax=float(np.abs(ref1).max()), data_residual=resid))
    # north_star's tolerance, rtol 1e-3 / atol 1e-4 with the absolute part following the tensor's magnitude: with random-init
    # weights x_0 / x0_pred are not image-scaled (|x| up to 475 for cfg1, 74 for the imagenet cases).  Measured drifts (logged
    # above; profiles/r02_fullsize_drift.jsonl): 1.2e-3 on |x| <= 475 after the 20 steps of cfg1 — the oracle itself sits 3e-4
    # from the reference there (gen_golden) — and 3e-5 .. 8e-4 on the short schedules.  The per-step tolerance is enforced
    # separately by test_fullsize_teacher_forced_steps (measured 2e-6 / 1.6e-5 of scale).
    sc0 = max(1.0, float(np.abs(ref0).max()))
    sc1 = max(1.0, float(np.abs(ref1).max()))
    assert_close(xs[0][:, :, ::4, ::4], ref0, 1e-3, 1e-4 * sc0, f"{key}: x_0 vs reference")
    assert_close(x0s[0][:, :, ::4, ::4], ref1, 1e-3, 1e-4 * sc1, f"{key}: x0_pred vs reference")
    sums = full[key + "_sums"]
    assert abs(xs[0].double().sum().item() - sums[0]) <= 2e-3 * sums[1]
    if sy == 0.0:
        assert resid <= 1e-4, f"{key}: |A x_0 - y| = {resid}"


@pytest.mark.parametrize("kind", ["celeba", "imagenet"])
def test_fullsize_teacher_forced_steps(full, kind):
    """One engine step from a given state, 256x256, both networks, five operators, DDNM and DDNM+: x0_t and xt_next against the
    oracle's arithmetic on the oracle network's eps — no trajectory, hence no chaos: north_star's tolerance applies."""
    from ddnm_b200 import sampler as ES
    _, _, fwd = oracle_net(kind)
    m = engine_net(kind)
    betas_c = SCH.linear_betas()
    abar = SCH.alpha_bar_table(betas_c)
    g = torch.Generator().manual_seed(4711 + len(kind))
    x_orig = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    ops = {n: FS.oracle_op(full, n) for n in ("sr4", "color", "inpaint", "wh", "deblur")}
    eops = {n: engine_operator(n, o) for n, o in ops.items()}
    ys = {n: o.A(x_orig.reshape(1, -1)) for n, o in ops.items()}
    worst = {}
    for (i, j) in ((900, 800), (300, 200), (0, -1)):
        xt = torch.randn(1, 3, 256, 256, generator=g) * (1.0 if i > 0 else 0.3)
        z = torch.randn(1, 3, 256, 256, generator=g)
        at, atn = abar[i + 1], abar[j + 1]
        with torch.no_grad():
            et = fwd(xt, torch.ones(1) * i)[:, :3]
        x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
        sc = max(1.0, x0_t.abs().max().item())
        for name, oop in ops.items():
            y = ys[name]
            resid = oop.A_pinv(oop.A(x0_t.reshape(1, -1)) - y)
            for sy in (0.0, 0.1):
                if sy == 0.0:
                    ref = atn.sqrt() * (x0_t - resid.reshape(x0_t.shape)) + (1 - atn).sqrt() * 0.85 * z + \
                        (1 - atn).sqrt() * ((1 - 0.85 ** 2) ** 0.5) * et
                else:
                    st, a = (1 - atn).sqrt(), atn.sqrt()
                    ref = a * (x0_t - oop.Lambda(resid.clone(), a, sy, st, 0.85).reshape(x0_t.shape)) + \
                        oop.Lambda_noise(z.reshape(1, -1), a, sy, st, 0.85, et.reshape(1, -1)).reshape(x0_t.shape)
                orig_pairs = ES.time_pairs
                ES.time_pairs = lambda *a_, **k_: [(i, j)]
                try:
                    args = (xt.to(dev), m, betas_c.to(dev), 0.85, eops[name], y.to(dev))
                    kw = dict(config=sampler_config(1000, 1, 1), noise=z[None].to(dev))
                    xs, x0s = ES.ddnm_diffusion(*args, **kw) if sy == 0.0 else ES.ddnm_plus_diffusion(*args, sy, **kw)
                finally:
                    ES.time_pairs = orig_pairs
                e0 = float((x0s[0] - x0_t).abs().max()) / sc
                e1 = float((xs[0] - ref).abs().max()) / sc
                worst[f"{name}_s{sy}_{i}"] = (e0, e1)
                assert_close(x0s[0], x0_t, 1e-3, 1e-4 * sc, f"{kind} {name} s{sy} step {i}->{j}: x0_t")
                assert_close(xs[0], ref, 1e-3, 1e-4 * sc, f"{kind} {name} s{sy} step {i}->{j}: xt_next")
    _log(dict(test="fullsize_teacher_forced_steps", net=kind, max_err_over_scale_x0t=max(v[0] for v in worst.values()),
              max_err_over_scale_xt_next=max(v[1] for v in worst.values())))


def test_inpainting_real_mask_bit_exact(full):
    """exp/inp_masks/mask.npy at 256x256 (index construction diffusion.py:466-470, operator svd_operators.py:324-439): pure
    indexing, so A, A_pinv and the projection must equal the oracle's bit for bit."""
    oop = FS.oracle_op(full, "inpaint")
    eop = engine_operator("inpaint", oop)
    assert eop.y_dim == 145314
    g = torch.Generator().manual_seed(5)
    B = 3
    x = torch.rand(B, 3, 256, 256, generator=g) * 2 - 1
    y = oop.A(x.reshape(B, -1))
    ye = eop.A(x.to(dev))
    assert ye.shape == y.shape and torch.equal(ye.cpu(), y), "A (gather) must be bit-exact"
    yq = y * 0.9 + 0.05
    assert torch.equal(eop.A_pinv(yq.to(dev)).cpu(), oop.A_pinv(yq.clone())), "A_pinv (scatter) must be bit-exact"
    assert torch.equal(eop.project(x.to(dev), yq.to(dev)).cpu().reshape(B, -1), oop.project(x, yq).reshape(B, -1)), "projection"
    # Lambda / Lambda_noise touch kept entries only; missing entries pass through unchanged
    v = torch.randn(B, 3 * 256 * 256, generator=g)
    e = torch.randn(B, 3 * 256 * 256, generator=g)
    for (a, sy, st) in LAMBDA_CASES:
        at, stt = torch.tensor(a), torch.tensor(st)
        assert_close(eop.Lambda(v.to(dev), at, sy, stt, 0.85), oop.Lambda(v.clone(), at, sy, stt, 0.85), 1e-5, 1e-6, "inpaint Lambda")
        assert_close(eop.Lambda_noise(v.to(dev), at, sy, stt, 0.85, e.to(dev)), oop.Lambda_noise(v.clone(), at, sy, stt, 0.85, e.clone()),
                     1e-5, 1e-6, "inpaint Lambda_noise")


@pytest.mark.parametrize("dim", [32, 256])
def test_deblur_uni_vs_reference(full, dim):
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
    eop = engine_op("deblur", o, dim)
    sub = (lambda z: z.reshape(B, -1)) if dim == 32 else (lambda z: z.reshape(B, -1)[:, ::61])
    xd, vd, ed = x.to(dev), v.to(dev), e.to(dev)
    y = o.A(x.reshape(B, -1))
    yq = y * 0.9 + 0.05
    assert_close(eop.A(xd), y, 1e-4, 1e-5, "deblur_uni A vs oracle")
    assert_close(sub(eop.A(xd)), full[f"{tag}_A"].reshape(B, -1), 1e-4, 1e-5, "deblur_uni A vs reference")
    assert_close(sub(eop.A_pinv(yq.to(dev))), full[f"{tag}_Apinv"].reshape(B, -1), 1e-4, 2e-5, "deblur_uni A_pinv vs reference")
    assert_close(sub(eop.project(xd, yq.to(dev))), full[f"{tag}_proj"].reshape(B, -1), 1e-4, 2e-5, "deblur_uni project vs reference")
    for ci, (a_, sy, st) in enumerate(LAMBDA_CASES):
        at, stt = torch.tensor(a_), torch.tensor(st)
        assert_close(sub(eop.Lambda(vd, at, sy, stt, 0.85)), full[f"{tag}_L{ci}"].reshape(B, -1), 1e-4, 2e-5, f"deblur_uni Lambda{ci}")
        assert_close(sub(eop.Lambda_noise(vd, at, sy, stt, 0.85, ed)), full[f"{tag}_Ln{ci}"].reshape(B, -1), 1e-4, 2e-5, f"deblur_uni Ln{ci}")


def test_chunked_noise_equals_full_tape():
    """The bounded-memory loop (chunks of pairs drawn on a side stream) consumes torch's generator exactly like one randn_like per
    pair (svd_ddnm.py:65,74): same seed -> same result as the full tape, with and without time travel."""
    from ddnm_b200 import sampler as ES
    from test_gpu_parity import _engine_model
    from oracle import unet_simple as U
    cfg = U.SimpleUNetConfig.tiny()
    m = _engine_model(cfg)
    oop = O.SuperResolution.make(3, 32, 4)
    eop = engine_op("sr4", oop, 32)
    g = torch.Generator().manual_seed(3)
    x_orig = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).to(dev)
    x_T = torch.randn(2, 3, 32, 32, generator=g).to(dev)
    y = eop.A(x_orig)
    betas = SCH.linear_betas().to(dev)
    for (T, tl, tr, sy) in ((7, 1, 1, 0.0), (6, 2, 2, 0.1)):
        conf = sampler_config(T, tl, tr)
        npairs = len(SCH.time_pairs(1000, T, tl, tr))
        torch.manual_seed(99)
        tape = torch.stack([torch.randn_like(x_T) for _ in range(npairs)])
        run = (lambda **k: ES.ddnm_diffusion(x_T, m, betas, 0.85, eop, y, config=conf, **k)) if sy == 0.0 else \
              (lambda **k: ES.ddnm_plus_diffusion(x_T, m, betas, 0.85, eop, y, sy, config=conf, **k))
        ref = run(noise=tape)
        old = ES.NOISE_CHUNK_BYTES
        try:
            for chunk_pairs in (1, 3, 1000):
                ES.NOISE_CHUNK_BYTES = chunk_pairs * x_T.numel() * 4
                torch.manual_seed(99)
                got = run()
                assert torch.equal(got[0][0], ref[0][0]) and torch.equal(got[1][0], ref[1][0]), (T, tl, tr, sy, chunk_pairs)
        finally:
            ES.NOISE_CHUNK_BYTES = old
