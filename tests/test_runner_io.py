"""Row f2/f3 of SURVEY §8: GeneralA and the runner's I/O step (data transforms, save_image bytes, PSNR).
CPU part: the oracle against the reference's golden outputs + the host PNG container.  GPU part: the CUDA path through the C ABI."""
import types

import numpy as np
import pytest
import torch

from oracle import operators as O
from oracle import runner_io as RIO

from helpers import assert_close

CASES = dict(rescaled=(True, False, False, False), logit=(False, True, False, False), deq=(True, False, True, True),
             plain=(False, False, False, False))


def _cfg(resc, logit, udq=False, gdq=False, channels=3, size=16):
    ns = types.SimpleNamespace
    return ns(data=ns(rescaled=resc, logit_transform=logit, uniform_dequantization=udq, gaussian_dequantization=gdq,
                      channels=channels, image_size=size))


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_general_a_matches_reference(gold):
    g = gold["general_a"]
    o = O.GeneralA(torch.from_numpy(g["U"]), torch.from_numpy(g["S"]), torch.from_numpy(g["V"]))
    x, yq = torch.from_numpy(g["x"]), torch.from_numpy(g["yq"])
    assert int((torch.from_numpy(g["S"]) == 0).sum()) == 2          # the ZERO = 1e-3 threshold branch (svd_operators.py:184-185)
    assert_close(o.A(x), g["y"], 1e-5, 2e-6, "GeneralA.A")
    assert_close(o.A_pinv(yq.clone()), g["pinv"], 1e-5, 2e-6, "GeneralA.A_pinv")
    assert_close(o.project(x, yq), g["proj"], 1e-5, 4e-6, "GeneralA project")
    with pytest.raises(NotImplementedError):
        o.Lambda(x, 0.5, 0.1, 0.1, 0.85)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_runner_io_matches_reference(gold, name):
    g = gold["runner_io"]
    resc, logit, udq, gdq = CASES[name]
    X, xm = torch.from_numpy(g["X"]), torch.from_numpy(g["xm"])
    un = torch.from_numpy(g[f"{name}_un"]) if udq else None
    gn = torch.from_numpy(g[f"{name}_gn"]) if gdq else None
    T = RIO.data_transform(X, resc, logit, un, gn)
    assert_close(T, g[f"{name}_T"], 1e-6, 1e-6, "data_transform")
    inv = RIO.inverse_data_transform(xm, resc, logit)
    assert_close(inv, g[f"{name}_inv"], 1e-6, 1e-6, "inverse_data_transform")
    assert np.array_equal(RIO.to_uint8_hwc(inv).numpy(), g[f"{name}_u8"])
    orig = RIO.inverse_data_transform(T, resc, logit)
    ps = torch.stack([RIO.psnr(inv[j], orig[j]) for j in range(inv.shape[0])])
    assert_close(ps, g[f"{name}_psnr"], 1e-5, 1e-5, "psnr")


def test_png_container_roundtrip_and_pil_agrees(gold, tmp_path):
    from ddnm_b200.runner import decode_png, encode_png, save_png
    img = gold["runner_io"]["rescaled_u8"][0]
    blob = encode_png(img)
    assert np.array_equal(decode_png(blob), img)
    gray = np.ascontiguousarray(img[:, :, :1])
    assert np.array_equal(decode_png(encode_png(gray)), gray)
    save_png(tmp_path / "a.png", img)
    try:
        from PIL import Image
    except ImportError:
        return
    assert np.array_equal(np.array(Image.open(tmp_path / "a.png").convert("RGB")), img)


# ------------------------------------------------------------------------------------------------ GPU: CUDA path vs golden
@pytest.mark.gpu
def test_general_a_engine_vs_reference_golden(gold):
    from ddnm_b200.operators import GeneralA
    g = gold["general_a"]
    arts = tuple(torch.from_numpy(g[k]) for k in ("U", "S", "V"))
    e = GeneralA(None, artefacts=arts)
    assert e.y_dim == 48 and e.x_dim == 192
    x, yq = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["yq"]).cuda()
    assert_close(e.A(x), g["y"], 1e-4, 1e-5, "GeneralA.A")
    assert_close(e.A_pinv(yq), g["pinv"], 1e-4, 1e-5, "GeneralA.A_pinv")
    assert_close(e.project(x, yq), g["proj"], 1e-4, 2e-5, "GeneralA project")
    with pytest.raises(NotImplementedError):
        e.Lambda(x, 0.5, 0.1, 0.1, 0.85)
    # constructor path: torch.svd on the matrix itself (sign / basis of the null space may differ between LAPACK builds, A and
    # A_pinv do not depend on it)
    e2 = GeneralA(torch.from_numpy(g["A"]))
    assert_close(e2.A(x), g["y"], 1e-4, 1e-5, "GeneralA(A).A")
    assert_close(e2.A_pinv(yq), g["pinv"], 1e-3, 1e-4, "GeneralA(A).A_pinv")


@pytest.mark.gpu
def test_sampler_with_general_a_vs_oracle(gold):
    """DDNM through a dense GeneralA operator on the tiny net (n = 3*32*32 = 3072, m = 256) against the oracle loop."""
    from oracle import sampler as S, schedule as SCH, unet_simple as U
    from ddnm_b200.model import Model
    from ddnm_b200.operators import GeneralA
    from ddnm_b200.sampler import ddnm_diffusion, ddnm_plus_diffusion
    from helpers import model_config, sampler_config
    cfg = U.SimpleUNetConfig.tiny()
    sd = U.init_state_dict(cfg, 1234)
    n, m = 3 * cfg.resolution ** 2, 256
    rng = torch.Generator().manual_seed(11)
    A = torch.randn(m, n, generator=rng) / n ** 0.5
    Um, Sm, Vm = torch.svd(A, some=False)
    oop = O.GeneralA(Um, Sm, Vm)
    eop = GeneralA(None, artefacts=(Um, Sm, Vm))
    T = 6
    npairs = len(SCH.time_pairs(1000, T, 1, 1))
    x_T = torch.randn(2, 3, cfg.resolution, cfg.resolution, generator=rng)
    y = oop.A((torch.rand(2, n, generator=rng) * 2 - 1))
    tape = [torch.randn(2, 3, cfg.resolution, cfg.resolution, generator=rng) for _ in range(npairs)]
    betas = SCH.linear_betas()
    with torch.no_grad():
        ox, _ = S.ddnm_sample(x_T, lambda a, b: U.forward(sd, a, b, cfg), betas, 0.85, oop, y, tape, t_sampling=T, travel_length=1,
                              travel_repeat=1, sigma_y=0.0)
    mdl = Model(model_config(cfg))
    mdl.load_state_dict(sd)
    xs, _ = ddnm_diffusion(x_T.cuda(), mdl, betas.cuda(), 0.85, eop, y.cuda(), config=sampler_config(T, 1, 1), noise=torch.stack(tape).cuda())
    assert_close(xs[0], ox, 1e-3, 3e-3, "GeneralA sampler vs oracle")
    assert_close(eop.A(xs[0].cuda()), y, 1e-3, 1e-3, "A x0 = y")
    with pytest.raises(NotImplementedError):       # no Lambda: the reference fails at its first DDNM+ step (svd_operators.py:93-97)
        ddnm_plus_diffusion(x_T.cuda(), mdl, betas.cuda(), 0.85, eop, y.cuda(), 0.1, config=sampler_config(T, 1, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_runner_io_kernels_vs_reference_golden(gold, name):
    from ddnm_b200 import runner as R
    g = gold["runner_io"]
    resc, logit, udq, gdq = CASES[name]
    cfg = _cfg(resc, logit, udq, gdq)
    X, xm = torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["xm"]).cuda()
    un = torch.from_numpy(g[f"{name}_un"]).cuda() if udq else None
    gn = torch.from_numpy(g[f"{name}_gn"]).cuda() if gdq else None
    Xc = X.clone()
    T = R.data_transform(cfg, X, un, gn)
    assert torch.equal(X, Xc), "data_transform must not mutate its input"
    assert_close(T, g[f"{name}_T"], 1e-5, 2e-6, "data_transform")
    if not logit:
        assert np.array_equal(T.cpu().numpy(), g[f"{name}_T"]), "affine transforms are bit-exact"
    assert_close(R.inverse_data_transform(cfg, xm), g[f"{name}_inv"], 1e-6, 1e-6, "inverse_data_transform")
    u8, ps, x01 = R.finish_images(cfg, xm, torch.from_numpy(g[f"{name}_T"]).cuda(), want_float=True)
    assert_close(x01, g[f"{name}_inv"], 1e-6, 1e-6, "finish_images float")
    got, want = u8.cpu().numpy(), g[f"{name}_u8"]
    if logit:   # expf differs from torch.sigmoid by an ulp: allow a handful of +-1 quantisation flips
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got != want).mean() < 2e-3
    else:
        assert np.array_equal(got, want), "uint8 bytes must equal what tvu.save_image encodes"
    assert_close(ps, g[f"{name}_psnr"], 1e-5, 1e-5, "psnr")
    u8b, psb, _ = R.finish_images(cfg, xm)                   # preview mode: no ground truth
    assert psb is None and torch.equal(u8b, u8)


@pytest.mark.gpu
def test_finish_images_full_size_properties():
    """256x256 batch: PSNR of an image against itself shifted by a constant, and against the closed form."""
    from ddnm_b200 import runner as R
    cfg = _cfg(True, False, size=256)
    rng = torch.Generator().manual_seed(3)
    x = (torch.rand(5, 3, 256, 256, generator=rng) * 1.6 - 0.8).cuda()
    d = 0.05
    u8, ps, x01 = R.finish_images(cfg, x + d, x, want_float=True)            # (x+d+1)/2 - (x+1)/2 = d/2 everywhere, no clamping
    want = 10 * np.log10(1 / (d / 2) ** 2)
    assert np.allclose(ps.cpu().numpy(), want, rtol=2e-5), (ps, want)
    ref = ((x + d + 1) / 2).clamp(0, 1).mul(255).add(0.5).clamp(0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    assert torch.equal(u8, ref)
    assert torch.equal(x01, ((x + d + 1) / 2).clamp(0, 1))


@pytest.mark.gpu
def test_restore_batch_matches_oracle_composition(gold, tmp_path):
    """The whole per-batch body of the runner (diffusion.py:533-603) on the tiny net, sr4 + noise, against the same steps composed
    from the oracle pieces; PNG files decode to the reported bytes."""
    from oracle import sampler as S, schedule as SCH, unet_simple as U
    from ddnm_b200 import runner as R
    from ddnm_b200.model import Model
    from helpers import engine_op, model_config, oracle_ops, sampler_config
    cfg = U.SimpleUNetConfig.tiny()
    sd = U.init_state_dict(cfg, 1234)
    res = cfg.resolution
    conf = sampler_config(5, 1, 1)
    conf.data = _cfg(True, False, channels=3, size=res).data
    oop = oracle_ops(gold["operators"], 32)["sr4"]
    eop = engine_op("sr4", oop, 32)
    mdl = Model(model_config(cfg))
    mdl.load_state_dict(sd)
    rng = torch.Generator().manual_seed(8)
    x01 = torch.rand(2, 3, res, res, generator=rng)
    x_T = torch.randn(2, 3, res, res, generator=rng)
    npairs = len(SCH.time_pairs(1000, 5, 1, 1))
    tape = [torch.randn(2, 3, res, res, generator=rng) for _ in range(npairs)]
    betas = SCH.linear_betas()
    out = R.restore_batch(conf, mdl, eop, "sr_averagepooling", x01, betas.cuda(), 0.85, sigma_y=0.0, image_folder=str(tmp_path),
                          idx_so_far=7, x_T=x_T.cuda(), noise=torch.stack(tape).cuda())
    # oracle composition
    xo = RIO.data_transform(x01, True, False)
    y = oop.A(xo.reshape(2, -1))
    with torch.no_grad():
        ox, _ = S.ddnm_sample(x_T, lambda a, b: U.forward(sd, a, b, cfg), betas, 0.85, oop, y, tape, t_sampling=5, travel_length=1,
                              travel_repeat=1, sigma_y=0.0)
    inv, orig = RIO.inverse_data_transform(ox, True, False), RIO.inverse_data_transform(xo, True, False)
    ps = torch.stack([RIO.psnr(inv[j], orig[j]) for j in range(2)])
    assert_close(out["y"], y, 1e-5, 1e-6, "y = A(data_transform(x))")
    assert_close(out["psnr"], ps, 2e-3, 2e-3, "psnr")
    want = RIO.to_uint8_hwc(inv).numpy()
    assert np.abs(out["images"].astype(int) - want.astype(int)).max() <= 1        # sampler tolerance -> at most one grey level
    assert np.array_equal(out["orig"], RIO.to_uint8_hwc(orig).numpy())
    apy = RIO.inverse_data_transform(oop.A_pinv(y).reshape(2, 3, res, res), True, False)
    assert np.abs(out["Apy"].astype(int) - RIO.to_uint8_hwc(apy).numpy().astype(int)).max() <= 1
    for i in range(2):
        for pat, key in (("{}_0.png", "images"), ("Apy/Apy_{}.png", "Apy"), ("Apy/orig_{}.png", "orig")):
            blob = (tmp_path / pat.format(7 + i)).read_bytes()
            assert np.array_equal(R.decode_png(blob), out[key][i])


@pytest.mark.gpu
@pytest.mark.parametrize("deg,opname,sigma_y", [("colorization", "color", 0.0), ("deblur_gauss", "deblur", 0.0), ("inpainting", "inpaint", 0.1)])
def test_restore_batch_preview_branches_and_noisy_path(gold, deg, opname, sigma_y):
    """The runner's per-degradation preview rules (diffusion.py:558-564: deblur shows y, colorization the repeated gray image,
    inpainting adds A^+A(1) - 1) and its noise / DDNM+ switch (:550-551, :587-590), against the oracle composition.  The noise
    added to y is drawn on the device, so the oracle replays the engine's own y."""
    from oracle import sampler as S, schedule as SCH, unet_simple as U
    from ddnm_b200 import runner as R
    from ddnm_b200.model import Model
    from helpers import engine_op, model_config, oracle_ops, sampler_config
    cfg = U.SimpleUNetConfig.tiny()
    sd = U.init_state_dict(cfg, 1234)
    res = cfg.resolution
    T = 4
    conf = sampler_config(T, 1, 1)
    conf.data = _cfg(True, False, channels=3, size=res).data
    oop = oracle_ops(gold["operators"], 32)[opname]
    eop = engine_op(opname, oop, 32)
    mdl = Model(model_config(cfg))
    mdl.load_state_dict(sd)
    rng = torch.Generator().manual_seed(21)
    x01 = torch.rand(2, 3, res, res, generator=rng)
    x_T = torch.randn(2, 3, res, res, generator=rng)
    npairs = len(SCH.time_pairs(1000, T, 1, 1))
    tape = [torch.randn(2, 3, res, res, generator=rng) for _ in range(npairs)]
    betas = SCH.linear_betas()
    out = R.restore_batch(conf, mdl, eop, deg, x01, betas.cuda(), 0.85, sigma_y=sigma_y, add_noise=sigma_y > 0, x_T=x_T.cuda(),
                          noise=torch.stack(tape).cuda())
    xo = RIO.data_transform(x01, True, False)
    y_clean = oop.A(xo.reshape(2, -1))
    y = out["y"].cpu()
    if sigma_y == 0:
        assert_close(y, y_clean, 1e-5, 2e-6, "y")
    else:   # y = A x + sigma_y * N(0, 1): right scale, and not the clean signal
        d = (y - y_clean) / sigma_y
        assert 0.8 < d.std().item() < 1.2 and abs(d.mean().item()) < 0.1
    with torch.no_grad():
        ox, _ = S.ddnm_sample(x_T, lambda a, b: U.forward(sd, a, b, cfg), betas, 0.85, oop, y, tape, t_sampling=T, travel_length=1,
                              travel_repeat=1, sigma_y=sigma_y)
    inv, orig = RIO.inverse_data_transform(ox, True, False), RIO.inverse_data_transform(xo, True, False)
    ps = torch.stack([RIO.psnr(inv[j], orig[j]) for j in range(2)])
    assert_close(out["psnr"], ps, 5e-3, 5e-3, "psnr")
    assert np.abs(out["images"].astype(int) - RIO.to_uint8_hwc(inv).numpy().astype(int)).max() <= 1
    # preview image rule of the degradation
    if deg[:6] == "deblur":
        apy = y.view(2, 3, res, res)
    elif deg == "colorization":
        apy = y.view(2, 1, res, res).repeat(1, 3, 1, 1)
    else:
        apy = oop.A_pinv(y).view(2, 3, res, res)
        apy = apy + oop.A_pinv(oop.A(torch.ones_like(apy).reshape(2, -1))).reshape(apy.shape) - 1
    want = RIO.to_uint8_hwc(RIO.inverse_data_transform(apy, True, False)).numpy()
    assert np.abs(out["Apy"].astype(int) - want.astype(int)).max() <= 1
