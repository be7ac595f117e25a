"""hq_demo's arbitrary-size DDNM restoration (mask-shift trick): CPU — the oracle against the results of
hq_demo's own p_sample_loop stored in tests/golden/hq.npz; GPU — ddnm_b200.hq.restore against both."""
import numpy as np
import pytest
import torch

from oracle import hq as HQO
from oracle import unet_openai as UO

from helpers import assert_close

JUMP = dict(t_T=6, n_sample=1, jump_length=2, jump_n_sample=2)
CASES = [("w384", (64, 96), 0.0), ("w320_noisy", (64, 80), 0.1), ("h320", (80, 64), 0.0)]


def hq_cfg():
    return UO.OpenAIUNetConfig(image_size=256, model_channels=64, num_res_blocks=1, channel_mult=(1, 1, 2, 2, 4, 4),
                               attention_resolutions=(32, 16, 8), num_head_channels=64, out_channels=6, num_classes=1000)


def hq_inputs(g, key, hw):
    h, w = hw
    gen = torch.Generator().manual_seed(int(g[key + "_seed"][0]))
    y_img = torch.rand(1, 3, h, w, generator=gen) * 2 - 1
    assert np.array_equal(y_img.numpy(), g[key + "_y"])
    n = HQO.count_draws(4 * h, 4 * w, JUMP)
    return y_img, [torch.randn(1, 3, 256, 256, generator=gen) for _ in range(n)]


def test_hq_schedule_and_windows():
    ts = HQO.get_schedule_jump(**JUMP)
    assert ts[0] == 5 and ts[-1] == -1 and all(abs(a - b) == 1 for a, b in zip(ts[:-1], ts[1:]))
    assert HQO.count_draws(256, 384, JUMP) == 1 + 2 * (len(ts) - 1)
    K = HQO.SpacedConstants(1000, 6)
    assert K.timestep_map[0] == 0 and K.timestep_map[-1] == 999 and len(K.timestep_map) == 6
    assert HQO.window_origin(0, 1, 1, 2, 256, 320) == (0, 64) and HQO.window_origin(1, 0, 2, 1, 320, 256) == (64, 0)


@pytest.mark.parametrize("case", CASES[:2], ids=lambda c: c[0])
def test_oracle_hq_matches_reference(gold, case):
    key, hw, sy = case
    g = gold["hq"]
    cfg = hq_cfg()
    sd = UO.init_state_dict(cfg, 1234)
    y_img, tape = hq_inputs(g, key, hw)
    with torch.no_grad():
        out = HQO.restore(lambda a, b, c: UO.forward(sd, a, b.float(), cfg, y=c), y_img, torch.tensor([950]), tape, scale=4, sigma_y=sy,
                          resize_y=True, respacing=6, jump=JUMP)
    assert out.shape == (1, 3, 4 * hw[0], 4 * hw[1])
    assert np.abs(out[:, :, ::2, ::2].numpy() - g[key + "_out_s2"]).max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_engine_hq_restore_vs_reference(gold, case):
    from ddnm_b200 import hq as HQ
    from ddnm_b200.model import create_model
    key, hw, sy = case
    g = gold["hq"]
    cfg = hq_cfg()
    m = create_model(image_size=256, num_channels=64, num_res_blocks=1, learn_sigma=True, class_cond=True, attention_resolutions="32,16,8",
                     num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True, use_fp16=False)
    m.load_state_dict(UO.init_state_dict(cfg, 1234))
    y_img, tape = hq_inputs(g, key, hw)
    out = HQ.restore(m, y_img.cuda(), torch.tensor([950]), deg="sr_averagepooling", scale=4, sigma_y=sy, resize_y=True,
                     timestep_respacing=6, schedule_jump_params=JUMP, noise=torch.stack(tape).cuda())
    assert out.shape == (1, 3, 4 * hw[0], 4 * hw[1]) and not out.is_cuda
    ref = g[key + "_out_s2"]
    assert_close(out[:, :, ::2, ::2], ref, 1e-3, 1e-4 * max(1.0, float(np.abs(ref).max())), f"hq {key} vs hq_demo")
    sums = g[key + "_sums"]
    assert abs(out.double().sum().item() - sums[0]) <= 1e-3 * sums[1]
