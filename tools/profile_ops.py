"""Workload for the ncu captures of the HBM-bound kernels: one 2-step DDNM and one 2-step DDNM+ sampling at
BASELINE's size (celeba `Model`, B = 16, 256x256) per operator, so that the fused per-step kernels (local_kernel, inpaint_kernel,
fwht_rows/cols_kernel + wh_spec_kernel, sgemm_kernel + mul_table_kernel in the deblur step) and gn_apply_kernel run on
real-sized data.  Run under ncu with -k regex:<kernel> (see tools/gpu_round2.sh); without ncu it just runs the loops."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddnm_b200 import operators as E                                   # noqa: E402
from ddnm_b200.model import Model                                       # noqa: E402
from ddnm_b200.sampler import sample_device                             # noqa: E402
from ddnm_b200.weights import random_state_dict                         # noqa: E402

ns = types.SimpleNamespace


def main():
    which = sys.argv[1:] or ["sr4", "color", "inpaint", "wh", "deblur"]
    dev = torch.device("cuda", 0)
    B = 16
    mcfg = ns(model=ns(type="simple", ch=128, out_ch=3, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.0,
                       in_channels=3, resamp_with_conv=True), data=ns(image_size=256), diffusion=ns(num_diffusion_timesteps=1000))
    model = Model(mcfg)
    model.load_state_dict(random_state_dict(mcfg, 1234))
    conf = ns(diffusion=ns(num_diffusion_timesteps=1000), time_travel=ns(T_sampling=2, travel_length=1, travel_repeat=1))
    import numpy as np
    betas = torch.from_numpy(np.linspace(1e-4, 2e-2, 1000, dtype="float64")).float().to(dev)
    torch.manual_seed(0)
    x_orig = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
    x_T = torch.randn(B, 3, 256, 256, device=dev)
    for name in which:
        if name == "sr4":
            op = E.SuperResolution(3, 256, 4, dev)
        elif name == "color":
            op = E.Colorization(256, dev)
        elif name == "inpaint":
            m = (torch.rand(256, 256) > 0.25)
            mr = torch.nonzero(~m.reshape(-1)).long().reshape(-1) * 3
            op = E.Inpainting(3, 256, torch.cat([mr, mr + 1, mr + 2]), dev)
        elif name == "wh":
            op = E.WalshHadamardCS(3, 256, 4, torch.randperm(256 ** 2), dev)
        elif name == "deblur":
            sigma = 10
            pdf = lambda z: torch.exp(torch.Tensor([-0.5 * (z / sigma) ** 2]))   # noqa: E731
            k = torch.Tensor([pdf(-2), pdf(-1), pdf(0), pdf(1), pdf(2)])
            op = E.Deblurring(k / k.sum(), 3, 256, dev)
        y = op.A(x_orig)
        for plus, sy in ((False, 0.0), (True, 0.1)):
            sample_device(x_T, model, betas, 0.85, op, y, sy, plus, conf)
        torch.cuda.synchronize()
        print("ran", name)


if __name__ == "__main__":
    main()
