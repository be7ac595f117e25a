"""Workload for ncu captures of single launches of the UNet forward: the celeba `Model` at B = 16, run EAGERLY (no CUDA graph) twice —
the first forward warms up, the second is the one to capture (`ncu -k regex:conv_tc_kernel -s <112 + index> -c 1 ...`; the index of
a layer among the 112 tensor-core launches of a forward comes from `Model.profile`)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddnm_b200.model import Model                                       # noqa: E402
from ddnm_b200.weights import random_state_dict                         # noqa: E402

ns = types.SimpleNamespace


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda", 0)
    mcfg = ns(model=ns(type="simple", ch=128, out_ch=3, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.0,
                       in_channels=3, resamp_with_conv=True), data=ns(image_size=256), diffusion=ns(num_diffusion_timesteps=1000))
    model = Model(mcfg)
    model.use_cuda_graph = False
    model.load_state_dict(random_state_dict(mcfg, 1234))
    torch.manual_seed(0)
    x = torch.randn(B, 3, 256, 256, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    for _ in range(2):
        model(x, t)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
