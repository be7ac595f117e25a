"""Time schedule and alpha-bar table of the DDNM sampler (oracle restatement).

Follows /root/reference/functions/svd_ddnm.py:167-206 (get_schedule_jump,
_check_times), :10-13 (compute_alpha) and guided_diffusion/diffusion.py:46-76,
92-98 (linear beta schedule, float64 linspace cast to float32).
"""
import numpy as np
import torch


def linear_betas(beta_start=1e-4, beta_end=2e-2, n=1000):
    # diffusion.py:61-63 (np.linspace in float64) then :98 (.float())
    return torch.from_numpy(np.linspace(beta_start, beta_end, n, dtype=np.float64)).float()


def alpha_bar_table(betas):
    """abar[k] = prod_{s<k}(1-beta_s) with abar[0] = 1, so that alpha-bar of time t is abar[t+1]
    and t = -1 gives exactly 1 (svd_ddnm.py:10-13)."""
    b = torch.cat([torch.zeros(1, dtype=betas.dtype), betas.cpu()], dim=0)
    return (1 - b).cumprod(dim=0)


def jump_schedule(t_sampling, travel_length, travel_repeat):
    """Sequence of sampling-time indices T-1 ... 0, -1 with RePaint-style jumps back
    (svd_ddnm.py:167-190).  Consecutive entries always differ by exactly 1."""
    budget = {}
    for j in range(0, t_sampling - travel_length, travel_length):
        budget[j] = travel_repeat - 1
    seq = []
    t = t_sampling
    while t >= 1:
        t -= 1
        seq.append(t)
        if budget.get(t, 0) > 0:
            budget[t] -= 1
            for _ in range(travel_length):
                t += 1
                seq.append(t)
    seq.append(-1)
    # the reference's sanity checks (svd_ddnm.py:192-206)
    assert seq[0] > seq[1] and seq[-1] == -1
    assert all(abs(a - b) == 1 for a, b in zip(seq[:-1], seq[1:]))
    assert all(-1 <= t <= t_sampling for t in seq)
    return seq


def time_pairs(num_timesteps, t_sampling, travel_length, travel_repeat):
    """(i, j) pairs in model-time units; j < i is a denoise step, j > i a travel-back step
    (svd_ddnm.py:23-38)."""
    skip = num_timesteps // t_sampling
    seq = jump_schedule(t_sampling, travel_length, travel_repeat)
    out = []
    for a, b in zip(seq[:-1], seq[1:]):
        i, j = a * skip, b * skip
        if j < 0:
            j = -1
        out.append((i, j))
    return out
