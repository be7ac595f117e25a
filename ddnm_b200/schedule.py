"""Host-side time schedule of the sampler (functions/svd_ddnm.py:167-206 get_schedule_jump/_check_times, :10-13
compute_alpha).  Pure integers and one fp32 cumprod; produced once per run and handed to the CUDA loop."""
import torch


def get_schedule_jump(T_sampling, travel_length, travel_repeat):
    jumps = {j: travel_repeat - 1 for j in range(0, T_sampling - travel_length, travel_length)}
    t, ts = T_sampling, []
    while t >= 1:
        t -= 1
        ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] -= 1
            for _ in range(travel_length):
                t += 1
                ts.append(t)
    ts.append(-1)
    _check_times(ts, -1, T_sampling)
    return ts


def _check_times(times, t_0, T_sampling):
    assert times[0] > times[1], (times[0], times[1])
    assert times[-1] == -1, times[-1]
    for t_last, t_cur in zip(times[:-1], times[1:]):
        assert abs(t_last - t_cur) == 1, (t_last, t_cur)
    for t in times:
        assert t >= t_0, (t, t_0)
        assert t <= T_sampling, (t, T_sampling)


def time_pairs(num_timesteps, T_sampling, travel_length, travel_repeat):
    skip = num_timesteps // T_sampling
    times = get_schedule_jump(T_sampling, travel_length, travel_repeat)
    pairs = []
    for i, j in zip(times[:-1], times[1:]):
        i, j = i * skip, j * skip
        if j < 0:
            j = -1
        pairs.append((i, j))
    return pairs


def alpha_bar_table(b):
    """cumprod table with the reference's ops: abar[t + 1] == compute_alpha(b, t); abar[0] == 1 serves t = -1."""
    beta = torch.cat([torch.zeros(1).to(b.device), b], dim=0)
    return (1 - beta).cumprod(dim=0).float().cpu()
