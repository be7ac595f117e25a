"""TEST INFRASTRUCTURE ONLY — a small differentiable stand-in for the runner's classifier guidance.

The reference builds ``cond_fn`` from an ImageNet noisy-image classifier (guided_diffusion/diffusion.py:181-189):
``grad_x log softmax(classifier(x, t))[y] * classifier_scale`` via autograd.  The fixtures and tests use the same recipe on a
fixed random linear "classifier" over 4x4-pooled pixels, so the guided sampling path (class-conditional denoiser, label
override, the ``x`` quirk, the ``et`` update) can be pinned without the 54M-parameter checkpoint.  Works on CPU and CUDA tensors.
"""
import torch
import torch.nn.functional as F


def make_toy_cond_fn(resolution=32, num_classes=1000, scale=1.0, seed=5):
    g = torch.Generator().manual_seed(seed)
    feat = 3 * (resolution // 4) ** 2
    W = torch.randn(num_classes, feat, generator=g) * 0.05

    def cond_fn(x, t, y):
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            f = F.avg_pool2d(x_in, 4).reshape(x_in.shape[0], -1)
            logits = (f @ W.t().to(x_in.device)) * (1.0 + t.to(x_in.device).float()[:, None] / 1000.0)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return torch.autograd.grad(selected.sum(), x_in)[0] * scale
    return cond_fn
