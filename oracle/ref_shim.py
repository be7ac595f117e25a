"""Run the UNMODIFIED reference samplers on a machine without their hard-coded CUDA device: ``.to('cuda')`` /
``torch.ones(..., device=cuda)`` (svd_ddnm.py:45,49,72,106,110,138) are redirected to the CPU and, optionally,
``torch.randn_like`` is fed from a noise tape so that the reference, the oracle and the CUDA engine consume identical draws.
Nothing in the reference's code is edited: the redirection is a context manager around the call.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch


class cpu_shim:
    """Redirect the reference's hard-coded device moves (svd_ddnm.py:45,72) and feed randn_like from a tape (None = real draws)."""

    def __init__(self, tape=None):
        self.tape = None if tape is None else list(tape)

    def __enter__(self):
        self._to, self._rl, self._ones = torch.Tensor.to, torch.randn_like, torch.ones
        orig_to, orig_ones = self._to, self._ones

        def ones(*a, **k):          # svd_ddnm.py:49,110 build the label vector with device=torch.device("cuda")
            k.pop("device", None)
            return orig_ones(*a, **k)
        torch.ones = ones

        def to(t, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                return t
            return orig_to(t, *a, **k)
        torch.Tensor.to = to
        if self.tape is not None:
            tape = self.tape

            def randn_like(x, *a, **k):
                z = tape.pop(0)
                assert z.shape == x.shape
                return z
            torch.randn_like = randn_like
        return self

    def __exit__(self, *exc):
        torch.Tensor.to, torch.randn_like, torch.ones = self._to, self._rl, self._ones
