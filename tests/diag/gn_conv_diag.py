"""Diagnostic (GPU): the fused GroupNorm convolution against fp64 for both descriptor modes, plus timing vs the unfused pair."""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ddnm_b200 import _lib   # noqa: E402

dev = "cuda"


def run(N, H, W, Cin, Cout, side_c=0, res=False, norm=True, silu=True, iters=0, seed=0):
    L = _lib.lib()
    torch.manual_seed(seed)
    x = torch.randn(N, Cin, H, W, device=dev) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev)
    g, be = torch.randn(Cin, device=dev), torch.randn(Cin, device=dev)
    side = torch.randn(N, side_c, H, W, device=dev) if side_c else None
    sw = torch.randn(Cout, side_c, 1, 1, device=dev) / side_c ** 0.5 if side_c else None
    r = torch.randn(N, Cout, H, W, device=dev) if res else None
    nhwc = lambda t: None if t is None else t.permute(0, 2, 3, 1).contiguous()   # noqa: E731
    out = torch.empty(N, H, W, Cout, device=dev)
    ms = C.c_float(0)
    xs, ss, rs = nhwc(x), nhwc(side), nhwc(r)
    _lib.check(L.ddnm_conv_gn_tc(_lib.ptr(xs), N, H, W, Cin, 32, _lib.ptr(g) if norm else None, _lib.ptr(be) if norm else None, 1e-6,
                                 int(silu), _lib.ptr(w), _lib.ptr(b), Cout, _lib.ptr(ss), side_c,
                                 _lib.ptr(sw.contiguous()) if sw is not None else None, _lib.ptr(rs), _lib.ptr(out), iters, C.byref(ms), None))
    torch.cuda.synchronize()
    xd = x.double().cpu()
    h = F.group_norm(xd, 32, g.double().cpu(), be.double().cpu(), 1e-6) if norm else xd
    if silu and norm:
        h = h * torch.sigmoid(h)
    ref = F.conv2d(h, w.double().cpu(), b.double().cpu(), padding=1)
    if side is not None:
        ref = ref + F.conv2d(side.double().cpu(), sw.double().cpu())
    if r is not None:
        ref = ref + r.double().cpu()
    got = out.permute(0, 3, 1, 2).double().cpu()
    err = (got - ref).abs().max().item()
    return err, ref.abs().max().item(), ms.value


if __name__ == "__main__" and len(sys.argv) == 1:
    L = _lib.lib()
    for mode in (0, 1):
        L.ddnm_tc_debug_gn_desc_mode(mode)
        for shape in [(2, 128, 128, 128, 128), (1, 256, 256, 64, 128), (2, 128, 128, 128, 256)]:
            try:
                err, sc, _ = run(*shape)
                print(f"desc_mode {mode} shape {shape}: max err {err:.3e} (ref absmax {sc:.2f})", flush=True)
            except Exception as e:
                print(f"desc_mode {mode} shape {shape}: FAILED {str(e)[:200]}", flush=True)
                sys.exit(1)


def timing():
    """fused kernel vs (gn_apply + conv_tc) on the celeba / imagenet wide-layer shapes, random data"""
    L = _lib.lib()
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    pf = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    L.ddnm_tc_debug_gn_pf_dist(pf)
    print(f"--- L2 prefetch distance {pf}")
    for (N, H, W, Cin, Cout, side) in [(16, 256, 256, 256, 128, 0), (16, 256, 256, 128, 128, 256), (16, 256, 256, 128, 128, 0),
                                       (16, 128, 128, 256, 128, 0), (8, 256, 256, 256, 256, 0), (8, 128, 128, 512, 256, 0)]:
        err, sc, ms = run(N, H, W, Cin, Cout, side_c=side, iters=5)
        ms_c, fl = C.c_float(0), C.c_double(0)
        _lib.check(L.ddnm_conv_tc_bench(N, H, W, Cin, Cout, 0, -5, C.byref(ms_c), C.byref(fl)))
        flops = 2.0 * N * H * W * Cout * (9 * Cin + side)
        # one more launch with the in-kernel clock counters on
        dbg.zero_()
        L.ddnm_tc_debug_gn_counters(_lib.ptr(dbg))
        run(N, H, W, Cin, Cout, side_c=side, iters=0)
        L.ddnm_tc_debug_gn_counters(None)
        d = dbg.reshape(148, 16).double().cpu()
        tr = d[d[:, 0] > 0]
        mm = d[d[:, 10] > 0]
        per = lambda k: (tr[:, k] / tr[:, 0]).mean().item()     # noqa: E731
        print(f"   transform group 0 (warp 0), clocks per unit of the group: wait-free-slot {per(1):.0f}, wait-loads {per(2):.0f}, convert+store {per(3):.0f}, "
              f"fence+arrive {per(4):.0f} (units per CTA {tr[:, 0].mean().item():.0f}); UMMA issuer: total {mm[:, 10].mean().item():.3g} clk, "
              f"waiting for A {100 * (mm[:, 11] / mm[:, 10]).mean().item():.0f}%, B {100 * (mm[:, 12] / mm[:, 10]).mean().item():.0f}%, "
              f"accumulator {100 * (mm[:, 13] / mm[:, 10]).mean().item():.0f}%", flush=True)
        print(f"timing N{N} {H}x{W} {Cin}->{Cout} side {side}: fused {ms:.3f} ms = {flops / ms / 1e9:.0f} TF/s (err {err:.1e}); "
              f"unfused conv alone (no side) {ms_c.value:.3f} ms = {fl.value / ms_c.value / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "timing":
    timing()
