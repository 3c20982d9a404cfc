"""Side measurement (not the headline bench): standard DCNv1 mode (DCN_VARIANT_DCNV1) against
torchvision.ops.deform_conv2d's own CUDA kernels on the same B200, fwd+bwd, CUDA-event timed.
    python tools/bench_vs_torchvision.py [B C O H W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchvision.ops import deform_conv2d as tv_dcn
import jittor_dcn_b200 as dcn

B, C, O, H, W = (int(a) for a in sys.argv[1:6]) if len(sys.argv) >= 6 else (256, 64, 64, 128, 128)
torch.manual_seed(0)
x = torch.randn(B, C, H, W, device="cuda")
off = torch.randn(B, 18, H, W, device="cuda") * 2
wt = torch.randn(O, C, 3, 3, device="cuda") * (2.0 / (C * 9)) ** 0.5
bias = torch.randn(O, device="cuda")
gout = torch.randn(B, O, H, W, device="cuda")


def ours():
    out = dcn.dcn_forward(x, off, wt, bias, 3, 1, 1, dcn.VARIANT_DCNV1)
    return dcn.dcn_backward(x, off, wt, gout, True, 3, 1, 1, dcn.VARIANT_DCNV1)


def theirs():
    xi, oi, wi, bi = (t.detach().requires_grad_(True) for t in (x, off, wt, bias))
    out = tv_dcn(xi, oi, wi, bi, stride=1, padding=1)
    return torch.autograd.grad(out, [xi, oi, wi, bi], gout)


def time_ms(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_ours, t_tv = time_ms(ours), time_ms(theirs)
go, gt = ours(), theirs()
err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(go, gt))
print(f"shape B={B} C={C} O={O} {H}x{W}: ours {t_ours:.2f} ms ({B / t_ours * 1e3:.0f} img/s), "
      f"torchvision CUDA {t_tv:.2f} ms ({B / t_tv * 1e3:.0f} img/s), speed-up {t_tv / t_ours:.2f}x, "
      f"max rel grad diff {err:.2e}")
