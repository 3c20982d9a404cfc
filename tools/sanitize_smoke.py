"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Shapes are tiny but route through the tensor path (both layouts, RW 16 and 32, fused and not)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jittor_dcn_b200 as dcn

torch.manual_seed(0)
cases = [(2, 64, 64, 16, 16, 3, 1, 1), (2, 16, 32, 32, 32, 3, 2, 1), (1, 128, 192, 12, 16, 3, 1, 1),
         (2, 32, 64, 16, 16, 3, 2, 1), (3, 5, 7, 9, 13, 3, 1, 1)]
for (B, C, O, H, W, k, s, p) in cases:
    for variant in (dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR):
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        x = torch.randn(B, C, H, W, device="cuda")
        off = torch.randn(B, 2 * k * k, Ho, Wo, device="cuda") * 3
        wt = torch.randn(O, C, k, k, device="cuda") * 0.1
        bias = torch.randn(O, device="cuda")
        gout = torch.randn(B, O, Ho, Wo, device="cuda")
        for flags in (0, dcn.FLAG_FORCE_SIMT):
            out = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant, flags=flags)
            grads = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, flags=flags)
        torch.cuda.synchronize()
        print("ok", (B, C, O, H, W), "variant", variant, float(out.abs().sum()))
# whole layer (offset conv on the engine: shifted-view kernels for C % 64 == 0, plain mode otherwise) + RoI pools
from jittor_dcn_b200.functional import dcn_layer_backward, dcn_layer_forward
for (B, C, O, H, W, s) in [(2, 64, 64, 16, 16, 1), (2, 16, 32, 32, 32, 2), (1, 128, 64, 12, 20, 1), (2, 64, 128, 16, 16, 2)]:
    for variant in (dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR):
        x = torch.randn(B, C, H, W, device="cuda")
        wt = torch.randn(O, C, 3, 3, device="cuda") * 0.1
        wo = torch.randn(18, C, 3, 3, device="cuda") * 0.02
        bo = torch.randn(18, device="cuda")
        Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
        gout = torch.randn(B, O, Ho, Wo, device="cuda")
        off, out = dcn_layer_forward(x, wo, bo, wt, None, 3, s, 1, variant)
        grads = dcn_layer_backward(x, off, wo, wt, gout, True, False, 3, s, 1, variant)
        torch.cuda.synchronize()
        print("ok layer", (B, C, O, H, W, s), "variant", variant, float(out.abs().sum()))
f = torch.randn(2, 40, 9, 11, device="cuda", requires_grad=True)
rois = torch.tensor([[0, 1., 1., 6., 5.], [1, -3., 2., 20., 30.]], device="cuda")
o = torch.randn(2, 1, 2, device="cuda", requires_grad=True)
dcn.DeformRoIPool(1)(f, rois, o).sum().backward()
torch.cuda.synchronize()
print("sanitize smoke done")
