"""Debug: whole-layer backward on the engine vs engine span + framework offset-conv backward on the SAME offsets."""
import sys
import torch
import torch.nn.functional as F
import jittor_dcn_b200 as dcn
from jittor_dcn_b200.functional import dcn_layer_forward, dcn_layer_backward, layer_workspace

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [(4, 64, 64, 128, 128, 1), (8, 16, 32, 128, 128, 2), (8, 128, 128, 56, 56, 1), (2, 64, 64, 128, 128, 1),
          (1, 64, 64, 128, 128, 1), (1, 64, 64, 64, 64, 1), (4, 64, 64, 64, 64, 1)]
for variant in (dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR):
    for (B, C, O, H, W, s) in shapes:
        torch.manual_seed(3)
        k, p = 3, 1
        Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
        x = torch.randn(B, C, H, W, device="cuda")
        wt = torch.randn(O, C, 3, 3, device="cuda") * (2.0 / (C * 9)) ** 0.5
        bias = torch.randn(O, device="cuda") * 0.1
        woff = torch.randn(18, C, 3, 3, device="cuda") * 0.01
        boff = torch.randn(18, device="cuda")
        gout = torch.randn(B, O, Ho, Wo, device="cuda")
        for rep in range(2):
            off, out = dcn_layer_forward(x, woff, boff, wt, bias, k, s, p, variant)
            gx, gwoff, gboff, gw, gb = dcn_layer_backward(x, off, woff, wt, gout, True, True, k, s, p, variant)
            # yardstick: engine span + torch autograd of the offset conv, same offsets
            gx2, goff2, gw2, gb2 = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant)
            xi = x.clone().requires_grad_(True)
            wo = woff.clone().requires_grad_(True)
            bo = boff.clone().requires_grad_(True)
            o2 = F.conv2d(xi, wo, bo, stride=s, padding=p)
            o2.backward(goff2)
            gx_ref = gx2 + xi.grad

            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max())
            print(f"variant {variant} shape {(B,C,O,H,W,s)} rep {rep}: off {rel(off, o2.detach()):.2e} gx {rel(gx, gx_ref):.2e} "
                  f"gx_dcn_only {rel(gx - xi.grad, gx2):.2e} gwoff {rel(gwoff, wo.grad):.2e} gboff {rel(gboff, bo.grad):.2e} "
                  f"gw {rel(gw, gw2):.2e} gb {rel(gb, gb2):.2e}")
            if rel(gx, gx_ref) > 1e-3 and rep == 0:
                d = (gx - gx_ref).abs()
                per_b = d.amax(dim=(1, 2, 3)).tolist()
                per_c = d.amax(dim=(0, 2, 3))
                per_row = d.amax(dim=(0, 1, 3))
                per_col = d.amax(dim=(0, 1, 2))
                print("   per image:", [f"{v:.3f}" for v in per_b])
                print("   bad channels:", (per_c > 1e-3 * float(gx_ref.abs().max())).nonzero().flatten().tolist()[:64])
                print("   bad rows:", (per_row > 1e-3 * float(gx_ref.abs().max())).nonzero().flatten().tolist()[:140])
                print("   bad cols:", (per_col > 1e-3 * float(gx_ref.abs().max())).nonzero().flatten().tolist()[:140])
                print("   count bad:", int((d > 1e-3 * float(gx_ref.abs().max())).sum()), "of", d.numel())
