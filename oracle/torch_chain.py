"""The reference's operator restated as the same chain of framework tensor ops.  TEST INFRA.

Why this exists next to the plain-C restatement (oracle/dcn_oracle.c): the reference spends
its time inside torch CPU kernels (grid_sampler_2d, mm, copy_ — SURVEY.md 3.4).  A scalar C
loop would be an unfairly slow CPU baseline, so ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs time THIS module: the same ops, the same N-fold materialisations,
the same autograd backward, on the GPU box's host cores.  /root/reference itself cannot
travel to the GPU box.

Pinned: tests/test_oracle_golden.py checks ``chain_forward(variant="torch")`` and its
autograd gradients bit-for-bit against tests/golden/ fixtures produced by the unmodified
reference class (same torch build => same kernels => identical bits).

  variant="torch"   follows train.py:95-140 op by op
  variant="jittor"  follows deform_conv.py:30-81 op by op with jt.* -> torch.*
                    (parity unpinned: jittor is not installable here)
"""
import torch
import torch.nn.functional as F


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def out_hw(H, W, k, s, p):
    (kh, kw), (sh, sw), (ph, pw) = _pair(k), _pair(s), _pair(p)
    return (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1


def chain_forward(x, offset, weight, bias, variant="torch", kernel_size=3, stride=1, padding=1):
    """x[B,C,H,W], offset[B,2N,Ho,Wo] (what offset_conv returned), weight[O,C,kh,kw]."""
    B, C, H, W = x.shape
    O = weight.shape[0]
    kh, kw = _pair(kernel_size)
    N = kh * kw
    Ho, Wo = offset.shape[2], offset.shape[3]

    # [B,2N,Ho,Wo] -> [B,Ho,Wo,N,2]          deform_conv.py:62 / train.py:102
    off = offset.view(B, 2, N, Ho, Wo).permute(0, 3, 4, 2, 1)
    # base grid: x = w, y = h, same for every tap          :64-66 / :104-107
    col = torch.arange(Wo, dtype=torch.float32).view(1, 1, Wo, 1).expand(B, Ho, Wo, N)
    row = torch.arange(Ho, dtype=torch.float32).view(1, Ho, 1, 1).expand(B, Ho, Wo, N)
    locs = torch.stack([col, row], dim=-1) + off  # :68 / :109

    if variant == "torch":       # train.py:111-112
        dx, dy = W - 1, H - 1
    elif variant == "jittor":    # deform_conv.py:34-38
        dx, dy = Wo - 1, Ho - 1
    else:
        raise ValueError(variant)
    nx = locs[..., 0] / dx * 2 - 1
    ny = locs[..., 1] / dy * 2 - 1
    grid = torch.stack([ny, nx], dim=-1)  # slot order as in :39 / :113

    # N physical copies of the input           :41-42 / :115-116
    xr = x.unsqueeze(1).repeat(1, N, 1, 1, 1).reshape(B * N, C, H, W)
    g = grid.permute(0, 3, 1, 2, 4).reshape(B * N, Ho, Wo, 2)  # :44-45 / :118-119
    smp = F.grid_sample(xr, g, mode="bilinear", padding_mode="zeros", align_corners=True)
    smp = smp.reshape(B, N, C, Ho, Wo).permute(0, 2, 3, 4, 1)  # [B,C,Ho,Wo,N]  :54 / :129

    if variant == "torch":
        cols = smp.reshape(B, Ho, Wo, -1).reshape(-1, C * N)      # train.py:130-131
    else:
        cols = smp.permute(0, 2, 3, 4, 1).reshape(B * Ho * Wo, N * C)  # deform_conv.py:72-73
    wm = weight.reshape(O, -1)
    flat = torch.matmul(cols, wm.t())  # :74-76 / :133-134
    out = flat.reshape(B, Ho, Wo, O).permute(0, 3, 1, 2)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def chain_forward_backward(x, offset, weight, bias, gout, **kw):
    """-> out, (gx, goff, gw, gb) through torch autograd, like train.py:249."""
    x = x.detach().clone().requires_grad_(True)
    offset = offset.detach().clone().requires_grad_(True)
    weight = weight.detach().clone().requires_grad_(True)
    leaves = [x, offset, weight]
    if bias is not None:
        bias = bias.detach().clone().requires_grad_(True)
        leaves.append(bias)
    out = chain_forward(x, offset, weight, bias, **kw)
    grads = torch.autograd.grad(out, leaves, gout)
    return out.detach(), grads


class ChainLayer(torch.nn.Module):
    """A whole layer (offset conv + chain) with the reference's parameter names — used by the
    CPU-baseline timing so that its work equals ``TorchDeformConv2d.forward`` + backward."""

    def __init__(self, C, O, kernel_size=3, stride=1, padding=1, bias=True, variant="torch"):
        super().__init__()
        self.k, self.s, self.p, self.variant = _pair(kernel_size), _pair(stride), _pair(padding), variant
        N = self.k[0] * self.k[1]
        self.offset_conv = torch.nn.Conv2d(C, 2 * N, self.k, self.s, self.p)
        self.weight = torch.nn.Parameter(torch.empty(O, C, *self.k))
        self.bias = torch.nn.Parameter(torch.zeros(O)) if bias else None

    def forward(self, x):
        return chain_forward(x, self.offset_conv(x), self.weight, self.bias, variant=self.variant,
                             kernel_size=self.k, stride=self.s, padding=self.p)
