"""Full-size property tests (BASELINE.json configs at their quoted sizes).  The CPU oracle cannot
run these sizes in seconds, so parity is checked through size-independent properties:

  * two independent CUDA implementations agree (tcgen05 path vs the generic kernels that the
    small-size tests pin to the oracle / reference goldens);
  * adjoint identities of the operator, which is linear in weight and in x:
        <forward(W) - forward(0), g> = <W, grad_weight(g)>,   <forward(x) - forward(0), g> = <x, grad_x(g)>;
  * linearity in the weights;  grad_bias = column sums of g;
  * a directional finite difference for grad_offset (the sample is piecewise linear in the offsets);
  * the zero-offset closed form of the Torch variant on a square stride-1 layer
    (SURVEY.md 8c fact 5): every tap of output (h, w) samples x[b, c, ~w, ~h].
"""
import numpy as np
import pytest
import torch

import jittor_dcn_b200 as dcn

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: B, C, O, H, W, k, s, p
    "cfg2": (256, 64, 64, 128, 128, 3, 1, 1),       # BASELINE configs[1]
    "cfg3": (64, 256, 256, 28, 28, 3, 1, 1),        # BASELINE configs[2]
    "det2": (64, 16, 32, 128, 128, 3, 2, 1),        # detector conv2 (configs[0]/[4] layer)
    "det5": (64, 128, 256, 16, 16, 3, 2, 1),        # detector conv5
    "c3": (32, 128, 128, 56, 56, 3, 1, 1),          # BASELINE configs[3] layers (batch cut to 32: the generic
    "c4": (32, 256, 256, 28, 28, 3, 1, 1),          #   kernels are the yardstick and take seconds at 128)
    "c5": (32, 512, 512, 14, 14, 3, 1, 1),          #   O = 512: two output-channel groups on the tensor path
}


def _data(name, seed=0, sigma=2.0):
    B, C, O, H, W, k, s, p = CONFIGS[name]
    g = torch.Generator(device="cuda").manual_seed(seed)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    off = torch.randn(B, 2 * k * k, Ho, Wo, device="cuda", generator=g) * sigma
    wt = torch.randn(O, C, k, k, device="cuda", generator=g) * (2.0 / (C * k * k)) ** 0.5
    bias = torch.randn(O, device="cuda", generator=g)
    gout = torch.randn(B, O, Ho, Wo, device="cuda", generator=g)
    return (x, off, wt, bias, gout), (k, s, p)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_tensor_path_agrees_with_generic_kernels(name, variant):
    (x, off, wt, bias, gout), (k, s, p) = _data(name)
    out_u = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant)
    out_s = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    assert _rel(out_u, out_s) < 1e-4
    del out_u, out_s
    gu = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant)
    gs = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    for a, b, nm in zip(gu, gs, ("gx", "goff", "gw", "gb")):
        assert _rel(a, b) < 1e-3, nm


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_adjoint_identities_and_linearity(name, variant):
    (x, off, wt, bias, gout), (k, s, p) = _data(name, seed=1)
    out = dcn.dcn_forward(x, off, wt, None, k, s, p, variant)
    gx, goff, gw, gb = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant)
    lhs = _dot(out, gout)
    # <out, g> is a cancelling sum of ~1e8 zero-mean terms: the meaningful scale is |out|*|g|
    # (a wrong index anywhere shifts the sum by ~|out||g|/sqrt(numel) ~ 1e4 x the tolerance)
    scale = float(out.double().norm() * gout.double().norm())
    assert abs(lhs - _dot(wt, gw)) < 3e-8 * scale      # linear in W
    assert abs(lhs - _dot(x, gx)) < 3e-8 * scale       # linear in x
    assert _rel(gb, gout.sum(dim=(0, 2, 3))) < 1e-4
    # linearity in the weights
    w2 = torch.randn_like(wt) * 0.1
    out12 = dcn.dcn_forward(x, off, wt + w2, None, k, s, p, variant)
    out2 = dcn.dcn_forward(x, off, w2, None, k, s, p, variant)
    assert _rel(out12, out + out2) < 1e-4


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
def test_grad_offset_directional_finite_difference(variant):
    (x, off, wt, bias, gout), (k, s, p) = _data("cfg3", seed=2, sigma=1.0)
    _, goff, _, _ = dcn.dcn_backward(x, off, wt, gout, False, k, s, p, variant)
    # Direction ALIGNED with the analytic gradient, so <goff, d> is a sum of positive terms.  (With
    # a random direction the sum cancels to ~sqrt(numel) and the ~0.3 % of samples that cross a
    # cell boundary within +-eps — where the sample has a kink — alone move it by ~5 %.)
    d = goff / goff.abs().max()
    eps = 1e-3
    fp = _dot(dcn.dcn_forward(x, off + eps * d, wt, None, k, s, p, variant), gout)
    fm = _dot(dcn.dcn_forward(x, off - eps * d, wt, None, k, s, p, variant), gout)
    fd = (fp - fm) / (2 * eps)
    an = _dot(goff, d)
    assert an > 0 and abs(fd - an) < 3e-2 * an, (fd, an)


def test_zero_offset_closed_form_torch_variant():
    """All N taps of output (h, w) sample x[b, c, ~w, ~h] (transposed), so with zero offsets the
    layer is a 1x1 convolution of the transposed input with the tap-summed weights — up to the
    float32 round-trip wobble of the coordinates, which moves a sample by < 1e-4 pixel."""
    B, C, O, S = 8, 64, 64, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    # smooth input so that a 1e-5-pixel coordinate wobble stays below the tolerance
    base = torch.randn(B, C, S // 8, S // 8, device="cuda", generator=g)
    x = torch.nn.functional.interpolate(base, size=(S, S), mode="bilinear", align_corners=True)
    off = torch.zeros(B, 18, S, S, device="cuda")
    wt = torch.randn(O, C, 3, 3, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5
    out = dcn.dcn_forward(x, off, wt, None, 3, 1, 1, dcn.VARIANT_TORCH)
    # Torch layout: out_flat[r, :] = sum_j S_b.flat[r*K + j] * Wm[:, j] with S_b[c,h,w,n] = x[c,w,h]
    smp = x.transpose(2, 3).unsqueeze(-1).expand(B, C, S, S, 9).reshape(B, S * S, C * 9)
    ref = torch.matmul(smp, wt.reshape(O, -1).t()).reshape(B, S, S, O).permute(0, 3, 1, 2)
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_stack_layers_bf16_operands_against_fp32_generic_kernels(name):
    """BASELINE configs[3] (bf16 operands, fp32 accumulate) at the layers' real extents, Jittor layout: the
    tensor path in bf16 storage mode against the generic fp32 kernels on the SAME bf16-rounded operands.
    Tolerances as in test_gpu_parity.py: forward 1e-2, gradients 2e-2 (one bf16 rounding per sample)."""
    (x, off, wt, bias, gout), (k, s, p) = _data(name, seed=4)
    xb, wb, gb16 = x.bfloat16(), wt.bfloat16(), gout.bfloat16()
    v = dcn.VARIANT_JITTOR
    out = dcn.dcn_forward(xb, off, wb, bias, k, s, p, v, operand=dcn.OPERAND_BF16)
    ref = dcn.dcn_forward(xb.float(), off, wb.float(), bias, k, s, p, v, flags=dcn.FLAG_FORCE_SIMT)
    assert _rel(out, ref) < 1e-2
    got = dcn.dcn_backward(xb, off, wb, gb16, True, k, s, p, v, operand=dcn.OPERAND_BF16)
    exp = dcn.dcn_backward(xb.float(), off, wb.float(), gb16.float(), True, k, s, p, v, flags=dcn.FLAG_FORCE_SIMT)
    for a, b, nm in zip(got, exp, ("gx", "goff", "gw", "gb")):
        assert _rel(a, b) < 2e-2, nm


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
def test_wide_layer_fp32_agrees_with_generic_kernels(variant):
    (x, off, wt, bias, gout), (k, s, p) = _data("c5", seed=5)
    out_u = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant)
    out_s = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    assert _rel(out_u, out_s) < 1e-4
    gu = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant)
    gs = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    for a, b, nm in zip(gu, gs, ("gx", "goff", "gw", "gb")):
        assert _rel(a, b) < 1e-3, nm


@pytest.mark.parametrize("shape", [(256, 16, 128, 128), (256, 32, 64, 64), (1024, 256, 8, 8)])
def test_bn_relu_at_detector_sizes_against_the_framework(shape):
    """relu(bn(x)) at the detector's activation sizes (train.py:146-170) against the framework's own CUDA
    BatchNorm2d + ReLU: forward, running statistics and all three gradients."""
    B, C, H, W = shape
    torch.manual_seed(7)
    ref = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        ref.weight.normal_(1.0, 0.2)
        ref.bias.normal_(0.0, 0.3)
    ours = dcn.BatchNormReLU2d(C).cuda()
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(B, C, H, W, device="cuda") * 1.7 + 0.4
    gy = torch.randn(B, C, H, W, device="cuda")
    # Among 1e7..3e7 elements a few pre-activations sit within float32 round-off of zero, where the two
    # implementations' ReLU masks may legitimately differ; each such flip moves that element's input gradient
    # and the channel's grad_gamma / grad_beta by O(|grad_y|).  The incoming gradient is therefore zeroed
    # within 1e-4 of the kink (< 0.1 % of the elements), which makes every gradient comparable again.
    with torch.no_grad():
        pre = torch.nn.functional.batch_norm(x, None, None, ref.weight, ref.bias, True, 0.1, ref.eps)
        clear = (pre.abs() > 1e-4).float()
        assert float(clear.mean()) > 0.999
        gy = gy * clear
    xr = x.clone().requires_grad_(True)
    yr = torch.relu(ref(xr))
    yr.backward(gy)
    xo = x.clone().requires_grad_(True)
    yo = ours(xo)
    yo.backward(gy)
    assert _rel(yo.detach(), yr.detach()) < 1e-5
    assert _rel(xo.grad, xr.grad) < 1e-4
    assert _rel(ours.weight.grad, ref.weight.grad) < 1e-4
    assert _rel(ours.bias.grad, ref.bias.grad) < 1e-4
    assert _rel(ours.running_mean, ref.running_mean) < 1e-5
    assert _rel(ours.running_var, ref.running_var) < 1e-5
