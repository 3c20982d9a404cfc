"""Fuzz of the companion offset convolution on the engine (csrc/dcn_conv.cu shifted-view kernels for C % 64 == 0,
csrc/dcn_conv_small.cu warp-MMA kernels for narrower layers, plain mode of the DCN kernels for what is left) over image widths that exercise every packing of the 128 tensor-core rows (1 .. 8
output rows per tile, partial tiles, more than one column segment per row), both strides, both staging layouts:
forward against `conv2d` in float64, backward (data + weight + bias gradient through dcn_layer_backward, DCN span
gradient subtracted) against torch's conv autograd."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import jittor_dcn_b200 as dcn
from jittor_dcn_b200.functional import dcn_layer_backward, dcn_layer_forward, layer_supported

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(77)
    out = []
    sizes = [(7, 9), (14, 14), (16, 30), (28, 28), (31, 33), (56, 56), (40, 64), (25, 100), (12, 126), (9, 128), (6, 130),
             (5, 200), (4, 257), (64, 62), (17, 127)]
    for (H, W) in sizes:
        for s in (1, 2):
            C = int(rng.choice([64, 64, 128, 16, 32, 16, 32, 48]))
            O = int(rng.choice([32, 64]))
            B = int(rng.integers(1, 4))
            variant = int(rng.integers(0, 2))
            out.append((B, C, O, H, W, s, variant))
    return out


@pytest.mark.parametrize("case", _cases())
def test_offset_conv_on_the_engine(case):
    B, C, O, H, W, s, variant = case
    if not layer_supported((B, C, H, W), O, 3, s, 1, variant):
        pytest.skip("whole-layer entry points do not cover this shape")
    g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xffff)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    wo = torch.randn(18, C, 3, 3, device="cuda", generator=g) * 0.05
    bo = torch.randn(18, device="cuda", generator=g)
    wt = torch.randn(O, C, 3, 3, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5
    Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
    gout = torch.randn(B, O, Ho, Wo, device="cuda", generator=g)
    off, out = dcn_layer_forward(x, wo, bo, wt, None, 3, s, 1, variant)
    ref = F.conv2d(x.double(), wo.double(), bo.double(), stride=s, padding=1)
    assert float((off.double() - ref).abs().max() / ref.abs().max()) < 2e-5, "offset conv forward"
    gx, gwo, gbo, gw, gb = dcn_layer_backward(x, off, wo, wt, gout, True, False, 3, s, 1, variant)
    gx_span, goff, gw_span, _ = dcn.dcn_backward(x, off, wt, gout, False, 3, s, 1, variant)
    x2, wo2, bo2 = x.clone().requires_grad_(True), wo.clone().requires_grad_(True), bo.clone().requires_grad_(True)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        F.conv2d(x2, wo2, bo2, stride=s, padding=1).backward(goff)
    finally:
        torch.backends.cudnn.allow_tf32 = prev

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(gx, gx_span + x2.grad) < 1e-3, "data gradient"
    assert rel(gwo, wo2.grad) < 1e-3, "offset-conv weight gradient"
    assert rel(gbo, bo2.grad) < 1e-3, "offset-conv bias gradient"
    assert rel(gw, gw_span) < 1e-4


@pytest.mark.parametrize("C,s", [(16, 2), (32, 2), (16, 1), (48, 1)])
def test_narrow_layers_run_the_warp_mma_kernels(C, s):
    """The detector's 16- and 32-channel layers: the companion conv must run on dcn_conv_small.cu (forward, and for
    C < 64 both backward kernels), not on the plain mode of the DCN kernels."""
    from jittor_dcn_b200 import _lib
    B, O, H, W = 2, 32, 32, 32
    variant = dcn.VARIANT_TORCH
    if not layer_supported((B, C, H, W), O, 3, s, 1, variant):
        pytest.skip("whole-layer entry points do not cover this shape")
    g = torch.Generator(device="cuda").manual_seed(C * 10 + s)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    wo = torch.randn(18, C, 3, 3, device="cuda", generator=g) * 0.05
    bo = torch.randn(18, device="cuda", generator=g)
    wt = torch.randn(O, C, 3, 3, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5
    Ho = (H + 2 - 3) // s + 1
    gout = torch.randn(B, O, Ho, Ho, device="cuda", generator=g)
    _lib.profile_begin()
    off, _ = dcn_layer_forward(x, wo, bo, wt, None, 3, s, 1, variant)
    dcn_layer_backward(x, off, wo, wt, gout, True, False, 3, s, 1, variant)
    torch.cuda.synchronize()
    names = set(_lib.profile_end())
    assert "conv_small_fwd_kernel" in names, names
    assert {"conv_small_dgrad_kernel", "conv_small_wgrad_kernel"} <= names, names
    assert not any(n.startswith("umma_offset_conv") for n in names), names
