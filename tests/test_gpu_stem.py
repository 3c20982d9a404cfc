"""GPU tests of the detector's conv1 on the engine (csrc/dcn_stem.cu; train.py:145,166) against the framework's conv2d:
fp32 both sides, the only difference is the summation order (forward 1e-5, gradients 1e-4 relative)."""
import pytest
import torch
import torch.nn as nn

import jittor_dcn_b200 as dcn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_framework_conv():
    """The framework's conv may run in TF32 by default (cudnn.allow_tf32); the comparison is fp32 against fp32."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [
    (8, 1, 16, 128, 128),      # the detector's conv1 at a small batch
    (3, 3, 32, 20, 36),        # three input channels, O = 32, a partial 128-pixel segment
    (2, 2, 16, 5, 8),          # tiny: every pixel touches the padding
    (2, 1, 16, 7, 260),        # rows longer than one warp's 128-pixel segment
    (1, 4, 16, 1, 4),          # a single row
])
def test_stem_conv_matches_the_framework_conv(shape):
    B, Cin, O, H, W = shape
    torch.manual_seed(B * 1000 + W)
    ref = nn.Conv2d(Cin, O, 3, 1, 1).cuda()
    eng = dcn.StemConv2d(Cin, O, 3, 1, 1).cuda()
    eng.load_state_dict(ref.state_dict())
    x = torch.randn(B, Cin, H, W, device="cuda")
    g = torch.randn(B, O, H, W, device="cuda")
    lib = dcn.load()
    lib.dcn_launch_count_reset()
    y = eng(x)
    assert int(lib.dcn_launch_count()) == 1, "the engine kernel must be the one that ran"
    y0 = ref(x)
    assert _rel(y, y0) < 1e-5, _rel(y, y0)
    y.backward(g)
    y0.backward(g)
    assert _rel(eng.weight.grad, ref.weight.grad) < 1e-4, _rel(eng.weight.grad, ref.weight.grad)
    assert _rel(eng.bias.grad, ref.bias.grad) < 1e-4


def test_stem_conv_without_bias_and_fallbacks():
    torch.manual_seed(1)
    lib = dcn.load()
    ref = nn.Conv2d(1, 16, 3, 1, 1, bias=False).cuda()
    eng = dcn.StemConv2d(1, 16, 3, 1, 1, bias=False).cuda()
    eng.load_state_dict(ref.state_dict())
    x = torch.randn(4, 1, 16, 16, device="cuda")
    assert _rel(eng(x), ref(x)) < 1e-5
    # an input that wants a gradient, a row length the kernels do not take, another stride: the framework's conv
    lib.dcn_launch_count_reset()
    xg = x.clone().requires_grad_(True)
    eng(xg).sum().backward()
    assert xg.grad is not None and int(lib.dcn_launch_count()) == 0
    x6 = torch.randn(2, 1, 8, 6, device="cuda")
    assert torch.equal(eng(x6), ref(x6)) and int(lib.dcn_launch_count()) == 0
    s2 = dcn.StemConv2d(1, 16, 3, 2, 1).cuda()
    assert s2(x).shape == (4, 16, 8, 8) and int(lib.dcn_launch_count()) == 0


def test_detector_uses_the_engine_for_conv1_and_keeps_the_state_dict():
    from jittor_dcn_b200.detector import EDNetDetection
    m = EDNetDetection(fused_bn_relu=True)
    assert isinstance(m.conv1, dcn.StemConv2d)
    plain = EDNetDetection(fused_bn_relu=False)
    assert type(plain.conv1) is nn.Conv2d
    assert list(m.state_dict().keys()) == list(plain.state_dict().keys())
