"""The whole layer on the engine: companion offset convolution (deform_conv.py:16-21,58 /
train.py:80-85,98) as a plain mode of the tcgen05 kernels + the DCN span, one autograd node.

  * offsets against the reference's own operator for that step, `conv2d` (torch CPU, float64);
  * the module against whole-module goldens made by the UNMODIFIED reference class (module_umma_*, live offset conv):
    state dict in, output / input gradient / all four parameter gradients out — north_star's tolerances;
  * the one-node layer against the two-node path (framework offset conv + engine span) at larger sizes and for the
    Jittor-semantics variant (whose reference cannot run here).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import jittor_dcn_b200 as dcn
from tests.util import golden, golden_names, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR, dcn.VARIANT_DCNV1])
@pytest.mark.parametrize("shape", [
    # B, C,  O,   H,  W,  k, s, p
    (2, 64, 64, 16, 16, 3, 1, 1),
    (3, 16, 32, 32, 32, 3, 2, 1),      # detector conv2 channels, stride 2
    (2, 32, 64, 20, 12, 3, 2, 1),      # non-square, partial last tile
    (1, 128, 32, 12, 12, 3, 1, 1),     # Torch layout: permuted staging, Cs = 8
    (2, 256, 256, 7, 7, 3, 1, 1),      # cfg3 channels
    (2, 64, 16, 9, 11, (1, 3), 1, (0, 1)),   # 6 offset channels (2N = 6 of 16 accumulator columns)
    (2, 64, 64, 10, 10, 3, 1, 2),      # padding 2: taps beyond the one-pixel frame
])
def test_offset_conv_forward_matches_conv2d(shape, variant):
    B, C, O, H, W, k, s, p = shape
    if not dcn.layer_supported((B, C, H, W), O, k, s, p, variant):
        pytest.skip("layer path does not cover this shape / layout")
    kh, kw = (k, k) if isinstance(k, int) else k
    g = torch.Generator().manual_seed(hash(shape) & 0xffff)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(2 * kh * kw, C, kh, kw, generator=g) * 0.05
    b = torch.randn(2 * kh * kw, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p)
    got = dcn.dcn_offset_conv_forward(x.cuda(), w.cuda(), b.cuda(), O, k, s, p, variant)
    assert tuple(got.shape) == tuple(ref.shape)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 2e-5
    got_nb = dcn.dcn_offset_conv_forward(x.cuda(), w.cuda(), None, O, k, s, p, variant)
    assert rel_err(got_nb.cpu().numpy(), F.conv2d(x.double(), w.double(), None, stride=s, padding=p).numpy()) < 2e-5


@pytest.mark.parametrize("name", golden_names("module_umma_"))
def test_whole_module_on_the_engine_matches_reference(name):
    g = golden(name)
    C, O, H, W, s = (int(v) for v in g["cfg"])
    m = dcn.TorchDeformConv2d(C, O, 3, s, 1).cuda()
    m.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("sd.")})
    x = _cuda(g["x"]).requires_grad_(True)
    assert m._whole_layer_on_engine(x), "this fixture is meant to run offset conv + DCN span on the engine"
    off = dcn.dcn_offset_conv_forward(x.detach(), m.offset_conv.weight.detach(), m.offset_conv.bias.detach(), O, 3, s, 1)
    assert rel_err(off.cpu().numpy(), g["offset"]) < 2e-5
    for keep in (False, True):
        m.keep_staged_input = keep
        m.zero_grad()
        x.grad = None
        out = m(x)
        out.backward(_cuda(g["gout"]))
        assert rel_err(out.detach().cpu().numpy(), g["out"]) < FWD_TOL
        assert rel_err(x.grad.cpu().numpy(), g["gx"]) < GRAD_TOL
        for pname, prm in m.named_parameters():
            assert rel_err(prm.grad.cpu().numpy(), g["grad." + pname]) < GRAD_TOL, pname


@pytest.mark.parametrize("variant_cls", [dcn.TorchDeformConv2d, dcn.TorchDeformConv2dJittorSemantics])
@pytest.mark.parametrize("shape", [(4, 64, 64, 128, 128, 1), (8, 256, 256, 28, 28, 1), (8, 16, 32, 128, 128, 2),
                                   (16, 128, 256, 16, 16, 2), (8, 128, 128, 56, 56, 1)])
def test_one_node_layer_equals_framework_offset_conv_plus_engine_span(shape, variant_cls):
    """BASELINE layer shapes (reduced batch): the whole layer on the engine against the round-1 composition —
    framework offset conv (forward value and autograd), engine DCN span, framework add of the two input-gradient terms.

    Both sides sample at the SAME offsets (the engine's; checked against the framework conv to 2e-5 first): the
    coordinate gradient is piecewise constant in the sampling position, so offsets that differ in the last bits (two
    correct fp32 convolutions with different summation orders) flip floor() for a handful of the 10^5..10^6 samples,
    and each flip moves one grad_offset element — and through the offset conv's backward a 3x3xC patch of grad_x — by
    up to ~1 % of max|grad_x| (measured 1.5-4 % at the 128x128 shapes when the framework's own offsets were used)."""
    B, C, O, H, W, s = shape
    torch.manual_seed(3)
    m = variant_cls(C, O, 3, s, 1).cuda()
    with torch.no_grad():
        m.offset_conv.weight.normal_(0, 0.01)
        m.offset_conv.bias.normal_(0, 1.0)
        m.bias.normal_(0, 0.1)
    x = torch.randn(B, C, H, W, device="cuda")
    gout = torch.randn(B, O, (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1, device="cuda")
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # the yardstick conv must be a real fp32 convolution
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        # the module, whole layer on the engine (one autograd node)
        xi = x.clone().requires_grad_(True)
        assert m._whole_layer_on_engine(xi)
        out = m(xi)
        out.backward(gout)
        got = [out.detach(), xi.grad] + [p.grad.clone() for p in m.parameters()]
        names = ["out", "gx"] + [n for n, _ in m.named_parameters()]
        # the composition, on the engine's offsets
        ow, ob = m.offset_conv.weight.detach(), m.offset_conv.bias.detach()
        off = dcn.dcn_offset_conv_forward(x, ow, ob, O, 3, s, 1, m.variant)
        x2 = x.clone().requires_grad_(True)
        ow2, ob2 = ow.clone().requires_grad_(True), ob.clone().requires_grad_(True)
        off_fw = F.conv2d(x2, ow2, ob2, stride=s, padding=1)
        assert rel_err(off.cpu().numpy(), off_fw.detach().cpu().numpy()) < 2e-5
        out2 = dcn.dcn_forward(x, off, m.weight.detach(), m.bias.detach(), 3, s, 1, m.variant)
        gx_dcn, goff, gw, gb = dcn.dcn_backward(x, off, m.weight.detach(), gout, True, 3, s, 1, m.variant)
        off_fw.backward(goff)
        ref = {"out": out2, "gx": gx_dcn + x2.grad, "weight": gw, "bias": gb, "offset_conv.weight": ow2.grad,
               "offset_conv.bias": ob2.grad}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    for a, nm in zip(got, names):
        tol = FWD_TOL if nm == "out" else GRAD_TOL
        assert rel_err(a.cpu().numpy(), ref[nm].cpu().numpy()) < tol, nm


def test_layer_backward_without_grad_x():
    torch.manual_seed(4)
    m = dcn.TorchDeformConv2d(16, 32, 3, 2, 1).cuda()
    with torch.no_grad():
        m.offset_conv.weight.normal_(0, 0.02)
        m.offset_conv.bias.normal_(0, 0.7)
    x = torch.randn(4, 16, 32, 32, device="cuda")          # no requires_grad: first layer of a net
    gout = torch.randn(4, 32, 16, 16, device="cuda")
    m(x).backward(gout)
    got = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    xi = x.clone().requires_grad_(True)
    m(xi).backward(gout)
    for a, b in zip(got, [p.grad for p in m.parameters()]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
