"""BatchNorm2d + ReLU post-op (csrc/dcn_bn.cu) against the reference's own modules for that span:
``relu(bn(x))`` with ``tnn.BatchNorm2d`` / ``tnn.ReLU`` (train.py:146-159, 167-170), evaluated by torch on the CPU
in float64 (forward) / float32 autograd.  Tolerances (max-abs error over max-abs value): forward and running
statistics 1e-5, gradients 1e-4 — float32 arithmetic with a different summation order."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import jittor_dcn_b200 as dcn
from tests.util import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [(4, 16, 32, 32), (3, 5, 7, 9), (8, 256, 8, 8), (2, 32, 64, 64), (6, 3, 1, 2), (5, 8, 6, 6)]


def _pair(C, seed):
    torch.manual_seed(seed)
    ref = nn.BatchNorm2d(C).double()
    with torch.no_grad():
        ref.weight.normal_(1.0, 0.3)
        ref.bias.normal_(0.0, 0.5)
        ref.running_mean.normal_(0.0, 0.2)
        ref.running_var.uniform_(0.5, 1.5)
    ours = dcn.BatchNormReLU2d(C)
    ours.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref.state_dict().items()})
    return ref, ours.cuda()


@pytest.mark.parametrize("shape", SHAPES)
def test_training_forward_backward_and_running_stats(shape):
    B, C, H, W = shape
    ref, ours = _pair(C, 3)
    for step in range(2):   # two steps: the running statistics must follow torch's update rule
        x = torch.randn(B, C, H, W) * 2.0 + 0.7
        gy = torch.randn(B, C, H, W)
        xr = x.double().requires_grad_(True)
        yr = torch.relu(ref(xr))
        yr.backward(gy.double())
        xo = x.cuda().requires_grad_(True)
        yo = ours(xo)
        assert rel_err(yo.detach().cpu().numpy(), yr.detach().numpy()) < 1e-5
        ours.zero_grad()
        yo.backward(gy.cuda())
        assert rel_err(xo.grad.cpu().numpy(), xr.grad.numpy()) < 1e-4
        assert rel_err(ours.weight.grad.cpu().numpy(), ref.weight.grad.numpy()) < 1e-4
        assert rel_err(ours.bias.grad.cpu().numpy(), ref.bias.grad.numpy()) < 1e-4
        ref.zero_grad()
    assert rel_err(ours.running_mean.cpu().numpy(), ref.running_mean.numpy()) < 1e-5
    assert rel_err(ours.running_var.cpu().numpy(), ref.running_var.numpy()) < 1e-5
    assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked)


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_eval_mode_uses_running_statistics(shape):
    B, C, H, W = shape
    ref, ours = _pair(C, 4)
    ref.eval()
    ours.eval()
    x = torch.randn(B, C, H, W)
    gy = torch.randn(B, C, H, W)
    xr = x.double().requires_grad_(True)
    yr = torch.relu(ref(xr))
    yr.backward(gy.double())
    xo = x.cuda().requires_grad_(True)
    yo = ours(xo)
    yo.backward(gy.cuda())
    assert rel_err(yo.detach().cpu().numpy(), yr.detach().numpy()) < 1e-5
    assert rel_err(xo.grad.cpu().numpy(), xr.grad.numpy()) < 1e-4
    assert rel_err(ours.weight.grad.cpu().numpy(), ref.weight.grad.numpy()) < 1e-4
    assert rel_err(ours.bias.grad.cpu().numpy(), ref.bias.grad.numpy()) < 1e-4
    assert torch.equal(ours.running_mean.cpu(), ref.running_mean.float())   # untouched in eval mode


def test_one_value_per_channel_is_refused_like_torch():
    m = dcn.BatchNormReLU2d(3).cuda()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 1, 1, device="cuda"))
    m.eval()
    assert m(torch.zeros(1, 3, 1, 1, device="cuda")).shape == (1, 3, 1, 1)


def test_relu_mask_of_backward_matches_forward_bitwise():
    """The backward pass recomputes the ReLU mask from x: wherever y == 0 the input gradient must carry no
    grad_y term.  With gamma > 0 and a one-hot grad_y on a clamped element, dx of a channel is then exactly 0."""
    torch.manual_seed(0)
    m = dcn.BatchNormReLU2d(4).cuda()
    x = torch.randn(2, 4, 8, 8, device="cuda", requires_grad=True)
    y = m(x)
    dead = (y == 0).nonzero()[0]
    gy = torch.zeros_like(y)
    gy[tuple(dead)] = 1.0
    y.backward(gy)
    assert float(x.grad.abs().max()) == 0.0


def test_host_tensors_and_state_dict_drop_in():
    ref = nn.BatchNorm2d(8)
    ours = dcn.BatchNormReLU2d(8)
    ours.load_state_dict(ref.state_dict())           # same keys and shapes as nn.BatchNorm2d
    x = torch.randn(4, 8, 10, 10)
    yr = torch.relu(ref(x))
    yo = ours(x)                                     # CPU tensors in, CPU tensors out (staged through the GPU)
    assert yo.device.type == "cpu"
    assert rel_err(yo.detach().numpy(), yr.detach().numpy()) < 1e-5
    assert rel_err(ours.running_var.numpy(), ref.running_var.numpy()) < 1e-5


def test_detector_fused_matches_framework_bn():
    """The harness detector with the fused post-op == the same detector with the framework's BatchNorm2d + ReLU."""
    from jittor_dcn_b200.detector import EDNetDetection, detection_loss, synthetic_canvases
    torch.manual_seed(1)
    a = EDNetDetection(fused_bn_relu=True).cuda()
    b = EDNetDetection(fused_bn_relu=False).cuda()
    b.load_state_dict(a.state_dict())
    x, labels, boxes = synthetic_canvases(8, torch.Generator().manual_seed(2), "cuda")
    la = detection_loss(*a(x), labels, boxes)
    lb = detection_loss(*b(x), labels, boxes)
    la.backward()
    lb.backward()
    assert abs(float(la) - float(lb)) < 1e-4 * abs(float(lb))
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if float(q.grad.abs().max()) < 1e-5:
            # a bias in front of a BatchNorm has zero gradient (the batch mean removes it): pure round-off
            assert float(p.grad.abs().max()) < 1e-5, n
            continue
        assert rel_err(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 2e-3, n
    for (n, p), (_, q) in zip(a.named_buffers(), b.named_buffers()):
        assert rel_err(p.float().cpu().numpy(), q.float().cpu().numpy()) < 1e-4, n


def test_batch_variance_with_mean_far_above_std():
    """A channel at mean 1e3, std 1e-2: E[x^2] - E[x]^2 in float32 loses every digit of the variance;
    the kernels accumulate sums shifted by a per-channel pivot instead.  Reference: float64 two-pass statistics.
    x itself carries a relative spacing of 6e-8 * 1e3 = 6e-5 around the mean, i.e. 0.6 % of std: the normalised
    output is compared where that input quantisation is the same for both sides (same float32 x)."""
    torch.manual_seed(11)
    B, C, H, W = 8, 6, 32, 32
    mean = torch.tensor([1e3, -1e3, 50.0, 0.0, 1e3, 3.0]).view(1, C, 1, 1)
    std = torch.tensor([1e-2, 1e-2, 1e-3, 1.0, 5.0, 1e-4]).view(1, C, 1, 1)
    x = (torch.randn(B, C, H, W, dtype=torch.float64) * std + mean).float()
    ours = dcn.BatchNormReLU2d(C).cuda()
    y = ours(x.cuda())
    xd = x.double()
    m = xd.mean(dim=(0, 2, 3))
    v = xd.var(dim=(0, 2, 3), unbiased=False)
    ref = torch.relu((xd - m.view(1, C, 1, 1)) / torch.sqrt(v.view(1, C, 1, 1) + ours.eps))
    assert rel_err(y.detach().cpu().numpy(), ref.numpy()) < 1e-3
    # running_var = 0.9 * 1 + 0.1 * unbiased batch variance
    n = B * H * W
    exp_rv = 0.9 + 0.1 * v * n / (n - 1)
    assert float(((ours.running_var.cpu().double() - exp_rv).abs() / exp_rv).max()) < 1e-5
    exp_rm = 0.1 * m
    assert float((ours.running_mean.cpu().double() - exp_rm).abs().max()) < 1e-4


@pytest.mark.parametrize("variant_cls", [dcn.TorchDeformConv2d, dcn.TorchDeformConv2dJittorSemantics])
@pytest.mark.parametrize("shape", [(2, 64, 64, 32, 32, 1), (2, 16, 32, 32, 32, 2), (2, 12, 20, 9, 11, 1)])
@pytest.mark.parametrize("engine_offset_conv", [True, False])
def test_relu_out_flag_and_folded_eval_batchnorm(shape, variant_cls, engine_offset_conv):
    """DCN_FLAG_RELU_OUT: the forward epilogue applies the ReLU on every kernel family (tensor path
    with and without the staged tensor-map store, whole-layer path, generic kernels for the odd shape), the autograd
    nodes mask grad_out, and fuse_eval_bn_relu(layer, bn) reproduces relu(bn(layer(x))) in eval mode."""
    B, C, O, H, W, s = shape
    torch.manual_seed(21)
    layer = variant_cls(C, O, 3, s, 1).cuda()
    layer.engine_offset_conv = engine_offset_conv
    with torch.no_grad():
        layer.offset_conv.weight.normal_(0, 0.02)
        layer.offset_conv.bias.normal_(0, 0.8)
        layer.bias.normal_(0, 0.3)
    x = torch.randn(B, C, H, W, device="cuda")
    # (a) the flag alone
    plain = layer(x).detach()
    relu_layer = variant_cls(C, O, 3, s, 1).cuda()
    relu_layer.load_state_dict(layer.state_dict())
    relu_layer.engine_offset_conv = engine_offset_conv
    relu_layer.engine_flags = dcn.FLAG_RELU_OUT
    xa = x.clone().requires_grad_(True)
    ya = relu_layer(xa)
    assert torch.equal(ya.detach(), torch.relu(plain))
    gout = torch.randn_like(ya)
    ya.backward(gout)
    xb = x.clone().requires_grad_(True)
    layer.zero_grad()
    torch.relu(layer(xb)).backward(gout)
    assert rel_err(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) < 1e-5
    for (n, pa), (_, pb) in zip(relu_layer.named_parameters(), layer.named_parameters()):
        assert rel_err(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < 1e-5, n
    # (b) eval-mode BatchNorm folded into the layer
    bn = torch.nn.BatchNorm2d(O).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.normal_(1.0, 0.2)
        bn.bias.normal_(0, 0.3)
    bn.eval()
    with pytest.raises(ValueError):
        dcn.fuse_eval_bn_relu(layer, torch.nn.BatchNorm2d(O).cuda().train())
    fused = dcn.fuse_eval_bn_relu(layer, bn)
    with torch.no_grad():
        ref = torch.relu(bn(layer(x)))
        got = fused(x)
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5
