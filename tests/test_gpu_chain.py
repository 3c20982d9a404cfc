"""Chained inference: channels-last activations between consecutive engine layers.  The chained
detector forward against the unmodified reference's eval outputs (tests/golden/detector_eval.npz), against the
layer-by-layer module path, and the launch list: ONE nchw -> channels-last pass for the four DCN stages."""
import numpy as np
import pytest
import torch

import jittor_dcn_b200 as dcn
from jittor_dcn_b200 import _lib
from jittor_dcn_b200.detector import EDNetDetection, chained_eval_forward
from tests.util import golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant_cls", [dcn.TorchDeformConv2d, dcn.TorchDeformConv2dJittorSemantics])
@pytest.mark.parametrize("fused_bn_relu", [True, False])
def test_chained_detector_forward(variant_cls, fused_bn_relu):
    ev = golden("detector_eval")
    m = EDNetDetection(dcn_cls=variant_cls, fused_bn_relu=fused_bn_relu).cuda().eval()
    m.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in ev.items() if k.startswith("sd.")})
    with torch.no_grad():                                   # live offsets and non-trivial statistics
        torch.manual_seed(4)
        for mod in m.modules():
            if isinstance(mod, dcn.TorchDeformConv2d):
                mod.offset_conv.weight.normal_(0, 0.01)
                mod.offset_conv.bias.normal_(0, 0.7)
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
    x = torch.as_tensor(np.ascontiguousarray(ev["x"])).cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref_cls, ref_box = m(x)
        _lib.profile_begin()
        cls, box = chained_eval_forward(m, x)
        torch.cuda.synchronize()
        prof = _lib.profile_end()
        cls2, box2 = chained_eval_forward(m, x)             # second call: cached workspaces, frames still zero
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert rel_err(cls.cpu().numpy(), ref_cls.cpu().numpy()) < 1e-4
    assert rel_err(box.cpu().numpy(), ref_box.cpu().numpy()) < 1e-4
    assert torch.equal(cls, cls2) and torch.equal(box, box2)
    # the four DCN stages share ONE layout pass (the first stage's), and none of them wrote NCHW except the last
    assert prof["nchw_to_nhwc_kernel"][0] == 1, prof
    assert "nhwc_to_nchw_kernel" not in prof
    n_dcn = prof["umma_fwd_kernel"][0]
    assert n_dcn == 4, prof


def test_chained_detector_matches_the_reference_eval_golden():
    """State dict of the unmodified reference detector (zero offsets, fresh BatchNorm statistics): class logits and
    boxes of its eval forward."""
    ev = golden("detector_eval")
    m = EDNetDetection().cuda().eval()
    m.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in ev.items() if k.startswith("sd.")})
    x = torch.as_tensor(np.ascontiguousarray(ev["x"])).cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        cls, box = chained_eval_forward(m, x)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert rel_err(cls.cpu().numpy(), ev["cls"]) < 1e-3
    assert rel_err(box.cpu().numpy(), ev["bbox"]) < 1e-3


def test_chained_forward_refuses_training_mode_and_cpu():
    m = EDNetDetection().cuda().train()
    with pytest.raises(ValueError):
        chained_eval_forward(m, torch.zeros(2, 1, 128, 128, device="cuda"))
    chain = dcn.ChainedDeformStages([(dcn.TorchDeformConv2d(16, 32, 3, 2, 1).cuda(), torch.nn.BatchNorm2d(32).cuda().eval())])
    with pytest.raises(ValueError):
        chain(torch.zeros(2, 16, 32, 32))


# ---- training: channels-last hand-over between a layer's post-op and the next layer ------------------------------
@pytest.mark.parametrize("variant_cls", [dcn.TorchDeformConv2d, dcn.TorchDeformConv2dJittorSemantics])
def test_training_step_channels_last_equals_nchw_path(variant_cls):
    """EDNetDetection(channels_last=True): relu(bn(x)) writes the next DCN layer's staged input directly and takes the
    gradient as that layer's channels-last grad_x accumulator.  One training step must reproduce the NCHW-path step
    (same kernels for the arithmetic; only the layout passes differ): loss, outputs, running statistics, all gradients —
    and its launch list holds NO nchw <-> channels-last transposition."""
    from jittor_dcn_b200.detector import detection_loss, synthetic_canvases
    torch.manual_seed(7)
    ref = EDNetDetection(dcn_cls=variant_cls, fused_bn_relu=True, channels_last=False).cuda().train()
    cl = EDNetDetection(dcn_cls=variant_cls, fused_bn_relu=True, channels_last=True).cuda().train()
    with torch.no_grad():
        for mod in ref.modules():
            if isinstance(mod, dcn.TorchDeformConv2d):
                mod.offset_conv.weight.normal_(0, 0.01)
                mod.offset_conv.bias.normal_(0, 0.7)
    cl.load_state_dict(ref.state_dict())
    x, labels, boxes = synthetic_canvases(6, torch.Generator().manual_seed(3), "cuda")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out_r = ref(x)
        loss_r = detection_loss(*out_r, labels, boxes)
        loss_r.backward()
        _lib.profile_begin()
        out_c = cl(x)
        loss_c = detection_loss(*out_c, labels, boxes)
        loss_c.backward()
        torch.cuda.synchronize()
        prof = _lib.profile_end()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert "nchw_to_nhwc_kernel" not in prof and "nhwc_to_nchw_kernel" not in prof, sorted(prof)
    assert prof["bn_relu_stage_kernel"][0] == 4 and prof["bn_bwd_unstage_kernel"][0] == 4, prof
    assert abs(float(loss_c) - float(loss_r)) < 1e-5 * abs(float(loss_r))
    for a, b in zip(out_c, out_r):
        assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 1e-5
    for (k, v), (_, w) in zip(cl.state_dict().items(), ref.state_dict().items()):
        if "running_" in k:
            assert rel_err(v.cpu().numpy(), w.cpu().numpy()) < 1e-5, k
    for (k, p), (_, q) in zip(cl.named_parameters(), ref.named_parameters()):
        scale = float(q.grad.abs().max())
        if scale < 1e-4:                 # biases in front of a BatchNorm: round-off of a zero sum on both sides
            assert float(p.grad.abs().max()) < 1e-3, k
        else:
            # atomics order differs between runs: float re-association, nothing else
            assert rel_err(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 2e-4, k


def test_staged_post_op_alone_matches_the_nchw_post_op():
    """BatchNormReLU2d.forward_staged against BatchNormReLU2d.forward + the layer's own staging, forward and backward,
    for a Torch-layout consumer (channel-permuted staging copy) and a pixel-row one."""
    for cls in (dcn.TorchDeformConv2d, dcn.TorchDeformConv2dJittorSemantics):
        torch.manual_seed(5)
        bn_a, bn_b = dcn.BatchNormReLU2d(32).cuda().train(), dcn.BatchNormReLU2d(32).cuda().train()
        layer_a, layer_b = cls(32, 64, 3, 2, 1).cuda(), cls(32, 64, 3, 2, 1).cuda()
        with torch.no_grad():
            bn_a.weight.normal_(1.0, 0.2)
            bn_a.bias.normal_(0, 0.3)
            layer_a.offset_conv.weight.normal_(0, 0.01)
            layer_a.offset_conv.bias.normal_(0, 0.7)
        bn_b.load_state_dict(bn_a.state_dict())
        layer_b.load_state_dict(layer_a.state_dict())
        x = torch.randn(3, 32, 24, 24, device="cuda")
        gout = torch.randn(3, 64, 12, 12, device="cuda")
        xa = x.clone().requires_grad_(True)
        ya = layer_a(bn_a(xa))
        ya.backward(gout)
        xb = x.clone().requires_grad_(True)
        yb = layer_b.forward_staged(bn_b.forward_staged(xb, layer_b), (24, 24))
        yb.backward(gout)
        assert rel_err(yb.detach().cpu().numpy(), ya.detach().cpu().numpy()) < 1e-6
        assert rel_err(xb.grad.cpu().numpy(), xa.grad.cpu().numpy()) < 2e-4
        for (k, p), (_, q) in zip(list(bn_b.named_parameters()) + list(layer_b.named_parameters()),
                                  list(bn_a.named_parameters()) + list(layer_a.named_parameters())):
            assert rel_err(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 2e-4, k
        assert rel_err(bn_b.running_var.cpu().numpy(), bn_a.running_var.cpu().numpy()) < 1e-6
