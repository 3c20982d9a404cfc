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
