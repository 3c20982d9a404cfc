"""DeformRoIPool / DeformPSRoIPool on the engine against the CPU restatement of the reference's op
chain (oracle/roi_pool_chain.py — parity unpinned: the reference classes are Jittor modules), forward values and the
autograd of features and offsets, including rois that leave the feature map (clamped corners, weights outside [0, 1])."""
import pytest
import torch

import jittor_dcn_b200 as dcn
from oracle import roi_pool_chain as chain
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _case(seed, B, C, H, W, R, scale):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, C, H, W, generator=g)
    xy = torch.rand(R, 2, generator=g) * torch.tensor([W, H]) / scale * 1.2 - 0.1 * W / scale   # some start outside
    wh = torch.rand(R, 2, generator=g) * torch.tensor([W, H]) / scale * 0.5
    wh[0] = 0.0                                                       # degenerate roi: the 1e-6 floor (:98-99)
    rois = torch.cat([torch.randint(0, B, (R, 1), generator=g).float(), xy, xy + wh], 1)
    return feats, rois, g


@pytest.mark.parametrize("shape", [(2, 8, 12, 16, 9, 1.0), (3, 300, 14, 14, 33, 0.25), (1, 64, 50, 38, 17, 0.5)])
def test_deform_roi_pool_matches_the_reference_chain(shape):
    B, C, H, W, R, scale = shape
    feats, rois, g = _case(1, B, C, H, W, R, scale)
    offsets = torch.randn(R, 1, 2, generator=g) * 0.3
    gout = torch.randn(R, C, 1, 1, generator=g)
    fr, orf = feats.clone().requires_grad_(True), offsets.clone().requires_grad_(True)
    ref = chain.deform_roi_pool(fr, rois, orf, scale)
    ref.backward(gout)
    fd, od = feats.cuda().requires_grad_(True), offsets.cuda().requires_grad_(True)
    m = dcn.DeformRoIPool(1, spatial_scale=scale)
    out = m.execute(fd, rois.cuda(), od)
    out.backward(gout.cuda())
    assert out.shape == (R, C, 1, 1)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-6
    assert rel_err(fd.grad.cpu().numpy(), fr.grad.numpy()) < 1e-5
    assert rel_err(od.grad.cpu().numpy(), orf.grad.numpy()) < 1e-4


@pytest.mark.parametrize("no_trans", [False, True])
@pytest.mark.parametrize("shape", [(2, 8, 12, 16, 9, 1.0), (2, 256, 28, 28, 40, 0.125)])
def test_deform_psroi_pool_matches_the_reference_chain(shape, no_trans):
    B, C, H, W, R, scale = shape
    feats, rois, g = _case(2, B, C, H, W, R, scale)
    offsets = torch.randn(R, 2, generator=g)
    gout = torch.randn(R, C, 1, 1, generator=g)
    fr, orf = feats.clone().requires_grad_(True), offsets.clone().requires_grad_(True)
    ref = chain.deform_psroi_pool(fr, rois, orf, scale, no_trans=no_trans, trans_std=0.1)
    ref.backward(gout)
    fd, od = feats.cuda().requires_grad_(True), offsets.cuda().requires_grad_(True)
    m = dcn.DeformPSRoIPool(1, spatial_scale=scale, no_trans=no_trans, trans_std=0.1)
    out = m(fd, rois.cuda(), od)
    out.backward(gout.cuda())
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-6
    assert rel_err(fd.grad.cpu().numpy(), fr.grad.numpy()) < 1e-5
    if no_trans:
        assert od.grad is None or float(od.grad.abs().max()) == 0.0
    else:
        assert rel_err(od.grad.cpu().numpy(), orf.grad.numpy()) < 1e-4


def test_only_one_bin_is_defined():
    for cls in (dcn.DeformRoIPool, dcn.DeformPSRoIPool):
        with pytest.raises(ValueError):
            cls(7)
        with pytest.raises(ValueError):
            cls((1, 2))
        cls((1, 1))
