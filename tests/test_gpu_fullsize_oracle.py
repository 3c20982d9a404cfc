"""The tensor path against the CPU ORACLE at the BASELINE sizes ("next" item 1).

test_gpu_fullsize.py compares the tcgen05 kernels with the repo's own generic kernels at these sizes; this file
removes that hop.  Samples are independent in out, grad_x and grad_offset, so the engine runs the FULL
batch and the oracle (oracle/dcn_oracle.c, pinned to the unmodified reference by tests/test_oracle_golden.py)
re-computes a handful of samples of that very batch:

    b = 0, 1, B/2, B-1 and the images that hold the first tile a persistent CTA wraps to
    (tile index = number of CTAs of the forward kernel, 148, and of a fused-backward tile chunk, 148 / 3 = 49)

and those slices are compared at north_star's tolerances (forward 1e-4, gradients 1e-3; max-abs error over max-abs
value).  grad_weight / grad_bias are sums over the batch: they are checked on an engine run over exactly those
samples (full spatial extent, same tile walk inside an image) against the oracle's sums.

bf16 operand mode (BASELINE configs[3]): the oracle runs in fp32 on the bf16-ROUNDED operands; tolerances 1e-2 /
2e-2 as in test_gpu_parity.py (one bf16 rounding per sample on top of fp32 round-off).
"""
import numpy as np
import pytest
import torch

import jittor_dcn_b200 as dcn
from oracle import dcn_oracle as orc

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: B, C, O, H, W, k, s, p        (BASELINE.json configs at their quoted sizes)
    "cfg2": (256, 64, 64, 128, 128, 3, 1, 1),      # configs[1]
    "cfg3": (64, 256, 256, 28, 28, 3, 1, 1),       # configs[2]
    "c3": (128, 128, 128, 56, 56, 3, 1, 1),        # configs[3] layers
    "c4": (128, 256, 256, 28, 28, 3, 1, 1),
    "c5": (128, 512, 512, 14, 14, 3, 1, 1),
    "det2": (1024, 16, 32, 128, 128, 3, 2, 1),     # configs[4] layers at g = 1
    "det5": (1024, 128, 256, 16, 16, 3, 2, 1),
}


def _samples(B, HW):
    tiles_per_image = (HW + 127) // 128
    picks = {0, 1, B // 2, B - 1, 148 // tiles_per_image, 49 // tiles_per_image, (148 + tiles_per_image - 1) // tiles_per_image}
    return sorted(b for b in picks if 0 <= b < B)


def _data(name, seed, sigma=2.0):
    B, C, O, H, W, k, s, p = CONFIGS[name]
    g = torch.Generator(device="cuda").manual_seed(seed)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    off = torch.randn(B, 2 * k * k, Ho, Wo, device="cuda", generator=g) * sigma
    wt = torch.randn(O, C, k, k, device="cuda", generator=g) * (2.0 / (C * k * k)) ** 0.5
    bias = torch.randn(O, device="cuda", generator=g)
    gout = torch.randn(B, O, Ho, Wo, device="cuda", generator=g)
    return x, off, wt, bias, gout


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _check(name, variant, operand, seed, fwd_tol, grad_tol):
    B, C, O, H, W, k, s, p = CONFIGS[name]
    x, off, wt, bias, gout = _data(name, seed)
    if operand == dcn.OPERAND_BF16:
        x, wt, gout = x.bfloat16(), wt.bfloat16(), gout.bfloat16()
    shp = dcn.make_shape(B, C, O, H, W, k, s, p, variant, operand=operand)
    assert [dcn._lib.path_name(shp, ph) for ph in (0, 1)] == ["umma", "umma"], "this test pins the tcgen05 path"
    Ho, Wo = dcn._lib.output_hw(shp)
    sel = _samples(B, Ho * Wo)
    idx = torch.as_tensor(sel, device="cuda")
    # ---- engine, full batch
    out = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant, operand=operand)
    out_sel = out[idx].cpu().numpy()
    del out
    gx, goff, _, _ = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, operand=operand)
    gx_sel, goff_sel = gx[idx].cpu().numpy(), goff[idx].cpu().numpy()
    del gx, goff
    # ---- engine, the selected samples only: grad_weight / grad_bias are batch sums
    xs, offs, gs = x[idx].contiguous(), off[idx].contiguous(), gout[idx].contiguous()
    _, _, gw_s, gb_s = dcn.dcn_backward(xs, offs, wt, gs, True, k, s, p, variant, operand=operand)
    torch.cuda.synchronize()
    # ---- oracle on the same samples (fp32 arithmetic on the operands as stored)
    osh = orc.make_shape(len(sel), C, O, H, W, k, s, p, variant)
    xn, on, wn, gn = (t.float().cpu().numpy() for t in (xs, offs, wt, gs))
    ref_out = orc.forward(osh, xn, on, wn, bias.cpu().numpy())
    ref_gx, ref_goff, ref_gw, ref_gb = orc.backward(osh, xn, on, wn, gn)
    assert _rel(out_sel, ref_out) < fwd_tol, ("out", sel)
    assert _rel(gx_sel, ref_gx) < grad_tol, ("grad_x", sel)
    assert _rel(goff_sel, ref_goff) < grad_tol, ("grad_offset", sel)
    assert _rel(gw_s.cpu().numpy(), ref_gw) < grad_tol, "grad_weight"
    assert _rel(gb_s.cpu().numpy(), ref_gb) < grad_tol, "grad_bias"
    # per-sample errors as well: one bad image must not hide behind the largest value of another
    for i, b in enumerate(sel):
        assert _rel(out_sel[i], ref_out[i]) < fwd_tol, ("out", b)
        assert _rel(gx_sel[i], ref_gx[i]) < grad_tol, ("grad_x", b)
        assert _rel(goff_sel[i], ref_goff[i]) < grad_tol, ("grad_offset", b)


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "c3", "c4", "det2", "det5"])
def test_tcgen05_path_against_oracle_at_baseline_size(name, variant):
    _check(name, variant, dcn.OPERAND_FP32, seed=11, fwd_tol=1e-4, grad_tol=1e-3)


def test_wide_layer_against_oracle_at_baseline_size():
    """C5 (512 -> 512): two output-channel groups; tensor path for the pixel-row layouts."""
    _check("c5", dcn.VARIANT_JITTOR, dcn.OPERAND_FP32, seed=12, fwd_tol=1e-4, grad_tol=1e-3)


@pytest.mark.parametrize("variant", [dcn.VARIANT_TORCH, dcn.VARIANT_JITTOR])
@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_bf16_operands_against_oracle_at_baseline_size(name, variant):
    B, C, O, H, W, k, s, p = CONFIGS[name]
    shp = dcn.make_shape(B, C, O, H, W, k, s, p, variant, operand=dcn.OPERAND_BF16)
    if [dcn._lib.path_name(shp, ph) for ph in (0, 1)] != ["umma", "umma"]:
        pytest.skip("shape not tiled by the tensor path in this layout (covered by test_gpu_fullsize.py)")
    _check(name, variant, dcn.OPERAND_BF16, seed=13, fwd_tol=1e-2, grad_tol=2e-2)


def test_dcnv1_mode_against_oracle_at_baseline_size():
    _check("cfg2", dcn.VARIANT_DCNV1, dcn.OPERAND_FP32, seed=14, fwd_tol=1e-4, grad_tol=1e-3)
