"""Torch column layout where gcd(Ho*Wo, C) is not a multiple of 16 (ResNet-50 C5: gcd(196, 512) = 4): the
materialised-sample + plain-GEMM path (csrc/dcn_gemm_path.cu, dcn_path_name == "gemm") against the C oracle (fp32
1e-4 / 1e-3; bf16 operands: oracle on bf16-rounded inputs, 1e-2 / 2e-2 as for the tensor path), against the generic
kernels at a full C5 layer, and through the module's autograd."""
import ctypes

import numpy as np
import pytest
import torch

import jittor_dcn_b200 as dcn
from oracle import dcn_oracle as orc
from tests.util import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # B  C    O    H   W   s   sigma
    (2, 64,  48,  14, 14, 1, 1.5),     # gcd(196, 64) = 4
    (3, 96,  40,  10, 12, 1, 2.0),     # gcd(120, 96) = 24: not a multiple of 16
    (2, 32,  64,  15, 15, 2, 1.0),     # stride 2, Ho*Wo = 64... gcd(64, 32) = 32 -> tensor path; kept as a control
    (1, 128, 96,  9,  11, 1, 3.0),     # gcd(99, 128) = 1, many out-of-bounds taps
    (2, 36,  32,  13, 7,  1, 1.0),     # C = 36: partial channel tile, C % 8 != 0 (fp32 only)
]


def _path(shape, operand=dcn.OPERAND_FP32):
    B, C, O, H, W, s, _ = shape
    shp = dcn.make_shape(B, C, O, H, W, 3, s, 1, dcn.VARIANT_TORCH, operand)
    lib = dcn.load()
    return lib.dcn_path_name(ctypes.byref(shp), 0), lib.dcn_path_name(ctypes.byref(shp), 1)


def _bf16_round(a):
    return torch.as_tensor(a).to(torch.bfloat16).float().numpy()


def _data(case, bf16):
    B, C, O, H, W, s, sigma = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    sh = orc.make_shape(B, C, O, H, W, 3, s, 1, dcn.VARIANT_TORCH)
    Ho, Wo = orc.out_hw(sh)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((B, 18, Ho, Wo)) * sigma).astype(np.float32)
    wt = (rng.standard_normal((O, C, 3, 3)) * (2.0 / (C * 9)) ** 0.5).astype(np.float32)
    bias = rng.standard_normal(O).astype(np.float32)
    gout = rng.standard_normal((B, O, Ho, Wo)).astype(np.float32)
    if bf16:
        x, wt, gout = _bf16_round(x), _bf16_round(wt), _bf16_round(gout)
    return sh, x, off, wt, bias, gout


@pytest.mark.parametrize("case", CASES)
def test_gemm_path_against_the_oracle_fp32(case):
    sh, x, off, wt, bias, gout = _data(case, False)
    s = case[5]
    ref_out = orc.forward(sh, x, off, wt, bias)
    ref = orc.backward(sh, x, off, wt, gout)
    tx, toff, twt, tb, tg = (torch.as_tensor(a).cuda() for a in (x, off, wt, bias, gout))
    out = dcn.dcn_forward(tx, toff, twt, tb, 3, s, 1, dcn.VARIANT_TORCH)
    grads = dcn.dcn_backward(tx, toff, twt, tg, True, 3, s, 1, dcn.VARIANT_TORCH)
    assert rel_err(out.cpu().numpy(), ref_out) < 1e-4, _path(case)
    for got, exp, nm in zip(grads, ref, ("gx", "goff", "gw", "gb")):
        assert rel_err(got.cpu().numpy(), exp) < 1e-3, (nm, _path(case))
    # accumulate / skip variants of grad_x
    gx0 = torch.ones_like(tx)
    acc = dcn.dcn_backward(tx, toff, twt, tg, True, 3, s, 1, dcn.VARIANT_TORCH, grad_x=gx0)
    assert rel_err((acc[0] - 1).cpu().numpy(), ref[0]) < 1e-3
    nogx = dcn.dcn_backward(tx, toff, twt, tg, True, 3, s, 1, dcn.VARIANT_TORCH, need_grad_x=False)
    assert nogx[0] is None and rel_err(nogx[2].cpu().numpy(), ref[2]) < 1e-3


def test_the_cases_route_where_they_are_meant_to():
    assert _path(CASES[0]) == (b"gemm", b"gemm")
    assert _path(CASES[1]) == (b"gemm", b"gemm")
    assert _path(CASES[2]) == (b"umma", b"umma")
    assert _path(CASES[3]) == (b"gemm", b"gemm")
    assert _path(CASES[4]) == (b"gemm", b"gemm")
    assert _path(CASES[4], dcn.OPERAND_BF16) == (b"simt", b"simt")     # C % 8 != 0: no bf16 staging copy
    assert _path((128, 512, 512, 14, 14, 1, 0), dcn.OPERAND_BF16) == (b"gemm", b"gemm")


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3]])
def test_gemm_path_bf16_operands(case):
    sh, x, off, wt, bias, gout = _data(case, True)
    s = case[5]
    ref_out = orc.forward(sh, x, off, wt, bias)
    ref = orc.backward(sh, x, off, wt, gout)
    bf = torch.bfloat16
    tx, twt, tg = (torch.as_tensor(a).cuda().to(bf) for a in (x, wt, gout))
    toff, tb = torch.as_tensor(off).cuda(), torch.as_tensor(bias).cuda()
    assert _path(case, dcn.OPERAND_BF16) == (b"gemm", b"gemm")
    out = dcn.dcn_forward(tx, toff, twt, tb, 3, s, 1, dcn.VARIANT_TORCH, operand=dcn.OPERAND_BF16)
    grads = dcn.dcn_backward(tx, toff, twt, tg, True, 3, s, 1, dcn.VARIANT_TORCH, operand=dcn.OPERAND_BF16)
    assert rel_err(out.cpu().numpy(), ref_out) < 1e-2
    for got, exp, nm in zip(grads, ref, ("gx", "goff", "gw", "gb")):
        assert rel_err(got.cpu().numpy(), exp) < 2e-2, nm


@pytest.mark.parametrize("operand", [dcn.OPERAND_FP32, dcn.OPERAND_BF16])
def test_full_c5_layer_against_the_generic_kernels(operand):
    """ResNet-50 C5 (512 -> 512 @ 14x14, BASELINE configs[3]) at batch 4 against the oracle-verified generic kernels."""
    B, C, O, H, W = 4, 512, 512, 14, 14
    g = torch.Generator(device="cuda").manual_seed(9)
    act = torch.bfloat16 if operand == dcn.OPERAND_BF16 else torch.float32
    x = torch.randn(B, C, H, W, device="cuda", generator=g).to(act)
    off = torch.randn(B, 18, H, W, device="cuda", generator=g) * 2
    wt = (torch.randn(O, C, 3, 3, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5).to(act)
    bias = torch.randn(O, device="cuda", generator=g)
    gout = torch.randn(B, O, H, W, device="cuda", generator=g).to(act)
    a = dcn.dcn_forward(x, off, wt, bias, 3, 1, 1, dcn.VARIANT_TORCH, operand=operand)
    b = dcn.dcn_forward(x, off, wt, bias, 3, 1, 1, dcn.VARIANT_TORCH, operand=operand, flags=dcn.FLAG_FORCE_SIMT)
    tol_f, tol_g = (1e-2, 2e-2) if operand == dcn.OPERAND_BF16 else (1e-4, 1e-3)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < tol_f
    ga = dcn.dcn_backward(x, off, wt, gout, True, 3, 1, 1, dcn.VARIANT_TORCH, operand=operand)
    gb = dcn.dcn_backward(x, off, wt, gout, True, 3, 1, 1, dcn.VARIANT_TORCH, operand=operand, flags=dcn.FLAG_FORCE_SIMT)
    for u, v, nm in zip(ga, gb, ("gx", "goff", "gw", "gb")):
        assert rel_err(u.cpu().numpy(), v.cpu().numpy()) < tol_g, nm


def test_module_autograd_on_the_gemm_path():
    torch.manual_seed(2)
    m = dcn.TorchDeformConv2d(64, 48, 3, 1, 1).cuda()
    with torch.no_grad():
        m.offset_conv.weight.normal_(0, 0.02)
        m.offset_conv.bias.normal_(0, 1.0)
    x = torch.randn(2, 64, 14, 14, device="cuda", requires_grad=True)
    out = m(x)
    out.square().mean().backward()
    assert out.shape == (2, 48, 14, 14) and torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
    for p in m.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_backward_reuses_the_forward_staging():
    """DCN_FLAG_XT_STAGED on the gemm path: staged x, the sampling plan and the materialised samples of the forward pass
    are reused by the backward pass (one scratch buffer for both phases)."""
    from jittor_dcn_b200.functional import staged_workspace
    sh, x, off, wt, bias, gout = _data(CASES[0], False)
    tx, toff, twt, tb, tg = (torch.as_tensor(a).cuda() for a in (x, off, wt, bias, gout))
    ws = staged_workspace(tx, twt, 3, 1, 1, dcn.VARIANT_TORCH)
    assert ws is not None
    out = dcn.dcn_forward(tx, toff, twt, tb, 3, 1, 1, dcn.VARIANT_TORCH, ws=ws)
    staged = dcn.dcn_backward(tx, toff, twt, tg, True, 3, 1, 1, dcn.VARIANT_TORCH, ws=ws, xt_staged=True)
    fresh = dcn.dcn_backward(tx, toff, twt, tg, True, 3, 1, 1, dcn.VARIANT_TORCH)
    for a, b in zip(staged, fresh):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
