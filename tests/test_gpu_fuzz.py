"""Shape fuzz: random problems that route to the tensor path, checked against the generic
kernels (which the fixed-shape tests pin to the oracle and the reference goldens).  Exercises the
tile edge cases: partial row tiles, instance counts that do not fill a tile, K not a multiple of
64 / 128, channel permutations (Cs > 1), 16-channel groups, 1x1 and rectangular taps, stride 2,
fused and unfused backward, both layouts, both operand modes."""
import ctypes

import numpy as np
import pytest
import torch

import jittor_dcn_b200 as dcn

pytestmark = pytest.mark.gpu


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        C = int(rng.choice([16, 32, 64, 128, 192, 256]))
        O = int(rng.choice([16, 32, 48, 64, 80, 128, 144, 256]))
        k = [(3, 3), (1, 1), (3, 1), (1, 3), (5, 5)][int(rng.integers(0, 5))]
        s = int(rng.choice([1, 2]))
        p = (k[0] // 2, k[1] // 2)
        if rng.random() < 0.6:   # extents with Ho*Wo % 16 == 0 so that the Torch layout tiles (gcd(Ho*Wo, C) % 16)
            H, W = 4 * s * int(rng.integers(2, 9)), 4 * s * int(rng.integers(2, 9))
        else:
            H, W = int(rng.integers(6, 40)), int(rng.integers(6, 40))
        B = int(rng.integers(1, 4))
        variant = int(rng.integers(0, 2))
        out.append((B, C, O, H, W, k, s, p, variant))
    return out


def _rel(a, b):
    den = float(b.double().abs().max())
    return float((a.double() - b.double()).abs().max() / (den if den > 0 else 1.0))


@pytest.mark.parametrize("case", _cases(2026, 90))
def test_fuzz_tensor_path_vs_generic(case):
    B, C, O, H, W, k, s, p, variant = case
    lib = dcn.load()
    shp = dcn.make_shape(B, C, O, H, W, k, s, p, variant)
    Ho, Wo = (H + 2 * p[0] - k[0]) // s + 1, (W + 2 * p[1] - k[1]) // s + 1
    if variant == dcn.VARIANT_JITTOR and min(Ho, Wo) < 2:
        pytest.skip("output extent 1: the reference divides by zero")
    paths = (lib.dcn_path_name(ctypes.byref(shp), 0), lib.dcn_path_name(ctypes.byref(shp), 1))
    if paths == (b"simt", b"simt"):
        pytest.skip("shape does not tile onto the tensor path")
    g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xffff)
    N = k[0] * k[1]
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    off = torch.randn(B, 2 * N, Ho, Wo, device="cuda", generator=g) * 2.0
    wt = torch.randn(O, C, *k, device="cuda", generator=g) * (2.0 / (C * N)) ** 0.5
    bias = torch.randn(O, device="cuda", generator=g)
    gout = torch.randn(B, O, Ho, Wo, device="cuda", generator=g)
    out_u = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant)
    out_s = dcn.dcn_forward(x, off, wt, bias, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    assert _rel(out_u, out_s) < 1e-4, ("fwd", paths)
    gu = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant)
    gs = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, flags=dcn.FLAG_FORCE_SIMT)
    for a, b, nm in zip(gu, gs, ("gx", "goff", "gw", "gb")):
        assert _rel(a, b) < 1e-3, (nm, paths)


@pytest.mark.parametrize("case", _cases(7, 30))
def test_fuzz_bf16_vs_fp32(case):
    B, C, O, H, W, k, s, p, variant = case
    lib = dcn.load()
    shp = dcn.make_shape(B, C, O, H, W, k, s, p, variant, operand=dcn.OPERAND_BF16)
    Ho, Wo = (H + 2 * p[0] - k[0]) // s + 1, (W + 2 * p[1] - k[1]) // s + 1
    if variant == dcn.VARIANT_JITTOR and min(Ho, Wo) < 2:
        pytest.skip("output extent 1")
    if lib.dcn_path_name(ctypes.byref(shp), 0) != b"umma" or lib.dcn_path_name(ctypes.byref(shp), 1) != b"umma":
        pytest.skip("bf16 needs the tensor path for forward and backward")
    g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xffff)
    N = k[0] * k[1]
    x = torch.randn(B, C, H, W, device="cuda", generator=g).bfloat16()
    off = torch.randn(B, 2 * N, Ho, Wo, device="cuda", generator=g) * 1.5
    wt = (torch.randn(O, C, *k, device="cuda", generator=g) * (2.0 / (C * N)) ** 0.5).bfloat16()
    gout = torch.randn(B, O, Ho, Wo, device="cuda", generator=g).bfloat16()
    out_b = dcn.dcn_forward(x, off, wt, None, k, s, p, variant, operand=dcn.OPERAND_BF16)
    out_f = dcn.dcn_forward(x.float(), off, wt.float(), None, k, s, p, variant)
    assert _rel(out_b, out_f) < 1e-2
    gb_ = dcn.dcn_backward(x, off, wt, gout, True, k, s, p, variant, operand=dcn.OPERAND_BF16)
    gf_ = dcn.dcn_backward(x.float(), off, wt.float(), gout.float(), True, k, s, p, variant)
    for a, b, nm in zip(gb_, gf_, ("gx", "goff", "gw", "gb")):
        assert _rel(a, b) < 2e-2, nm
