"""ctypes front-end of oracle/dcn_oracle.c (the plain-C restatement).  TEST INFRASTRUCTURE.

Every function takes/returns numpy float32 arrays in the reference's layouts:
``x[B,C,H,W]``, ``off[B,2N,Ho,Wo]``, ``weight[O,C,kh,kw]``, ``bias[O]``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdcn_oracle.so")

VARIANT_JITTOR = 0  # deform_conv.py:56-81
VARIANT_TORCH = 1   # train.py:95-140
VARIANT_DCNV1 = 2   # torchvision.ops.deform_conv2d semantics


class DcnShape(ctypes.Structure):
    """Mirror of include/dcn_b200.h:DcnShape."""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "C", "O", "H", "W", "kh", "kw", "sh", "sw", "ph", "pw",
                 "variant", "operand", "flags")]


def build(force=False):
    src = os.path.join(_HERE, "dcn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        for name in ("dcn_oracle_corners", "dcn_oracle_sample", "dcn_oracle_forward",
                     "dcn_oracle_backward", "dcn_oracle_threads"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def make_shape(B, C, O, H, W, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH):
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    return DcnShape(B, C, O, H, W, kh, kw, sh, sw, ph, pw, variant, 0, 0)


def out_hw(s):
    return (s.H + 2 * s.ph - s.kh) // s.sh + 1, (s.W + 2 * s.pw - s.kw) // s.sw + 1


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def corners(s, off):
    """-> y0[B,N,Ho,Wo] i32, x0 i32, w4[B,N,Ho,Wo,4] f32, fxy[B,N,Ho,Wo,2] f32"""
    Ho, Wo = out_hw(s)
    N = s.kh * s.kw
    off, poff = _f32(off)
    assert off.shape == (s.B, 2 * N, Ho, Wo), off.shape
    y0 = np.empty((s.B, N, Ho, Wo), np.int32)
    x0 = np.empty_like(y0)
    w4 = np.empty((s.B, N, Ho, Wo, 4), np.float32)
    fxy = np.empty((s.B, N, Ho, Wo, 2), np.float32)
    rc = lib().dcn_oracle_corners(ctypes.byref(s), poff, y0.ctypes.data_as(ctypes.c_void_p),
                                  x0.ctypes.data_as(ctypes.c_void_p),
                                  w4.ctypes.data_as(ctypes.c_void_p),
                                  fxy.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return y0, x0, w4, fxy


def sample(s, x, off):
    """-> S[B,C,Ho,Wo,N], the tensor the reference materialises at train.py:129."""
    Ho, Wo = out_hw(s)
    N = s.kh * s.kw
    x, px = _f32(x)
    off, poff = _f32(off)
    S = np.empty((s.B, s.C, Ho, Wo, N), np.float32)
    rc = lib().dcn_oracle_sample(ctypes.byref(s), px, poff, S.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return S


def forward(s, x, off, weight, bias=None):
    Ho, Wo = out_hw(s)
    x, px = _f32(x)
    off, poff = _f32(off)
    weight, pw = _f32(weight)
    pb = None
    if bias is not None:
        bias, pb = _f32(bias)
    out = np.empty((s.B, s.O, Ho, Wo), np.float32)
    rc = lib().dcn_oracle_forward(ctypes.byref(s), px, poff, pw, pb,
                                  out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return out


def backward(s, x, off, weight, gout):
    """-> grad_x, grad_offset, grad_weight, grad_bias"""
    Ho, Wo = out_hw(s)
    N = s.kh * s.kw
    x, px = _f32(x)
    off, poff = _f32(off)
    weight, pw = _f32(weight)
    gout, pg = _f32(gout)
    assert gout.shape == (s.B, s.O, Ho, Wo)
    gx = np.empty((s.B, s.C, s.H, s.W), np.float32)
    goff = np.empty((s.B, 2 * N, Ho, Wo), np.float32)
    gw = np.empty((s.O, s.C, s.kh, s.kw), np.float32)
    gb = np.empty((s.O,), np.float32)
    rc = lib().dcn_oracle_backward(ctypes.byref(s), px, poff, pw, pg,
                                   gx.ctypes.data_as(ctypes.c_void_p),
                                   goff.ctypes.data_as(ctypes.c_void_p),
                                   gw.ctypes.data_as(ctypes.c_void_p),
                                   gb.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return gx, goff, gw, gb
