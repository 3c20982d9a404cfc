"""Shared helpers for the tests (CPU side)."""
import glob
import os

import numpy as np

from oracle import dcn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def shape_from_cfg(cfg, variant):
    B, C, O, H, W, kh, kw, sh, sw, ph, pw = (int(v) for v in cfg)
    return orc.make_shape(B, C, O, H, W, (kh, kw), (sh, sw), (ph, pw), variant)


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative error' of north_star's tolerances)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def stencil_from_corners(y0, x0, w4, H, W):
    """Dense [B,N,Ho,Wo,H*W] stencil: each in-bounds corner's weight at its pixel."""
    B, N, Ho, Wo = y0.shape
    st = np.zeros((B, N, Ho, Wo, H * W), np.float32)
    idx = np.indices((B, N, Ho, Wo))
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy = y0.astype(np.int64) + dy
        xx = x0.astype(np.int64) + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        pix = np.where(ok, yy * W + xx, 0)
        sel = tuple(i[ok] for i in idx) + (pix[ok],)
        st[sel] = w4[..., k][ok]
    return st
