"""Loads the UNMODIFIED reference classes out of /root/reference/train.py.  TEST INFRA.

``import train`` is impossible in the build container (matplotlib at train.py:6 and jittor at
train.py:297 are not installed), so the ``ClassDef`` nodes are cut out of the file with ``ast``
and executed verbatim in a namespace that holds exactly the names they use
(``torch``, ``tnn``, ``F``, ``math`` — train.py:3,8-11).  Nothing is copied into this repo;
the source is read where it lies.  The reference does not exist on the GPU box:
only oracle/make_golden.py (run here) uses this module.
"""
import ast
import math
import os

import torch
import torch.nn as tnn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("DCN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "train.py"))


def load_reference_classes(names=("TorchDeformConv2d", "TorchEDNetDetection")):
    path = os.path.join(REFERENCE_ROOT, "train.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"torch": torch, "tnn": tnn, "F": F, "math": math}
    wanted = [node for node in tree.body if isinstance(node, ast.ClassDef) and node.name in names]
    assert len(wanted) == len(names), [n.name for n in wanted]
    module = ast.Module(body=wanted, type_ignores=[])
    exec(compile(module, path, "exec"), ns)
    return {n: ns[n] for n in names}


class FixedOffsets(tnn.Module):
    """Stands in for ``offset_conv`` so that reference and engine see identical offsets."""

    def __init__(self, offsets):
        super().__init__()
        self.offsets = offsets

    def forward(self, x):
        return self.offsets


def run_reference_layer(C, O, k, s, p, x, off, weight, bias, gout=None):
    """Forward (+ autograd backward) of the unmodified TorchDeformConv2d with injected offsets.

    Returns dict of numpy arrays: out, and if ``gout`` is given gx, goff, gw, gb.
    """
    cls = load_reference_classes(("TorchDeformConv2d",))["TorchDeformConv2d"]
    m = cls(C, O, k, s, p, bias=bias is not None)
    with torch.no_grad():
        m.weight.copy_(torch.as_tensor(weight))
        if bias is not None:
            m.bias.copy_(torch.as_tensor(bias))
    xt = torch.as_tensor(x).clone().requires_grad_(gout is not None)
    ot = torch.as_tensor(off).clone().requires_grad_(gout is not None)
    m.offset_conv = FixedOffsets(ot)
    out = m(xt)
    res = {"out": out.detach().numpy().copy()}
    if gout is not None:
        params = [xt, ot, m.weight] + ([m.bias] if bias is not None else [])
        grads = torch.autograd.grad(out, params, torch.as_tensor(gout))
        res["gx"], res["goff"], res["gw"] = (g.numpy().copy() for g in grads[:3])
        if bias is not None:
            res["gb"] = grads[3].numpy().copy()
    return res
