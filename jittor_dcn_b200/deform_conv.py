"""`DeformConv2d` — drop-in for the Jittor class in the reference's deform_conv.py:6-81.

The reference imports it as ``from deform_conv import DeformConv2d`` (train.py:299,
test.py:12): putting this package directory on ``sys.path`` ahead of the reference's own
file makes that import resolve here.  Jittor is pinned to the CPU in the reference
(``jt.flags.use_cuda = 0``, train.py:301), so Vars are staged host -> pinned -> B200 and back;
the arithmetic is the same C-ABI engine the PyTorch module uses (variant DCN_VARIANT_JITTOR).

Jittor cannot be installed in the build container (no network), so this module is
import-guarded and only syntax-checked there.
"""
import math

import numpy as np

from . import _lib

try:  # pragma: no cover - jittor is absent in the build container
    import jittor as jt
    from jittor import nn
    HAVE_JITTOR = True
except ImportError:  # keep the module importable for documentation / syntax checks
    jt = None
    nn = None
    HAVE_JITTOR = False


def _engine_forward_host(x, offset, weight, bias, k, s, p):
    import torch
    from .functional import dcn_forward
    dev = torch.device("cuda", torch.cuda.current_device())
    t = [None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory().to(dev, non_blocking=True)
         for a in (x, offset, weight, bias)]
    out = dcn_forward(t[0], t[1], t[2], t[3], k, s, p, _lib.VARIANT_JITTOR)
    return out.cpu().numpy()


def _engine_backward_host(x, offset, weight, gout, has_bias, k, s, p):
    import torch
    from .functional import dcn_backward
    dev = torch.device("cuda", torch.cuda.current_device())
    t = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory().to(dev, non_blocking=True)
         for a in (x, offset, weight, gout)]
    gx, goff, gw, gb = dcn_backward(t[0], t[1], t[2], t[3], has_bias, k, s, p, _lib.VARIANT_JITTOR)
    return tuple(None if g is None else g.cpu().numpy() for g in (gx, goff, gw, gb))


if HAVE_JITTOR:  # pragma: no cover

    class _DeformConvCore(jt.Function):
        """Stands where the reference has nn.grid_sample + jt.matmul (deform_conv.py:62-81)."""

        def execute(self, x, offset, weight, bias, k, s, p):
            self.saved = (x.data, offset.data, weight.data, bias is not None, k, s, p)
            b = None if bias is None else bias.data
            return jt.array(_engine_forward_host(x.data, offset.data, weight.data, b, k, s, p))

        def grad(self, grad_out):
            x, offset, weight, has_bias, k, s, p = self.saved
            gx, goff, gw, gb = _engine_backward_host(x, offset, weight, grad_out.data, has_bias, k, s, p)
            return (jt.array(gx), jt.array(goff), jt.array(gw),
                    jt.array(gb) if has_bias else None, None, None, None)

    class DeformConv2d(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=True):
            super().__init__()
            self.in_channels = in_channels
            self.out_channels = out_channels
            self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
            self.stride = stride if isinstance(stride, tuple) else (stride, stride)
            self.padding = padding if isinstance(padding, tuple) else (padding, padding)
            self.N = self.kernel_size[0] * self.kernel_size[1]
            # companion offset conv (deform_conv.py:16-21), zero-initialised (:27-28)
            self.offset_conv = nn.Conv(in_channels, 2 * self.N, kernel_size=self.kernel_size,
                                       stride=self.stride, padding=self.padding)
            std = math.sqrt(2.0 / (in_channels * self.kernel_size[0] * self.kernel_size[1]))
            self.weight = jt.init.gauss([out_channels, in_channels, *self.kernel_size], mean=0.0, std=std)
            self.bias = jt.init.constant(shape=[out_channels], value=0.0) if bias else None
            self.offset_conv.weight = jt.zeros_like(self.offset_conv.weight)
            self.offset_conv.bias = jt.zeros_like(self.offset_conv.bias)

        def execute(self, x):
            offset = self.offset_conv(x)
            return _DeformConvCore()(x, offset, self.weight, self.bias, self.kernel_size,
                                     self.stride, self.padding)

else:

    class DeformConv2d:  # noqa: D401 - placeholder that fails loudly
        """Placeholder: constructing it without jittor installed raises."""

        def __init__(self, *a, **kw):
            raise ImportError("jittor is not installed: use jittor_dcn_b200.TorchDeformConv2d "
                              "(or TorchDeformConv2dJittorSemantics for the Jittor operator's maths)")
