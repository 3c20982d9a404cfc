"""Tensor-level entry points: torch tensors in, raw device pointers across the C ABI.

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic of
the operator happens in libdcn_b200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import (FLAG_ACCUM_GRAD_X, FLAG_FORCE_SIMT, FLAG_NO_GRAD_X, FLAG_RELU_OUT, FLAG_XT_STAGED, OPERAND_BF16, OPERAND_FP32,
                   PHASE_BACKWARD, PHASE_FORWARD, PHASE_LAYER_BACKWARD, PHASE_LAYER_FORWARD, VARIANT_DCNV1,
                   VARIANT_JITTOR, VARIANT_TORCH)

_workspaces = {}
_captured = []   # scratch buffers whose addresses are baked into CUDA graphs: alive until clear_workspaces()


def _workspace(device, stream_id, nbytes):
    """Caller-owned scratch, one growing buffer per (device, stream).

    While the stream is being captured into a CUDA graph the call gets a PRIVATE buffer that is never freed or
    handed out again (a replay reads and writes the captured address; the shared cache may be re-allocated by a
    later, larger eager call).  clear_workspaces() drops everything — only once such graphs are gone."""
    if torch.cuda.is_current_stream_capturing():
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _captured.append(ws)
        return ws
    key = (device.index, stream_id)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def clear_workspaces():
    """Release the cached scratch buffers (all devices / streams) and the ones captured CUDA graphs point at.
    Call only when no such graph will be replayed again."""
    _workspaces.clear()
    del _captured[:]


def _dev_ready(t, dtype=torch.float32):
    """Dense, 16-byte aligned tensor of the wanted dtype on its current CUDA device."""
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _compute_device(x):
    if x.is_cuda:
        return x.device
    if not torch.cuda.is_available():
        raise _lib.DcnError("jittor_dcn_b200 needs a CUDA device (B200, sm_100a): there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(t, dev):
    """Host tensors are staged through pinned memory (the reference feeds CPU tensors,
    train.py:239); CUDA tensors are used in place."""
    if t is None or t.device == dev:
        return t
    if t.device.type == "cpu":
        t = t.detach().contiguous()
        try:
            t = t.pin_memory()
        except RuntimeError:
            pass
        return t.to(dev, non_blocking=True)
    return t.to(dev)


def _shape_of(x, weight, kernel_size, stride, padding, variant, operand, flags):
    B, C, H, W = x.shape
    O = weight.shape[0]
    return _lib.make_shape(B, C, O, H, W, kernel_size, stride, padding, variant, operand, flags)


def staged_workspace(x, weight, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH,
                     operand=OPERAND_FP32, flags=0):
    """A private scratch buffer big enough for BOTH phases of one layer call, or None when either
    phase would not run on the tensor path.  Passing it to dcn_forward and then to dcn_backward
    (xt_staged=True) lets the backward pass reuse the staged copy of x (DCN_FLAG_XT_STAGED)."""
    lib = _lib.load()
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, operand, flags)
    names = [_lib.path_name(shp, ph) for ph in (PHASE_FORWARD, PHASE_BACKWARD)]
    if names not in (["umma", "umma"], ["gemm", "gemm"]):      # both keep the forward's staging at the workspace head
        return None
    need = max(lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_FORWARD),
               lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_BACKWARD))
    return torch.empty(need, dtype=torch.uint8, device=x.device)


def dcn_forward(x, offset, weight, bias, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH,
                operand=OPERAND_FP32, flags=0, ws=None):
    """out[B,O,Ho,Wo] = engine forward on CUDA tensors (no autograd)."""
    lib = _lib.load()
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, operand, flags)
    Ho, Wo = _lib.output_hw(shp)
    N = shp.kh * shp.kw
    if tuple(offset.shape) != (shp.B, 2 * N, Ho, Wo):
        raise ValueError(f"offset shape {tuple(offset.shape)} != {(shp.B, 2 * N, Ho, Wo)}")
    if tuple(weight.shape) != (shp.O, shp.C, shp.kh, shp.kw):
        raise ValueError(f"weight shape {tuple(weight.shape)} != {(shp.O, shp.C, shp.kh, shp.kw)}")
    act = torch.bfloat16 if operand == OPERAND_BF16 else torch.float32
    x, weight = _dev_ready(x, act), _dev_ready(weight, act)
    offset, bias = _dev_ready(offset), _dev_ready(bias)
    out = torch.empty((shp.B, shp.O, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device)
        need = lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_FORWARD)
        if ws is None:
            ws = _workspace(x.device, stream.cuda_stream, need)
        rc = lib.dcn_forward(ctypes.byref(shp), _ptr(x), _ptr(offset), _ptr(weight), _ptr(bias),
                             _ptr(out), _ptr(ws), ws.numel(), ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_forward")
    return out


def dcn_backward(x, offset, weight, grad_out, has_bias, kernel_size=3, stride=1, padding=1,
                 variant=VARIANT_TORCH, operand=OPERAND_FP32, flags=0, need_grad_x=True, ws=None,
                 xt_staged=False, grad_x=None):
    """-> grad_x (or None), grad_offset, grad_weight, grad_bias (or None); all float32.
    ws / xt_staged: the buffer from staged_workspace() that the matching dcn_forward used.
    grad_x: an existing float32 gradient of x's shape to ACCUMULATE into (DCN_FLAG_ACCUM_GRAD_X; e.g. the
    term the companion offset conv contributes); without it the flag is refused — the library would add into
    whatever the freshly allocated output buffer happens to hold."""
    lib = _lib.load()
    if grad_x is not None:
        if not need_grad_x:
            raise ValueError("grad_x given but need_grad_x is False")
        if grad_x.dtype != torch.float32 or tuple(grad_x.shape) != tuple(x.shape) or not grad_x.is_contiguous() \
                or grad_x.data_ptr() % 16:
            raise ValueError("grad_x must be a dense, 16-byte aligned float32 tensor of x's shape")
        flags |= FLAG_ACCUM_GRAD_X
    elif flags & FLAG_ACCUM_GRAD_X:
        raise ValueError("FLAG_ACCUM_GRAD_X needs the tensor to accumulate into: pass grad_x=")
    if not need_grad_x:
        flags |= FLAG_NO_GRAD_X
    if xt_staged and ws is not None:
        flags |= FLAG_XT_STAGED
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, operand, flags)
    Ho, Wo = _lib.output_hw(shp)
    N = shp.kh * shp.kw
    act = torch.bfloat16 if operand == OPERAND_BF16 else torch.float32
    x, weight, grad_out = _dev_ready(x, act), _dev_ready(weight, act), _dev_ready(grad_out, act)
    offset = _dev_ready(offset)
    dev = x.device
    gx = grad_x if grad_x is not None else (torch.empty(x.shape, dtype=torch.float32, device=dev) if need_grad_x else None)
    goff = torch.empty((shp.B, 2 * N, Ho, Wo), dtype=torch.float32, device=dev)
    gw = torch.empty(weight.shape, dtype=torch.float32, device=dev)
    gb = torch.empty((shp.O,), dtype=torch.float32, device=dev) if has_bias else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        need = lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_BACKWARD)
        if ws is None:
            ws = _workspace(dev, stream.cuda_stream, need)
        rc = lib.dcn_backward(ctypes.byref(shp), _ptr(x), _ptr(offset), _ptr(weight), _ptr(grad_out),
                              _ptr(gx), _ptr(goff), _ptr(gw), _ptr(gb), _ptr(ws), ws.numel(),
                              ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_backward")
    return gx, goff, gw, gb


# ---- whole layer: companion offset conv + DCN span on the engine -----------------------------------
def layer_supported(x_shape, out_channels, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH,
                    operand=OPERAND_FP32, flags=0):
    """True when dcn_layer_forward / dcn_layer_backward cover the shape (both phases on the tensor path)."""
    B, C, H, W = (int(v) for v in x_shape)
    shp = _lib.make_shape(B, C, int(out_channels), H, W, kernel_size, stride, padding, variant, operand, flags)
    return all(_lib.path_name(shp, ph) == "umma" for ph in (PHASE_LAYER_FORWARD, PHASE_LAYER_BACKWARD))


def layer_workspace(x, weight, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH, flags=0):
    """One scratch buffer big enough for dcn_layer_forward AND dcn_layer_backward of this layer: handing the same
    buffer to both lets the backward pass reuse the staged copy of x (DCN_FLAG_XT_STAGED)."""
    lib = _lib.load()
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, OPERAND_FP32, flags)
    need = max(lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_FORWARD),
               lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_BACKWARD))
    return torch.empty(need, dtype=torch.uint8, device=x.device)


def dcn_offset_conv_forward(x, offset_weight, offset_bias, out_channels, kernel_size=3, stride=1, padding=1,
                            variant=VARIANT_TORCH, flags=0, ws=None, xt_staged=False):
    """offset[B,2N,Ho,Wo] = the companion offset convolution on the engine (a plain mode of the tcgen05 forward
    kernel).  `variant` / out_channels describe the DCN layer the offsets are for: the staged copy of x is laid out
    for it, so a following dcn_forward(..., ws=ws, xt_staged=True) reuses it."""
    lib = _lib.load()
    B, C, H, W = x.shape
    if xt_staged and ws is not None:
        flags |= FLAG_XT_STAGED
    shp = _lib.make_shape(B, C, int(out_channels), H, W, kernel_size, stride, padding, variant, OPERAND_FP32, flags)
    Ho, Wo = _lib.output_hw(shp)
    N = shp.kh * shp.kw
    x, offset_weight, offset_bias = _dev_ready(x), _dev_ready(offset_weight), _dev_ready(offset_bias)
    if tuple(offset_weight.shape) != (2 * N, C, shp.kh, shp.kw):
        raise ValueError(f"offset_weight shape {tuple(offset_weight.shape)} != {(2 * N, C, shp.kh, shp.kw)}")
    offset = torch.empty((B, 2 * N, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device)
        need = lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_FORWARD)
        if ws is None:
            ws = _workspace(x.device, stream.cuda_stream, need)
        rc = lib.dcn_offset_conv_forward(ctypes.byref(shp), _ptr(x), _ptr(offset_weight), _ptr(offset_bias), _ptr(offset),
                                         _ptr(ws), ws.numel(), ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_offset_conv_forward")
    return offset


def dcn_layer_forward(x, offset_weight, offset_bias, weight, bias, kernel_size=3, stride=1, padding=1,
                      variant=VARIANT_TORCH, flags=0, ws=None):
    """-> (offset, out): offset conv + DCN forward, x staged once (no autograd)."""
    lib = _lib.load()
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, OPERAND_FP32, flags)
    Ho, Wo = _lib.output_hw(shp)
    N = shp.kh * shp.kw
    x, weight, bias = _dev_ready(x), _dev_ready(weight), _dev_ready(bias)
    offset_weight, offset_bias = _dev_ready(offset_weight), _dev_ready(offset_bias)
    offset = torch.empty((shp.B, 2 * N, Ho, Wo), dtype=torch.float32, device=x.device)
    out = torch.empty((shp.B, shp.O, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device)
        need = lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_FORWARD)
        if ws is None:
            ws = _workspace(x.device, stream.cuda_stream, need)
        rc = lib.dcn_layer_forward(ctypes.byref(shp), _ptr(x), _ptr(offset_weight), _ptr(offset_bias), _ptr(weight),
                                   _ptr(bias), _ptr(offset), _ptr(out), _ptr(ws), ws.numel(),
                                   ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_layer_forward")
    return offset, out


def dcn_layer_backward(x, offset, offset_weight, weight, grad_out, has_offset_bias=True, has_bias=True, kernel_size=3,
                       stride=1, padding=1, variant=VARIANT_TORCH, flags=0, need_grad_x=True, ws=None, xt_staged=False):
    """-> grad_x (or None), grad_offset_weight, grad_offset_bias (or None), grad_weight, grad_bias (or None).
    grad_offset stays inside the workspace; the offset conv's data gradient is accumulated on chip-side buffers."""
    lib = _lib.load()
    if not need_grad_x:
        flags |= FLAG_NO_GRAD_X
    if xt_staged and ws is not None:
        flags |= FLAG_XT_STAGED
    shp = _shape_of(x, weight, kernel_size, stride, padding, variant, OPERAND_FP32, flags)
    x, weight, grad_out = _dev_ready(x), _dev_ready(weight), _dev_ready(grad_out)
    offset, offset_weight = _dev_ready(offset), _dev_ready(offset_weight)
    dev = x.device
    gx = torch.empty(x.shape, dtype=torch.float32, device=dev) if need_grad_x else None
    gwoff = torch.empty(offset_weight.shape, dtype=torch.float32, device=dev)
    gboff = torch.empty((offset_weight.shape[0],), dtype=torch.float32, device=dev) if has_offset_bias else None
    gw = torch.empty(weight.shape, dtype=torch.float32, device=dev)
    gb = torch.empty((shp.O,), dtype=torch.float32, device=dev) if has_bias else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        need = lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_BACKWARD)
        if ws is None:
            ws = _workspace(dev, stream.cuda_stream, need)
        rc = lib.dcn_layer_backward(ctypes.byref(shp), _ptr(x), _ptr(offset), _ptr(offset_weight), _ptr(weight),
                                    _ptr(grad_out), _ptr(gx), _ptr(gwoff), _ptr(gboff), _ptr(gw), _ptr(gb), _ptr(ws),
                                    ws.numel(), ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_layer_backward")
    return gx, gwoff, gboff, gw, gb


class DeformLayerFunction(torch.autograd.Function):
    """The whole reference module forward (offset conv + sampling + GEMM, deform_conv.py:56-81 / train.py:95-140)
    and its autograd as ONE node on the engine: x is staged once per step, grad_offset never reaches HBM as a
    framework tensor, and the two data-gradient terms are summed before the single transposition back to NCHW."""

    @staticmethod
    def forward(ctx, x, offset_weight, offset_bias, weight, bias, cfg):
        kernel_size, stride, padding, variant, flags, keep_staged = cfg
        ctx.ws = layer_workspace(x, weight, kernel_size, stride, padding, variant, flags) if keep_staged else None
        offset, out = dcn_layer_forward(x, offset_weight, offset_bias, weight, bias, kernel_size, stride, padding,
                                        variant, flags, ws=ctx.ws)
        # DCN_FLAG_RELU_OUT: the epilogue applied the ReLU; the backward entry points want the gradient of the
        # pre-activation sum, i.e. grad_out masked with out > 0
        ctx.relu = bool(flags & FLAG_RELU_OUT)
        ctx.save_for_backward(x, offset, offset_weight, weight, *((out,) if ctx.relu else ()))
        ctx.cfg = cfg
        ctx.has = (offset_bias is not None, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, offset, offset_weight, weight = ctx.saved_tensors[:4]
        if ctx.relu:
            grad_out = grad_out * (ctx.saved_tensors[4] > 0)
        kernel_size, stride, padding, variant, flags, _ = ctx.cfg
        gx, gwoff, gboff, gw, gb = dcn_layer_backward(
            x, offset, offset_weight, weight, grad_out, ctx.has[0], ctx.has[1], kernel_size, stride, padding, variant,
            flags & ~FLAG_ACCUM_GRAD_X, need_grad_x=ctx.needs_input_grad[0], ws=ctx.ws, xt_staged=ctx.ws is not None)
        ctx.ws = None
        return gx, gwoff, gboff, gw, gb, None


def deform_layer(x, offset_weight, offset_bias, weight, bias=None, kernel_size=3, stride=1, padding=1,
                 variant=VARIANT_TORCH, flags=0, keep_staged=False):
    """Differentiable whole layer on the engine (CUDA float32 tensors; check layer_supported() first)."""
    cfg = (_lib._pair(kernel_size), _lib._pair(stride), _lib._pair(padding), variant, flags, bool(keep_staged))
    return DeformLayerFunction.apply(x, offset_weight, offset_bias, weight, bias, cfg)


def dcn_corners(offset, in_hw, kernel_size=3, stride=1, padding=1, variant=VARIANT_TORCH):
    """Sampling geometry only: y0, x0 [B,N,Ho,Wo] int32 and w4 [B,N,Ho,Wo,4] float32."""
    lib = _lib.load()
    B = offset.shape[0]
    shp = _lib.make_shape(B, 1, 1, in_hw[0], in_hw[1], kernel_size, stride, padding, variant)
    Ho, Wo = _lib.output_hw(shp)
    N = shp.kh * shp.kw
    offset = _dev_ready(offset)
    dev = offset.device
    y0 = torch.empty((B, N, Ho, Wo), dtype=torch.int32, device=dev)
    x0 = torch.empty_like(y0)
    w4 = torch.empty((B, N, Ho, Wo, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        rc = lib.dcn_debug_corners(ctypes.byref(shp), _ptr(offset), _ptr(y0), _ptr(x0), _ptr(w4),
                                   ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcn_debug_corners")
    return y0, x0, w4


class DeformConvFunction(torch.autograd.Function):
    """autograd node standing where the reference has grid_sample + matmul + their autograd
    (train.py:102-140 forward, train.py:249 backward)."""

    @staticmethod
    def forward(ctx, x, offset, weight, bias, cfg):
        kernel_size, stride, padding, variant, operand, flags = cfg[:6]
        keep_staged = len(cfg) > 6 and cfg[6]
        home = x.device
        dev = _compute_device(x)
        xd, od, wd = _to_device(x, dev), _to_device(offset, dev), _to_device(weight, dev)
        bd = _to_device(bias, dev)
        act = torch.bfloat16 if operand == OPERAND_BF16 else torch.float32
        ctx.ws = None
        if keep_staged and xd.dtype == act and xd.is_contiguous():
            # private scratch that lives until backward: its head keeps the staged copy of x
            ctx.ws = staged_workspace(xd, wd, kernel_size, stride, padding, variant, operand, flags)
        out = dcn_forward(xd, od, wd, bd, kernel_size, stride, padding, variant, operand, flags, ws=ctx.ws)
        ctx.relu = bool(flags & FLAG_RELU_OUT)    # see DeformLayerFunction
        ctx.save_for_backward(xd, od, wd, *((out,) if ctx.relu else ()))
        ctx.cfg, ctx.home, ctx.has_bias = cfg, home, bias is not None
        ctx.dtypes = (x.dtype, offset.dtype, weight.dtype)
        return out if home == dev else out.to(home)

    @staticmethod
    def backward(ctx, grad_out):
        xd, od, wd = ctx.saved_tensors[:3]
        kernel_size, stride, padding, variant, operand, flags = ctx.cfg[:6]
        flags &= ~FLAG_ACCUM_GRAD_X     # autograd sums gradient terms itself; the output buffer here is fresh
        dev = xd.device
        grad_out = _to_device(grad_out, dev)
        if ctx.relu:
            grad_out = grad_out * (ctx.saved_tensors[3] > 0)
        gx, goff, gw, gb = dcn_backward(xd, od, wd, grad_out, ctx.has_bias,
                                        kernel_size, stride, padding, variant, operand, flags,
                                        need_grad_x=ctx.needs_input_grad[0], ws=ctx.ws,
                                        xt_staged=ctx.ws is not None)
        ctx.ws = None
        home = ctx.home

        def back(t, dtype):
            if t is None:
                return None
            t = t.to(dtype) if t.dtype != dtype else t
            return t if home == dev else t.to(home)

        return (back(gx, ctx.dtypes[0]), back(goff, ctx.dtypes[1]), back(gw, ctx.dtypes[2]),
                back(gb, torch.float32), None)


def deform_conv2d(x, offset, weight, bias=None, kernel_size=3, stride=1, padding=1,
                  variant=VARIANT_TORCH, operand=OPERAND_FP32, flags=0, keep_staged=False):
    """Differentiable DeformConv2d core: everything after the offset conv.
    keep_staged: hold a private scratch buffer (backward-phase size) from forward to backward so
    that the backward pass reuses the staged copy of x instead of transposing it again."""
    cfg = (_lib._pair(kernel_size), _lib._pair(stride), _lib._pair(padding), variant, operand, flags,
           bool(keep_staged))
    return DeformConvFunction.apply(x, offset, weight, bias, cfg)


def deform_conv2d_v1(input, offset, weight, bias=None, stride=1, padding=0):
    """Standard deformable convolution v1 with the call signature (and semantics) of
    ``torchvision.ops.deform_conv2d(input, offset, weight, bias, stride, padding)`` for one offset
    group, no mask, dilation 1 — the same kernels as the reference's operators with the textbook
    coordinate generator.  Differentiable in input, offset, weight and bias."""
    kernel_size = (int(weight.shape[2]), int(weight.shape[3]))
    return deform_conv2d(input, offset, weight, bias, kernel_size, stride, padding, variant=VARIANT_DCNV1)


# ---- post-op: BatchNorm2d + ReLU (train.py:167-170) -------------------------------------
class BatchNormReLUFunction(torch.autograd.Function):
    """relu(batch_norm(x)) on the engine's own kernels (csrc/dcn_bn.cu).  running_mean / running_var are
    updated in place in training mode, exactly as nn.BatchNorm2d does."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps):
        lib = _lib.load()
        home = x.device
        dev = _compute_device(x)
        xd = _dev_ready(_to_device(x, dev))
        B, C = xd.shape[0], xd.shape[1]
        HW = xd.numel() // (B * C)
        wd, bd = _dev_ready(_to_device(weight, dev)), _dev_ready(_to_device(bias, dev))
        rm, rv = _to_device(running_mean, dev), _to_device(running_var, dev)
        y = torch.empty_like(xd)
        saved = torch.empty(4 * C, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            need = lib.dcn_bn_workspace_bytes(C)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            rc = lib.dcn_bn_relu_forward(B, C, HW, 1 if training else 0, _ptr(xd), _ptr(wd), _ptr(bd), _ptr(rm),
                                         _ptr(rv), float(momentum), float(eps), _ptr(y), _ptr(saved), _ptr(ws),
                                         ws.numel(), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_bn_relu_forward")
        if training and running_mean is not None and rm is not running_mean:
            running_mean.copy_(rm)       # host-resident module: write the updated statistics back
            running_var.copy_(rv)
        ctx.save_for_backward(xd, saved)
        ctx.training, ctx.home = bool(training), home
        ctx.has_affine = (weight is not None, bias is not None)
        return y if home == dev else y.to(home)

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        xd, saved = ctx.saved_tensors
        dev = xd.device
        B, C = xd.shape[0], xd.shape[1]
        HW = xd.numel() // (B * C)
        gy = _dev_ready(_to_device(grad_y, dev))
        gx = torch.empty_like(xd) if ctx.needs_input_grad[0] else None
        gg = torch.empty(C, dtype=torch.float32, device=dev) if ctx.has_affine[0] else None
        gb = torch.empty(C, dtype=torch.float32, device=dev) if ctx.has_affine[1] else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            ws = torch.empty(lib.dcn_bn_workspace_bytes(C), dtype=torch.uint8, device=dev)
            rc = lib.dcn_bn_relu_backward(B, C, HW, 1 if ctx.training else 0, _ptr(xd), _ptr(gy), _ptr(saved),
                                          _ptr(gx), _ptr(gg), _ptr(gb), _ptr(ws), ws.numel(),
                                          ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_bn_relu_backward")
        home = ctx.home

        def back(t):
            return t if t is None or home == dev else t.to(home)

        return back(gx), back(gg), back(gb), None, None, None, None, None


def batch_norm_relu(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    """relu(F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)) for NCHW float32."""
    return BatchNormReLUFunction.apply(x, weight, bias, running_mean, running_var, training, momentum, eps)


# ---- training: channels-last hand-over between a layer's post-op and the next layer ----------------
# The post-op writes the NEXT layer's staged input (framed channels-last copy at the head of that layer's workspace);
# the tensor that travels between the two autograd nodes is a float32 view of that workspace head, and its gradient
# is the consumer's channels-last grad_x accumulator (DCN_FLAG_GRAD_X_FRAMED) — nothing crosses the boundary in NCHW.
def stem_conv_supported(x, weight):
    """dcn_stem_conv_* takes this convolution (csrc/dcn_stem.cu: 3 x 3, stride 1, padding 1, Cin <= 4, O in {16, 32},
    W % 4 == 0, float32 CUDA tensors, no input gradient wanted)."""
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and tuple(weight.shape[2:]) == (3, 3) and weight.shape[1] == x.shape[1] <= 4
            and weight.shape[0] in (16, 32) and x.shape[3] % 4 == 0 and not x.requires_grad)


class StemConvFunction(torch.autograd.Function):
    """conv2d(x, weight, bias, stride 1, padding 1) of the detector's first layer (train.py:145,166) on the engine's two
    streaming kernels.  x is the network input: no gradient flows to it."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib.load()
        xd, wd = _dev_ready(x), _dev_ready(weight)
        bd = _dev_ready(bias) if bias is not None else None
        B, Cin, H, W = xd.shape
        O = wd.shape[0]
        out = torch.empty(B, O, H, W, dtype=torch.float32, device=xd.device)
        with torch.cuda.device(xd.device):
            stream = torch.cuda.current_stream(xd.device)
            rc = lib.dcn_stem_conv_forward(B, Cin, O, H, W, _ptr(xd), _ptr(wd), _ptr(bd), _ptr(out),
                                           ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_stem_conv_forward")
        ctx.save_for_backward(xd)
        ctx.O, ctx.has_bias = O, bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        xd, = ctx.saved_tensors
        B, Cin, H, W = xd.shape
        go = _dev_ready(grad_out)
        gw = torch.empty(ctx.O, Cin, 3, 3, dtype=torch.float32, device=xd.device)
        gb = torch.empty(ctx.O, dtype=torch.float32, device=xd.device)
        with torch.cuda.device(xd.device):
            stream = torch.cuda.current_stream(xd.device)
            rc = lib.dcn_stem_conv_backward(B, Cin, ctx.O, H, W, _ptr(xd), _ptr(go), _ptr(gw), _ptr(gb),
                                            ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_stem_conv_backward")
        return None, gw, gb if ctx.has_bias else None


def stem_conv(x, weight, bias=None):
    """The detector's conv1 on the engine; the caller checks stem_conv_supported first."""
    return StemConvFunction.apply(x, weight, bias)


def _framed_view(ws, B, H, W, C):
    n = B * (H + 3) * (W + 2) * C
    return ws[:4 * n].view(torch.float32).view(B, H + 3, W + 2, C)


def _storage_tensor(t):
    """uint8 tensor over the WHOLE storage `t` lives in (the workspace a staged tensor is the head of)."""
    return torch.empty(0, dtype=torch.uint8, device=t.device).set_(t.untyped_storage())


def _consumer_shape(x_shape, consumer):
    out_channels, kernel_size, stride, padding, variant, flags = consumer
    B, C, H, W = (int(v) for v in x_shape)
    return _lib.make_shape(B, C, int(out_channels), H, W, kernel_size, stride, padding, variant, OPERAND_FP32, flags)


class BatchNormReLUStagedFunction(torch.autograd.Function):
    """relu(batch_norm(x)) whose result is written as the staged input of `consumer` (a DCN layer).  Returns the float32
    view [B, H + 3, W + 2, C] of the consumer's workspace head."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, consumer):
        lib = _lib.load()
        x = _dev_ready(x)
        B, C, H, W = x.shape
        shp = _consumer_shape(x.shape, consumer)
        need = max(lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_FORWARD),
                   lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_BACKWARD))
        if need == 0:
            raise _lib.DcnError("staged post-op: the consumer layer does not run on the tensor path")
        dev = x.device
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        saved = torch.empty(4 * C, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            bws = torch.empty(lib.dcn_bn_workspace_bytes(C), dtype=torch.uint8, device=dev)
            rc = lib.dcn_bn_relu_forward_staged(ctypes.byref(shp), 1 if training else 0, _ptr(x), _ptr(weight), _ptr(bias),
                                                _ptr(running_mean), _ptr(running_var), float(momentum), float(eps),
                                                _ptr(ws), _ptr(saved), _ptr(bws), bws.numel(),
                                                ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_bn_relu_forward_staged")
        ctx.save_for_backward(x, saved)
        ctx.consumer, ctx.training = consumer, bool(training)
        ctx.has_affine = (weight is not None, bias is not None)
        return _framed_view(ws, B, H, W, C)

    @staticmethod
    def backward(ctx, grad_staged):
        lib = _lib.load()
        x, saved = ctx.saved_tensors
        C = x.shape[1]
        dev = x.device
        shp = _consumer_shape(x.shape, ctx.consumer)
        g = _dev_ready(grad_staged)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gg = torch.empty(C, dtype=torch.float32, device=dev) if ctx.has_affine[0] else None
        gb = torch.empty(C, dtype=torch.float32, device=dev) if ctx.has_affine[1] else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            bws = torch.empty(lib.dcn_bn_workspace_bytes(C), dtype=torch.uint8, device=dev)
            rc = lib.dcn_bn_relu_backward_staged(ctypes.byref(shp), 1 if ctx.training else 0, _ptr(x), _ptr(g), _ptr(saved),
                                                 _ptr(gx), _ptr(gg), _ptr(gb), _ptr(bws), bws.numel(),
                                                 ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_bn_relu_backward_staged")
        return gx, gg, gb, None, None, None, None, None, None


def batch_norm_relu_staged(x, weight, bias, running_mean, running_var, training, momentum, eps, consumer):
    """consumer = (out_channels, kernel_size, stride, padding, variant, flags) of the DCN layer that reads the result."""
    return BatchNormReLUStagedFunction.apply(x, weight, bias, running_mean, running_var, training, momentum, eps, consumer)


class DeformLayerFramedFunction(torch.autograd.Function):
    """The whole layer (offset conv + DCN span) on an input that is ALREADY staged: `xt` is the view a staged post-op
    returned (head of this layer's workspace).  Its gradient is returned in the same framed channels-last layout."""

    @staticmethod
    def forward(ctx, xt, offset_weight, offset_bias, weight, bias, cfg):
        lib = _lib.load()
        (C, H, W), kernel_size, stride, padding, variant, flags = cfg
        B = int(xt.shape[0])
        if tuple(xt.shape) != (B, H + 3, W + 2, C) or xt.dtype != torch.float32 or not xt.is_contiguous():
            raise ValueError(f"staged input of shape {tuple(xt.shape)} does not match the layer input {(B, C, H, W)}")
        ws = _storage_tensor(xt)
        shp = _lib.make_shape(B, C, int(weight.shape[0]), H, W, kernel_size, stride, padding, variant, OPERAND_FP32,
                              flags | FLAG_XT_STAGED)
        need = max(lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_FORWARD),
                   lib.dcn_workspace_bytes(ctypes.byref(shp), PHASE_LAYER_BACKWARD))
        if need == 0 or ws.numel() < need or xt.data_ptr() != ws.data_ptr():
            raise _lib.DcnError("staged input is not the head of a workspace of this layer")
        Ho, Wo = _lib.output_hw(shp)
        N = shp.kh * shp.kw
        dev = xt.device
        weight, bias = _dev_ready(weight), _dev_ready(bias)
        offset_weight, offset_bias = _dev_ready(offset_weight), _dev_ready(offset_bias)
        offset = torch.empty((B, 2 * N, Ho, Wo), dtype=torch.float32, device=dev)
        out = torch.empty((B, shp.O, Ho, Wo), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            rc = lib.dcn_layer_forward(ctypes.byref(shp), _ptr(ws), _ptr(offset_weight), _ptr(offset_bias), _ptr(weight),
                                       _ptr(bias), _ptr(offset), _ptr(out), _ptr(ws), ws.numel(),
                                       ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_layer_forward")
        ctx.relu = bool(flags & FLAG_RELU_OUT)
        ctx.save_for_backward(xt, offset, offset_weight, weight, *((out,) if ctx.relu else ()))
        ctx.cfg = cfg
        ctx.has = (offset_bias is not None, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        xt, offset, offset_weight, weight = ctx.saved_tensors[:4]
        if ctx.relu:
            grad_out = grad_out * (ctx.saved_tensors[4] > 0)
        (C, H, W), kernel_size, stride, padding, variant, flags = ctx.cfg
        B = int(xt.shape[0])
        ws = _storage_tensor(xt)
        need_gx = ctx.needs_input_grad[0]
        fl = (flags & ~FLAG_ACCUM_GRAD_X) | FLAG_XT_STAGED | _lib.FLAG_GRAD_X_FRAMED | (0 if need_gx else FLAG_NO_GRAD_X)
        shp = _lib.make_shape(B, C, int(weight.shape[0]), H, W, kernel_size, stride, padding, variant, OPERAND_FP32, fl)
        dev = xt.device
        grad_out = _dev_ready(grad_out)
        gxt = torch.empty_like(xt) if need_gx else None
        gwoff = torch.empty(offset_weight.shape, dtype=torch.float32, device=dev)
        gboff = torch.empty((offset_weight.shape[0],), dtype=torch.float32, device=dev) if ctx.has[0] else None
        gw = torch.empty(weight.shape, dtype=torch.float32, device=dev)
        gb = torch.empty((shp.O,), dtype=torch.float32, device=dev) if ctx.has[1] else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            rc = lib.dcn_layer_backward(ctypes.byref(shp), _ptr(ws), _ptr(offset), _ptr(offset_weight), _ptr(weight),
                                        _ptr(grad_out), _ptr(gxt), _ptr(gwoff), _ptr(gboff), _ptr(gw), _ptr(gb), _ptr(ws),
                                        ws.numel(), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dcn_layer_backward")
        return gxt, gwoff, gboff, gw, gb, None


def deform_layer_framed(xt, in_chw, offset_weight, offset_bias, weight, bias=None, kernel_size=3, stride=1, padding=1,
                        variant=VARIANT_TORCH, flags=0):
    """Whole layer on a staged input (see batch_norm_relu_staged); in_chw = (C, H, W) of the layer's logical input."""
    cfg = (tuple(int(v) for v in in_chw), _lib._pair(kernel_size), _lib._pair(stride), _lib._pair(padding), variant, flags)
    return DeformLayerFramedFunction.apply(xt, offset_weight, offset_bias, weight, bias, cfg)
