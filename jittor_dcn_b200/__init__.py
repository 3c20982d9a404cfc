"""jittor_dcn_b200 — B200-native engine behind the reference's DeformConv2d / TorchDeformConv2d.

Only what the hot path needs: the C-ABI library (csrc/ -> libdcn_b200.so), its ctypes
binding and the two module classes that mirror the reference's operator interface.
"""
from ._lib import (DcnError, DcnShape, FLAG_ACCUM_GRAD_X, FLAG_FORCE_SIMT, FLAG_NO_GRAD_X, FLAG_RELU_OUT,
                   OPERAND_BF16, OPERAND_FP32, VARIANT_DCNV1, VARIANT_JITTOR, VARIANT_TORCH, load, make_shape)
from .deform_conv import DeformConv2d
from .functional import (batch_norm_relu, clear_workspaces, dcn_backward, dcn_corners, dcn_forward,
                         dcn_layer_backward, dcn_layer_forward, dcn_offset_conv_forward, deform_conv2d, deform_conv2d_v1,
                         deform_layer, layer_supported)
from .chain import ChainedDeformStages
from .roi_pool import DeformPSRoIPool, DeformRoIPool
from .torch_module import (BatchNormReLU2d, StemConv2d, TorchDeformConv2d, TorchDeformConv2dJittorSemantics,
                           fuse_eval_bn_relu)

__all__ = [
    "DeformConv2d", "TorchDeformConv2d", "BatchNormReLU2d", "batch_norm_relu", "TorchDeformConv2dJittorSemantics", "deform_conv2d",
    "dcn_forward", "dcn_backward", "dcn_corners", "deform_conv2d_v1", "VARIANT_DCNV1", "load", "make_shape", "DcnShape", "DcnError",
    "VARIANT_JITTOR", "VARIANT_TORCH", "OPERAND_FP32", "OPERAND_BF16", "FLAG_ACCUM_GRAD_X",
    "FLAG_FORCE_SIMT", "FLAG_NO_GRAD_X", "FLAG_RELU_OUT", "clear_workspaces", "dcn_offset_conv_forward", "dcn_layer_forward", "dcn_layer_backward", "deform_layer",
    "layer_supported", "fuse_eval_bn_relu", "DeformRoIPool", "DeformPSRoIPool", "ChainedDeformStages", "StemConv2d",
]
