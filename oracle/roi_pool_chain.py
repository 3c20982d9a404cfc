"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's DeformRoIPool / DeformPSRoIPool
(`/root/reference/deform_conv.py:83-157`, `:160-241`) with the same op chain in torch (jt.maximum -> torch.maximum,
jt.floor / clamp / meshgrid / advanced indexing one to one), differentiable through torch autograd.

PARITY UNPINNED: the reference classes are Jittor modules and jittor cannot be installed in the build container; the
reference's tests hold no vectors for them.  Float index tensors (`offsets[:, ph_grid * pooled_w + pw_grid, 0]`,
:113-114 / :209-210) are cast to integers here, which is what the expression can only mean.  Only pooled 1 x 1 runs in
the reference (the final reshape, :157 / :241), and only that is restated.
"""
import torch


def deform_roi_pool(features, rois, offsets, spatial_scale=1.0):
    """deform_conv.py:90-157 with output_size = (1, 1).  offsets [R, 1, 2]."""
    B, C, H, W = features.shape
    num_rois = rois.shape[0]
    pooled_h = pooled_w = 1
    batch_indices = rois[:, 0].long()                                        # :95
    roi_coords = rois[:, 1:5] * spatial_scale                                # :96
    x1, y1, x2, y2 = roi_coords[:, 0], roi_coords[:, 1], roi_coords[:, 2], roi_coords[:, 3]
    roi_w = torch.maximum(x2 - x1, torch.tensor(1e-6))                       # :98
    roi_h = torch.maximum(y2 - y1, torch.tensor(1e-6))                       # :99
    ph_grid = torch.zeros(1)                                                 # :101-105 for one bin
    pw_grid = torch.zeros(1)
    bin_w = roi_w[:, None] / pooled_w                                        # :107
    bin_h = roi_h[:, None] / pooled_h
    bin_cx = x1[:, None] + (pw_grid + 0.5) * bin_w                           # :110
    bin_cy = y1[:, None] + (ph_grid + 0.5) * bin_h
    idx = (ph_grid * pooled_w + pw_grid).long()
    offset_x = offsets[:, idx, 0] * roi_w[:, None]                           # :113
    offset_y = offsets[:, idx, 1] * roi_h[:, None]                           # :114
    cx = bin_cx + offset_x                                                   # :116
    cy = bin_cy + offset_y
    return _sample(features, batch_indices, cx, cy, num_rois, C, H, W)


def deform_psroi_pool(features, rois, offsets, spatial_scale=1.0, no_trans=False, trans_std=0.1):
    """deform_conv.py:172-241 with output_size = part_size = (1, 1).  offsets [R, 2]."""
    B, C, H, W = features.shape
    num_rois = rois.shape[0]
    batch_indices = rois[:, 0].long()                                        # :179
    roi_coords = rois[:, 1:5] * spatial_scale
    x1, y1, x2, y2 = roi_coords[:, 0], roi_coords[:, 1], roi_coords[:, 2], roi_coords[:, 3]
    roi_w = torch.maximum(x2 - x1, torch.tensor(1e-6))                       # :182
    roi_h = torch.maximum(y2 - y1, torch.tensor(1e-6))
    ph_flat = torch.zeros(1)
    pw_flat = torch.zeros(1)
    part_idx = (ph_flat * 1 + pw_flat).long()                                # :191
    bin_w = roi_w[:, None] / 1                                               # :193
    bin_h = roi_h[:, None] / 1
    bin_cx = x1[:, None] + (pw_flat + 0.5) * bin_w                           # :196
    bin_cy = y1[:, None] + (ph_flat + 0.5) * bin_h
    if not no_trans:
        trans_x = offsets[:, part_idx * 2] * roi_w[:, None] * trans_std      # :200
        trans_y = offsets[:, part_idx * 2 + 1] * roi_h[:, None] * trans_std  # :201
        cx = bin_cx + trans_x
        cy = bin_cy + trans_y
    else:
        cx, cy = bin_cx, bin_cy
    # one bin: channel_idx = c_out * 1 + 0 = c (:224-226), so the gather is the same as DeformRoIPool's
    return _sample(features, batch_indices, cx, cy, num_rois, C, H, W)


def _sample(features, batch_indices, cx, cy, num_rois, C, H, W):
    """deform_conv.py:120-157 (= :208-241): clamp the corner indices, THEN form the fractions."""
    x0 = torch.floor(cx).long()
    x1 = x0 + 1
    y0 = torch.floor(cy).long()
    y1 = y0 + 1
    x0 = torch.clamp(x0, 0, W - 1)
    x1 = torch.clamp(x1, 0, W - 1)
    y0 = torch.clamp(y0, 0, H - 1)
    y1 = torch.clamp(y1, 0, H - 1)
    dx = cx - x0.float()
    dy = cy - y0.float()
    w00 = (1 - dx) * (1 - dy)
    w01 = (1 - dx) * dy
    w10 = dx * (1 - dy)
    w11 = dx * dy
    batch_idx = batch_indices[:, None].repeat(1, 1)

    def corner(yy, xx, w):
        feats = features[batch_idx.reshape(-1), :, yy.reshape(-1), xx.reshape(-1)]
        feats = feats.reshape(num_rois, 1, C).permute(0, 2, 1)
        return (feats * w[:, None, :]).sum(dim=2)

    output = corner(y0, x0, w00) + corner(y1, x0, w01) + corner(y0, x1, w10) + corner(y1, x1, w11)
    return output.reshape(num_rois, C, 1, 1)
