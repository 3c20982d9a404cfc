"""Harness model for BASELINE configs 1 and 5: the reference's toy detector topology
(`TorchEDNetDetection`, train.py:142-175) with its four DeformConv2d layers running on the B200
engine.  The DCN layers are ours, and with them (fused_bn_relu) the ops either side of them: relu(bn(x)) and conv1, the
producer of the first DCN layer's input (StemConv2d); pooling, the two heads and the optimiser are stock framework
ops, as in the reference.  Module and parameter names follow the reference so
that its checkpoints (train.py:293, test.py:17-22) load unchanged.
"""
import torch
import torch.nn as nn

from .torch_module import BatchNormReLU2d, StemConv2d, TorchDeformConv2d

# (name suffix, in, out) of the stride-2 DCN stages, train.py:149-158
_DCN_STAGES = ((2, 16, 32), (3, 32, 64), (4, 64, 128), (5, 128, 256))


class EDNetDetection(nn.Module):
    def __init__(self, num_classes=10, groups=2, dcn_cls=TorchDeformConv2d, fused_bn_relu=True, channels_last=False):
        super().__init__()
        # channels_last (needs fused_bn_relu): every relu(bn(x)) in front of a DCN layer writes that layer's staged
        # channels-last input directly and takes its gradient in the same layout: no NCHW <-> channels-last
        # transposition between the layers, in either direction
        self.channels_last = bool(channels_last) and fused_bn_relu
        # fused_bn_relu: relu(bn(x)) as ONE engine op (BatchNormReLU2d, same state-dict keys as nn.BatchNorm2d);
        # False keeps the framework's BatchNorm2d + ReLU as the reference has them
        self.fused_bn_relu = fused_bn_relu
        bn_cls = BatchNormReLU2d if fused_bn_relu else nn.BatchNorm2d
        del groups  # dead argument in the reference too (train.py:143,145)
        # conv1 on the engine's streaming kernels when the engine path is on (same parameters / state-dict keys)
        self.conv1 = (StemConv2d if fused_bn_relu else nn.Conv2d)(1, 16, 3, 1, 1)
        self.bn1 = bn_cls(16)
        self.relu = nn.ReLU(inplace=True)
        for idx, cin, cout in _DCN_STAGES:
            layer = dcn_cls(cin, cout, 3, 2, 1)
            # the backward pass reuses the forward pass's staged copy of x (one transpose of x per layer and step
            # instead of two; the scratch buffer then lives from forward to backward)
            layer.keep_staged_input = True
            setattr(self, f"conv{idx}", layer)
            setattr(self, f"bn{idx}", bn_cls(cout))
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.fc_cls = nn.Linear(256, num_classes)
        self.fc_bbox = nn.Linear(256, 4)

    def forward(self, x):
        if self.channels_last and x.is_cuda:
            return self._forward_channels_last(x)
        act = (lambda t: t) if self.fused_bn_relu else self.relu
        x = act(self.bn1(self.conv1(x)))
        for idx, _, _ in _DCN_STAGES:
            x = act(getattr(self, f"bn{idx}")(getattr(self, f"conv{idx}")(x)))
        feat = self.gap(x).flatten(1)
        return self.fc_cls(feat), torch.sigmoid(self.fc_bbox(feat))

    def _forward_channels_last(self, x):
        raw = self.conv1(x)                                   # NCHW, framework conv
        bn = self.bn1
        for idx, _, _ in _DCN_STAGES:
            layer = getattr(self, f"conv{idx}")
            hw = raw.shape[2:]
            xt = bn.forward_staged(raw, layer)                # relu(bn(raw)) -> the layer's staged channels-last input
            raw = layer.forward_staged(xt, hw)                # offset conv + DCN span, NCHW out (batch statistics next)
            bn = getattr(self, f"bn{idx}")
        x = bn(raw)
        feat = self.gap(x).flatten(1)
        return self.fc_cls(feat), torch.sigmoid(self.fc_bbox(feat))


def chained_eval_forward(model, x):
    """Inference forward of an EDNetDetection with channels-last activations between its four DCN stages:
    conv1 / bn1 / ReLU on the framework, then ONE layout pass into the engine, four chained `relu(bn(dcn(x)))` stages
    (jittor_dcn_b200/chain.py) and one NCHW store for the heads.  Same result as `model.eval()(x)`."""
    from .chain import ChainedDeformStages
    if model.training:
        raise ValueError("chained_eval_forward is inference only: call model.eval() first")
    chain = getattr(model, "_chained_stages", None)
    if chain is None:
        chain = ChainedDeformStages([(getattr(model, f"conv{i}"), getattr(model, f"bn{i}")) for i, _, _ in _DCN_STAGES])
        object.__setattr__(model, "_chained_stages", chain)     # not a sub-module: holds folded COPIES of the weights
    with torch.no_grad():
        act = (lambda t: t) if model.fused_bn_relu else model.relu
        h = act(model.bn1(model.conv1(x)))
        h = chain(h)
        feat = model.gap(h).flatten(1)
        return model.fc_cls(feat), torch.sigmoid(model.fc_bbox(feat))


def detection_loss(cls_logits, bbox, labels, boxes):
    """CE + 5 * smooth-L1(beta=1), the recipe of train.py:195-199,247."""
    return nn.functional.cross_entropy(cls_logits, labels) + \
        5.0 * nn.functional.smooth_l1_loss(bbox, boxes, beta=1.0)


def synthetic_canvases(batch, generator=None, device="cpu"):
    """MNISTDet-shaped synthetic data (prepare_data.py:8-29): a 28x28 patch of U(0,1) noise on a
    zero 1x128x128 canvas, label in [0,10), box = [x, y, x+28, y+28] / 128."""
    g = generator
    x = torch.zeros(batch, 1, 128, 128)
    pos = torch.randint(0, 101, (batch, 2), generator=g)
    patch = torch.rand(batch, 28, 28, generator=g)
    for i in range(batch):
        px, py = int(pos[i, 0]), int(pos[i, 1])
        x[i, 0, py:py + 28, px:px + 28] = patch[i]
    labels = torch.randint(0, 10, (batch,), generator=g)
    boxes = torch.stack([pos[:, 0], pos[:, 1], pos[:, 0] + 28, pos[:, 1] + 28], 1).float() / 128.0
    return x.to(device), labels.to(device), boxes.to(device)
