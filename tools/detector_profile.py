"""Kernel table (torch.profiler / CUPTI: engine AND framework kernels) of the detector training step, BASELINE configs[4]
on one GPU.  Usage: python tools/detector_profile.py [batch] > profiles/<name>.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_dcn_b200 as dcn  # noqa: E402
from jittor_dcn_b200.detector import EDNetDetection, detection_loss, synthetic_canvases  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = EDNetDetection(dcn_cls=dcn.TorchDeformConv2d, fused_bn_relu=True, channels_last=True).to(dev)
with torch.no_grad():
    for m in model.modules():
        if isinstance(m, dcn.TorchDeformConv2d):
            m.offset_conv.weight.normal_(0, 0.01)
            m.offset_conv.bias.normal_(0, 1.0)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
x, labels, boxes = synthetic_canvases(B, torch.Generator().manual_seed(100), dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = detection_loss(*model(x), labels, boxes)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.count / N, e.device_time_total / N / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"detector training step, batch {B}, eager, {N} steps averaged: {tot:.3f} ms of kernels per step")
for k, c, ms in rows:
    print(f"{ms:9.4f} ms  {100 * ms / tot:5.1f} %  x{c:5.1f}  {k[:150]}")
