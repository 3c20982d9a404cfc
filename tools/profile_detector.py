"""Where does the detector training step (BASELINE configs[4]) spend its GPU time?  torch.profiler table.
usage: python tools/profile_detector.py [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from jittor_dcn_b200.detector import EDNetDetection, detection_loss, synthetic_canvases

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = EDNetDetection().cuda()
for mod in m.modules():
    if hasattr(mod, "offset_conv"):
        with torch.no_grad():
            mod.offset_conv.bias.normal_(0, 1.0)
opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)
x, labels, boxes = synthetic_canvases(B, torch.Generator().manual_seed(0), "cuda")

def step():
    opt.zero_grad(set_to_none=True)
    loss = detection_loss(*m(x), labels, boxes)
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
