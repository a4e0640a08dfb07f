"""Diagnostic: device time of vlg_b200.one_hot_layout on a C2-sized class map (168 MB of fp32 one-hot layout)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, vlg_b200
lab = torch.randint(0, 20, (16, 256, 512), device="cuda")
for _ in range(3): vlg_b200.one_hot_layout(lab, 20)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): vlg_b200.one_hot_layout(lab, 20)
e1.record(); torch.cuda.synchronize(); print("one_hot_layout 16x256x512: %.1f us" % (e0.elapsed_time(e1) * 20))
