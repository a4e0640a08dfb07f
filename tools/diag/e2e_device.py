"""Diagnostic: the device work of one captured end-to-end step, kernel by kernel (CUPTI through torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench, vlg_b200 as vlg
from torch.profiler import profile, ProfilerActivity
cx = bench.Ctx()
fz = bench.Fused(cx, "c2", 16, use_graph=False)
host = bench.u8_host_inputs(fz.sets[0], fz.K)
buf = {k: v.to(cx.dev) for k, v in host.items()}
crit = vlg.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))
def device_step():
    src = vlg.ingest(buf["src_u8"], buf["src_seg_u8"], n_classes=fz.K, want_label=False, want_one_hot=True)
    tgt = vlg.ingest(buf["tgt_u8"], buf["tgt_seg_u8"], n_classes=fz.K, want_label=True)
    a, b = src["frames"].requires_grad_(True), src["one_hot"].requires_grad_(True)
    f = buf["flow"].detach().requires_grad_(True)
    crit(a, b, f, tgt["frames"], tgt["label"]).backward()
    return crit.last_terms
step = vlg.CapturedStep(device_step)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): step()
e1.record(); torch.cuda.synchronize()
print("captured step alone: %.1f us" % (e0.elapsed_time(e1) * 20))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
tot = 0
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    print(f"  {ev.device_time_total / 5:8.1f} us x{ev.count / 5:.0f}  {ev.key[:90]}")
    tot += ev.device_time_total / 5
print("sum %.1f us" % tot)
