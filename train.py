#!/usr/bin/env python
"""Training-step harness: torch GridNet-style flow producer -> fused
warp+loss op (hand-written sm_100a kernels) -> DDP gradient all-reduce -> Adam.

    python train.py [--steps 20] [--warmup 3] [--batch 16] [--height 256] [--width 512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 train.py ...

Shape of the step follows src/trainer.py:168-258: inputs cat([seg1, frame1, frame2, seg2]) (8
channels, src/trainer.py:461), loss 40*L1 + 20*(GD+SSIM) + 10*CE (+ TV on the flow), one
all-reduce for the loss vector instead of one per scalar (src/trainer.py:381-386).  Unlike the
reference (Appendix B #2) gradients are zeroed every step.  Data: synthetic, Cityscapes-shaped.
Prints one JSON line with iterations/s (max over ranks, CUDA events).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def run_training(steps=20, warmup=3, batch=16, height=256, width=512, classes=20, seed=1024, lr=2e-4, verbose=False,
                 amp="none"):
    import vlg_b200
    from vlg_b200.producer import FlowGridNet, flow_nhw2
    from vlg_b200 import parallel
    from bench import make_inputs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    own_pg = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=dev)
        own_pg = True

    torch.backends.cudnn.benchmark = True         # src/main.py:122 does the same
    torch.manual_seed(seed)                       # src/main.py:121 seeds with 1024
    net = FlowGridNet(in_channels=8).to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank])
    opt = torch.optim.Adam(net.parameters(), lr=lr, betas=(0.5, 0.999))     # src/main.py:139-141
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))

    d = make_inputs(batch, height, width, classes, 1.0, 0.0, "f32", dev, seed=seed + rank)
    frame1, frame2, frame3 = d["src_rgb"], d["tgt_rgb"].flip(0).contiguous(memory_format=torch.channels_last), d["tgt_rgb"]
    seg2_onehot = d["src_layout"]
    seg2 = seg2_onehot.argmax(1, keepdim=True).float()
    seg1 = seg2.roll(1, 0)
    seg3 = d["tgt_label"]
    x = torch.cat([seg1, frame1, frame2, seg2], 1).contiguous(memory_format=torch.channels_last)

    def step():
        opt.zero_grad(set_to_none=True)
        # the producer (library convolutions) may run under bf16 autocast; the flow and everything on the
        # hand-written path stay fp32
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(amp == "bf16")):
            flow, _ = net(x)
        loss = crit(frame2, seg2_onehot, flow_nhw2(flow.float()), frame3, seg3)
        loss.backward()
        opt.step()
        return parallel.sync_loss_vector(crit.last_terms, "reference") if world > 1 else crit.last_terms

    for _ in range(max(warmup, 1)):
        terms = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    first = None
    for i in range(steps):
        terms = step()
        if i == 0:
            first = terms[5].item()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {
        "iters_per_s": 1e3 / ms.item(), "ms_per_iter": ms.item(), "n_gpus": world, "global_batch": batch * world,
        "per_gpu_batch": batch, "resolution": [height, width], "producer": "FlowGridNet 3x6 (32/64/96), %s, torch/cuDNN" % ("bf16 autocast" if amp == "bf16" else "fp32"),
        "loss_first": first, "loss_last": terms[5].item(),
        "params": sum(p.numel() for p in net.parameters()),
    }
    if own_pg:
        dist.destroy_process_group()
    return out, rank


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (weak scaling)")
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--amp", default="none", choices=["none", "bf16"], help="autocast dtype of the torch producer")
    args = ap.parse_args()
    out, rank = run_training(args.steps, args.warmup, args.batch, args.height, args.width, amp=args.amp)
    if rank == 0:
        print(json.dumps({"metric": "train iters/s", **out}))


if __name__ == "__main__":
    main()
