#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: warp+loss fwd/bwd Mpixel/s and
HBM GB/s vs peak) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c5|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused op over one batch of synthetic Cityscapes-shaped input
(BASELINE.json configs[1]: 16x256x512, K=20, fp32): flow-guided warp of RGB + layout, all loss
terms, gradients to flow, src_rgb and src_layout.  Weak scaling: every rank processes one such
batch; the only exchange is one NCCL all-reduce of the 8-float loss vector.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
public module API from pinned host buffers, `roofline` is live CUDA-event timing of the dominant
kernel against MEASURED_PEAKS.json, `cpu_baseline` is the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (N, H, W, K, flow sigma px, far fraction, dtype)
    "c1": (2, 128, 256, 20, 4.0, 0.0, "f32"),
    "c2": (16, 256, 512, 20, 4.0, 0.0, "f32"),
    "c2calm": (16, 256, 512, 20, 0.5, 0.0, "f32"),    # tuning: same shape, flow so small that no tap leaves the staged windows
    "c3": (8, 1024, 2048, 20, 4.0, 0.0, "bf16"),
    "c5": (16, 375, 1242, 20, 48.0, 0.05, "f32"),
}
BYTES_PER_PX = {"f32": dict(step=220, pass1=128, pass2=92), "bf16": dict(step=122, pass1=76, pass2=46)}
FALLBACK_HBM_GBS = 6650.0


def make_inputs(N, H, W, K, sigma, far, dtype, device, seed=1024):
    """Synthetic Cityscapes-shaped batch: normalised RGB, block-constant one-hot
    layouts, block-constant int64 labels, box-smoothed Gaussian flow."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    src_rgb = (torch.rand(N, 3, H, W, generator=g) - mean) / std
    tgt_rgb = (torch.rand(N, 3, H, W, generator=g) - mean) / std
    blk = 32

    def labels():
        l = torch.randint(0, K, (N, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
        return l.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()

    lab_src, lab_tgt = labels(), labels()
    src_layout = torch.zeros(N, K, H, W).scatter_(1, lab_src[:, None], 1.0)
    flow = torch.randn(N, 2, H, W, generator=g) * sigma
    flow = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(flow, (4, 4, 4, 4), mode="replicate"), 9, 1)
    if far > 0:
        m = torch.rand(N, 1, H, W, generator=g) < far
        flow = torch.where(m, (torch.rand(N, 2, H, W, generator=g) - 0.5) * 2 * max(H, W), flow)
    flow = flow.permute(0, 2, 3, 1).contiguous()
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    out = dict(src_rgb=src_rgb.to(tdt), src_layout=src_layout.to(tdt), flow=flow, tgt_rgb=tgt_rgb.to(tdt),
               tgt_label=lab_tgt)
    if device is not None:
        cl = lambda t: t.to(device).contiguous(memory_format=torch.channels_last)
        out = dict(src_rgb=cl(out["src_rgb"]), src_layout=cl(out["src_layout"]), flow=flow.to(device),
                   tgt_rgb=cl(out["tgt_rgb"]), tgt_label=lab_tgt.to(device))
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_throughput(N, H, W, K, sigma, far, n_sample, iters):
    """Oracle port (torch CPU composition of the reference's losses + grid_sample) on a bounded
    sample of the workload: fwd+bwd to flow, src_rgb, src_layout.  Returns Mpixel/s."""
    from oracle import torch_oracle as TO
    d = make_inputs(n_sample, H, W, K, sigma, far, "f32", None)
    ts = []
    for i in range(iters + 1):
        t0 = time.perf_counter()
        TO.warp_loss_fwd_bwd(d["src_rgb"], d["src_layout"], d["flow"], d["tgt_rgb"], d["tgt_label"], w_tv=0.5)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])  # first call = warm-up
    t = ts[len(ts) // 2]
    return n_sample * H * W / t / 1e6, t


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is
    pure Python/PyTorch with no compiled code of its own, so this is the oracle port
    (oracle/torch_oracle.py) on the box's host cores, all threads."""
    if rank != 0:
        return
    N, H, W, K, sigma, far, dtype = WORKLOADS[args.workload]
    n_sample = min(N, 2)
    ts = []
    try:   # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it may
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    from oracle import torch_oracle as TO
    d = make_inputs(n_sample, H, W, K, sigma, far, "f32", None)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        TO.warp_loss_fwd_bwd(d["src_rgb"], d["src_layout"], d["flow"], d["tgt_rgb"], d["tgt_label"], w_tv=0.5)
        if i >= args.warmup:
            ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    val = n_sample * H * W / t / 1e6
    sample = f"{n_sample}x{H}x{W} slice of the {N}x{H}x{W} batch per step, torch CPU oracle port, fwd+bwd"
    print(json.dumps({
        "impl": "reference", "metric": "warp+loss fwd/bwd throughput", "value": val, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {N}x{H}x{W} K={K} fp32 fwd+bwd (bounded sample {n_sample}x{H}x{W} per step)"},
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--no-eager", action="store_true", help="skip timing the torch CUDA eager incumbent")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (iters/s) measurement")
    ap.add_argument("--flow-grad-only", action="store_true", help="sources are data: no d_src (128 B/px variant)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import vlg_b200
    from vlg_b200 import _cabi
    from vlg_b200 import ops as vops

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=dev)

    N, H, W, K, sigma, far, dtype = WORKLOADS[args.workload]
    P = N * H * W
    with_src = not args.flow_grad_only
    lib = _cabi.load()
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16

    # two input sets, alternated, so no step re-reads lines its predecessor left in L2
    sets = [make_inputs(N, H, W, K, sigma, far, dtype, dev, seed=1024 + 7 * rank + s) for s in range(2)]
    cfg = vops.WarpLossConfig(w_tv=0.5)
    prob = vops._problem(N, H, W, K, tdt, cfg)
    ws = vops._workspace(prob, with_src, dev)
    # same problem without the far path: vlg_warp_bwd_src then launches pass2_kernel ALONE, so that CUDA
    # events bracket exactly one kernel (the far-path launches are idle for near flows and timed separately)
    prob_p2 = vops._problem(N, H, W, K, tdt, vops.WarpLossConfig(w_tv=0.5, assume_near=True))
    loss = torch.zeros(_cabi.LOSS_SLOTS, dtype=torch.float32, device=dev)
    d_c = torch.empty(N, H, W, 2, dtype=torch.float32, device=dev)
    d_a = vops.empty_nhwc((N, 3, H, W), tdt, dev) if with_src else None
    d_b = vops.empty_nhwc((N, K, H, W), tdt, dev) if with_src else None
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    ptr = vops._ptr

    def launch_fused(s, sp_):
        vops.check(lib.vlg_warp_loss_fwd_bwd(C.byref(prob), ptr(s["src_rgb"]), ptr(s["src_layout"]), ptr(s["flow"]),
                                             ptr(s["tgt_rgb"]), ptr(s["tgt_label"]), ptr(loss), ptr(d_c), ptr(d_a),
                                             ptr(d_b), None, ptr(ws), ws.numel(), sp_))

    def launch_pass1(s, sp_):
        vops.check(lib.vlg_warp_loss_pass1(C.byref(prob), ptr(s["src_rgb"]), ptr(s["src_layout"]), ptr(s["flow"]),
                                           ptr(s["tgt_rgb"]), ptr(s["tgt_label"]), ptr(loss), ptr(d_c), None, int(with_src),
                                           ptr(ws), ws.numel(), sp_))

    def launch_pass2(s, sp_):
        if with_src:
            vops.check(lib.vlg_warp_bwd_src(C.byref(prob), ptr(s["flow"]), ptr(d_a), ptr(d_b), ptr(ws), ws.numel(), sp_))

    graphs = None
    launches_per_step = None

    def step(i, evs=None):
        """One pass of the hot path.  The timed region uses the fused entry point (the last pass-1
        CTA reduces the loss vector); with `evs` the same work is issued piecewise so that CUDA
        events can bracket each kernel."""
        s = sets[i & 1]
        if evs is None and dist is not None:
            # data parallel: the loss vector is complete after pass 1 (its last CTA reduces), so the path's only
            # exchange -- one all-reduce of 8 floats -- is issued there and overlaps pass 2
            if graphs is not None: graphs[i & 1][0].replay()
            else: launch_pass1(s, sp)
            work = dist.all_reduce(loss, async_op=True)
            if graphs is not None: graphs[i & 1][1].replay()
            else: launch_pass2(s, sp)
            work.wait()
            return
        if evs is None and graphs is not None:
            graphs[i & 1].replay()                  # the same launches, captured once per input set
        elif evs is None:
            launch_fused(s, sp)
        else:
            vops.check(lib.vlg_warp_loss_bwd_out(C.byref(prob), ptr(s["src_rgb"]), ptr(s["src_layout"]), ptr(s["flow"]),
                                                 ptr(s["tgt_rgb"]), ptr(s["tgt_label"]), ptr(d_c), None, int(with_src),
                                                 ptr(ws), ws.numel(), sp))
            evs[1].record(stream)
            vops.check(lib.vlg_reduce_partials(C.byref(prob), ptr(loss), ptr(ws), ws.numel(), sp))
            evs[2].record(stream)
            if with_src:
                vops.check(lib.vlg_warp_bwd_src(C.byref(prob_p2), ptr(s["flow"]), ptr(d_a), ptr(d_b), ptr(ws), ws.numel(), sp))
        if dist is not None:
            dist.all_reduce(loss)          # the path's only exchange: one 8-float loss vector
        if evs: evs[3].record(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # CUDA graph of the step's launches (memset + count + pass 1 + far path + pass 2), one per input set
    if not args.no_graph:
        try:
            n_before = vlg_b200.launch_count()
            gs = []
            for k in range(2):
                if dist is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        launch_fused(sets[k], C.c_void_p(torch.cuda.current_stream().cuda_stream))
                    gs.append(g)
                else:   # two graphs per input set: the all-reduce is issued between them
                    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        launch_pass1(sets[k], C.c_void_p(torch.cuda.current_stream().cuda_stream))
                    with torch.cuda.graph(g2):
                        launch_pass2(sets[k], C.c_void_p(torch.cuda.current_stream().cuda_stream))
                    gs.append((g1, g2))
            launches_per_step = (vlg_b200.launch_count() - n_before) // 2
            graphs = gs
            for i in range(4):
                step(i)
            barrier()
        except Exception as exc:   # capture unsupported: plain launches
            print(f"[bench] CUDA graph capture failed ({exc}); using plain launches", file=sys.stderr)
            graphs = None

    # ---- timed region 1: device-resident throughput (`value`) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = vlg_b200.launch_count()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    launches = vlg_b200.launch_count() - n0
    if graphs is not None:
        launches = launches_per_step * args.steps   # replayed from the graph: the host counter does not move
    ms_total = e0.elapsed_time(e1)
    t_step = torch.tensor([ms_total / args.steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    ms_per_step = t_step.item()
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing (same stream, CUDA events between the launches) ----
    reps = min(args.steps, 20)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
    barrier()
    for i in range(reps):
        evs[i][0].record(stream)
        step(i, evs[i])
    barrier()
    k1 = sorted(e[0].elapsed_time(e[1]) for e in evs)[reps // 2]   # pass-1 stage: memset + count_valid + rgb strip + layout tile
    k2 = sorted(e[2].elapsed_time(e[3]) for e in evs)[reps // 2]   # pass2_kernel alone (+ the all-reduce when world > 1)
    kr = sorted(e[1].elapsed_time(e[2]) for e in evs)[reps // 2]

    # ---- per-kernel device times of the fused step (CUPTI through torch.profiler; informational) ----
    kernels_us = None
    if rank == 0:
        try:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for i in range(10):
                    launch_fused(sets[i & 1], sp)
                torch.cuda.synchronize()
            kernels_us = {e.key.split("(")[0].replace("void ", "").replace("vlg::", ""): round(e.device_time_total / 10, 2)
                          for e in prof.key_averages() if e.device_time_total > 0}
        except Exception as exc:   # profiler unavailable: the event timings above stand on their own
            kernels_us = {"unavailable": str(exc)}

    # ---- timed region 2: end to end through the public module API from pinned host memory ----
    # The source layout travels as the class-id map the reference's dataset hands out (float32 [N,1,H,W],
    # src/folder.py:97-99) and is one-hot encoded on the device like src/models/net_utils.py:14-24
    # (`vlg_one_hot`): 4 bytes per pixel over PCIe instead of 4*K.
    host = {k: v.cpu().pin_memory() for k, v in sets[0].items() if k != "src_layout"}
    host["src_seg"] = sets[0]["src_layout"].argmax(1, keepdim=True).float().cpu().pin_memory()
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # Double-buffered: while step i computes, the inputs of step i+1 travel on a copy stream into the other
    # device buffer set (what a DataLoader prefetcher does).  Every step still uploads one full input set
    # and reads its result back; the timed region holds exactly n_e2e uploads, n_e2e steps, n_e2e reads.
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    for b_ in dbuf:   # keep the NHWC storage of the image tensors
        for k in ("src_rgb", "tgt_rgb"):
            b_[k] = torch.empty_strided(host[k].shape, host[k].stride(), dtype=host[k].dtype, device=dev)
    up_done = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(j):
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                dbuf[j][k].copy_(v, non_blocking=True)
            up_done[j].record(copy_stream)

    def e2e_step(i):
        j = i & 1
        upload(j ^ 1)                      # inputs of the NEXT step (its buffer was released by the previous read-back)
        stream.wait_event(up_done[j])
        cur = dbuf[j]
        a = cur["src_rgb"].detach().requires_grad_(with_src)
        b = vlg_b200.one_hot_layout(cur["src_seg"], K, tdt).requires_grad_(with_src)
        f = cur["flow"].detach().requires_grad_(True)
        total = crit(a, b, f, cur["tgt_rgb"], cur["tgt_label"])
        total.backward()
        return crit.last_terms.cpu()       # device -> host read of the step's result (synchronises)

    upload(0)
    for i in range(4):
        e2e_step(i)
    barrier()
    n_e2e = max(4, min(args.steps, 10)) & ~1
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for i in range(n_e2e):
        e2e_step(i)
    e3.record(stream)
    barrier()
    t_e2e = torch.tensor([e2.elapsed_time(e3) / n_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)

    # ---- incumbent on the same GPU: the oracle composition run by torch CUDA eager (ATen kernels) ----
    eager = None
    if rank == 0 and dtype == "f32" and not args.no_eager:
        from oracle import torch_oracle as TO
        s0 = sets[0]
        def eager_step():
            a = s0["src_rgb"].detach().requires_grad_(with_src)
            b = s0["src_layout"].detach().requires_grad_(with_src)
            f = s0["flow"].detach().requires_grad_(True)
            o = TO.warp_loss(a, b, f, s0["tgt_rgb"], s0["tgt_label"], w_tv=0.5)
            o["total"].backward()
        for _ in range(2):
            eager_step()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(5):
            eager_step()
        g1.record(stream)
        torch.cuda.synchronize()
        ms_eager = g0.elapsed_time(g1) / 5
        eager = {"ms_per_step": ms_eager, "value": P / (ms_eager * 1e-3) / 1e6, "unit": "Mpixel/s",
                 "what": "oracle composition (F.grid_sample + losses + autograd) in torch CUDA eager, same inputs, checker only"}

    # ---- boundary fusion: renorm + flip + NCHW->NHWC of rgb frames, one pass, HBM-bound ----
    aux = None
    if rank == 0:
        fr = torch.rand(8, 3, 1024, 2048, device=dev)          # 403 MB read+written per call: larger than the L2
        for _ in range(3):
            vlg_b200.prepare_frames(fr, flip=True)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(10):
            vlg_b200.prepare_frames(fr, flip=True)
        a1.record(stream)
        torch.cuda.synchronize()
        ms_fa = a0.elapsed_time(a1) / 10
        gbs = fr.numel() * 4 * 2 / (ms_fa * 1e-3) / 1e9
        aux = {"frame_affine": {"what": "vlg_frame_affine 8x3x1024x2048 fp32 NCHW -> normalised, flipped NHWC (24 B/px algorithmic; "
                                        "includes the output allocation of the Python wrapper)",
                                "ms": ms_fa, "achieved_GBs": gbs}}
        del fr
        # ---- rollout (BASELINE config 4): 5 autoregressive steps at 512x1024 with LABEL layout sources ----
        def timed(fn, reps):
            for _ in range(2):
                fn()
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record(stream)
            for _ in range(reps):
                fn()
            t1_.record(stream)
            torch.cuda.synchronize()
            return t0_.elapsed_time(t1_) / reps
        rn, rh, rw = 4, 512, 1024                                # 32 / 8 GPUs per rank
        rd = make_inputs(rn, rh, rw, K, sigma, 0.0, "f32", dev, seed=77)
        rlab = rd["src_layout"].argmax(1)
        rflows = [make_inputs(rn, rh, rw, K, sigma, 0.0, "f32", dev, seed=78 + t)["flow"] for t in range(5)]
        ms_ro = timed(lambda: vlg_b200.rollout(rd["src_rgb"], rlab, lambda t, im, lb: rflows[t], steps=5), 5)
        ro_px = rn * rh * rw * 5
        # per step and pixel: coords 8 + rgb 12 + label 8 read, rgb 12 + label 8 written = 48 B
        aux["rollout"] = {"what": "vlg_warp_fwd_labels x5 (4x512x1024 per GPU, label sources fed back: 48 B/px algorithmic "
                                  "against 200 B/px with dense one-hot layouts); includes the Python wrapper's allocations",
                          "ms": ms_ro, "Mpixel_steps_per_s": ro_px / (ms_ro * 1e-3) / 1e6,
                          "achieved_GBs": ro_px * 48 / (ms_ro * 1e-3) / 1e9}
        # ---- validation forward (BASELINE config 3 shape, one image set that exceeds the L2): warp + argmax, outputs materialised ----
        vn, vh, vw = 2, 1024, 2048
        vd = make_inputs(vn, vh, vw, K, sigma, 0.0, dtype, dev, seed=79)
        ms_fw = timed(lambda: vlg_b200.warp(vd["src_rgb"], vd["src_layout"], vd["flow"]), 5)
        fw_bpp = 200 if dtype == "f32" else 108
        aux["warp_fwd"] = {"what": f"vlg_warp_fwd {vn}x{vh}x{vw} {dtype}: warped rgb + layout + argmax written ({fw_bpp} B/px algorithmic)",
                           "ms": ms_fw, "Mpixel_per_s": vn * vh * vw / (ms_fw * 1e-3) / 1e6,
                           "achieved_GBs": vn * vh * vw * fw_bpp / (ms_fw * 1e-3) / 1e9}
        del rd, rflows, vd

    # ---- training step: torch flow producer -> fused op -> DDP -> Adam, iters/s ----
    train = None
    if args.workload == "c2" and not args.no_train:
        from train import run_training
        train, _ = run_training(steps=max(3, min(args.steps, 10)), warmup=2, batch=N, height=H, width=W, classes=K)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bpp = BYTES_PER_PX[dtype]
    step_bytes = bpp["step"] if with_src else bpp["pass1"]
    value = world * P / (ms_per_step * 1e-3) / 1e6
    # dominant single kernel: pass2_rec_kernel (deterministic source gradient from pass 1's tap records); without source gradients the
    # pass-1 stage (rgb strip + layout tile kernels, not separately callable) is reported instead
    dom_name, dom_ms, dom_bytes = ("pass2_rec_kernel", k2, bpp["pass2"]) if with_src else ("pass-1 stage (rgb_strip_kernel + lay_tile_kernel)", k1, bpp["pass1"])
    achieved = P * dom_bytes / (dom_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full`
    # capture of this same workload (profiles/r01_ncu_traffic.json); null for other workloads
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if args.workload == "c2" and with_src and os.path.exists(tpath):
        k = json.load(open(tpath))["kernels"].get(dom_name)
        if k:
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
    out = {
        "metric": "warp+loss fwd/bwd throughput", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {N}x{H}x{W} K={K} {dtype} warp+loss fwd+bwd per GPU"
                               + ("" if with_src else " (flow-grad only)"),
                   "per_gpu_pixels": P, "grads": "flow,src_rgb,src_layout" if with_src else "flow",
                   "l2_policy": "inputs (%.0f MB/step) exceed the 126 MB L2; two input sets alternate" % (P * 120 / 1e6),
                   "parallelism": f"dp{world}", "launch": "cuda graph replay" if graphs is not None else "direct launches",
                   "exchange": ("one NCCL all-reduce of the 8-float loss vector per step, issued after pass 1 and overlapped with pass 2"
                                if world > 1 else "none (single GPU)")},
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": P * dom_bytes,
                     "algorithmic_bytes_per_px": dom_bytes, "kernel_ms": dom_ms},
        "roofline_step": {"algorithmic_bytes_per_px": step_bytes,
                          "achieved": P * step_bytes / (ms_per_step * 1e-3) / 1e9,
                          "frac": P * step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                          "kernel_ms": {"pass1_stage(memset+count+rgb_strip+lay_tile)": k1, "reduce(standalone)": kr, "pass2_rec_kernel": k2},
                          "kernels_us_cupti": kernels_us},
        "e2e": {"value": world * P / (t_e2e.item() * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": _cabi.LOSS_SLOTS * 4, "ms_per_step": t_e2e.item(),
                "inputs": "pinned host: src_rgb, tgt_rgb, flow (fp32), tgt_label (int64), source layout as the dataset's float32 "
                          "class-id map (src/folder.py:97-99), one-hot encoded on the device (src/models/net_utils.py:14-24); "
                          "double-buffered: step i+1 uploads on a copy stream while step i computes"},
        "gpu_launches": launches,
        "clocks": clocks,
        "train": train,
        "torch_cuda_eager": eager,
        "aux": aux,
    }
    if aux:
        for k_ in aux:
            aux[k_]["frac_of_hbm_peak"] = aux[k_]["achieved_GBs"] / peak
    if world == 1 and not args.no_cpu_baseline:
        n_sample = min(N, 4)
        v, t = cpu_oracle_throughput(N, H, W, K, sigma, far, n_sample, iters=5)
        out["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{n_sample}x{H}x{W} slice of the batch, torch CPU oracle port fwd+bwd, median of 5 ({t*1e3:.0f} ms/iter)"}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
