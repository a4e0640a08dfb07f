"""GPU tests of the training step around the fused op (shape of src/trainer.py:168-258):
producer -> WarpLoss (torch.autograd.Function over the C ABI) -> backward -> optimiser, and its data-parallel
conventions (src/trainer.py:113 DDP, :381-386 Trainer.sync).

  * the fused op's d_flow against central finite differences of its own forward (directional derivatives);
  * a two-shard step under the `global` convention (vlg_problem_t.global_N) equals the one-rank step on the
    concatenated batch: loss vector and d_flow to 1e-6, producer gradients to 5e-5 (cuDNN's own reduction order) -- what a
    2-rank DDP run computes, emulated in one process on one GPU (shards run back to back, gradients SUMMED as NCCL's
    all-reduce would);
  * the `reference` convention (local means, averaged by sync) reproduces Trainer.sync;
  * one optimiser step through train.run_training's step shape reduces the loss and agrees with the same step driven
    by the oracle composition in torch CUDA autograd.
"""
import numpy as np
import pytest
import torch

import vlg_b200
from vlg_b200 import _cabi, parallel
from vlg_b200.producer import FlowGridNet, flow_nhw2
from conftest import parity_errors
from oracle import torch_oracle as TO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _batch(N, H, W, K, seed):
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    f1 = ((torch.rand(N, 3, H, W, generator=g) - mean) / std)
    f2 = ((torch.rand(N, 3, H, W, generator=g) - mean) / std)
    f3 = ((torch.rand(N, 3, H, W, generator=g) - mean) / std)
    blk = 8
    lab = lambda: torch.randint(0, K, (N, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g).repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()
    seg1, seg2, seg3 = lab(), lab(), lab()
    cl = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)
    return dict(frame1=cl(f1), frame2=cl(f2), frame3=cl(f3), seg1=seg1.to(DEV), seg2=seg2.to(DEV), seg3=seg3.to(DEV))


def _net_input(b):
    # cat([seg1, frame1, frame2, seg2]) as src/trainer.py:461 (class-id maps as float channels, src/folder.py:97-99)
    return torch.cat([b["seg1"][:, None].float(), b["frame1"], b["frame2"], b["seg2"][:, None].float()], 1).contiguous(memory_format=torch.channels_last)


def _small_net(seed=0):
    torch.manual_seed(seed)
    net = FlowGridNet(in_channels=8, widths=(8, 12, 16), cols=2, max_flow=2.5).to(DEV)
    with torch.no_grad():      # the flow head starts at zero (producer.py): give it a signal so that gradients flow
        for p in net.flow_head.parameters():
            p.add_(0.05 * torch.randn_like(p))
    return net


def _grads(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])


def test_d_flow_matches_finite_differences_of_the_forward():
    """gradcheck-style: <d_flow, v> against (L(f + eps v) - L(f - eps v)) / (2 eps) of the op's own forward, for random
    directions v.  fp32 forward => a loose bar; kinks of the bilinear interpolant / |.| terms add O(eps) errors."""
    N, H, W, K = 2, 24, 40, 20
    g = torch.Generator().manual_seed(5)
    b = _batch(N, H, W, K, seed=3)
    layout = torch.randn(N, K, H, W, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    flow = (torch.rand(N, H, W, 2, generator=g) * 0.8 + 0.1).to(DEV)     # fractional parts away from the cell borders
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))
    f = flow.clone().requires_grad_(True)
    crit(b["frame2"], layout, f, b["frame3"], b["seg3"]).backward()
    d_flow = f.grad.double()
    eps = 4e-3
    for trial in range(4):
        v = torch.randn(N, H, W, 2, generator=g).to(DEV)
        if trial < 2:      # aligned with the gradient: no cancellation in <d_flow, v>
            v = v.abs() * torch.sign(d_flow).float()
        with torch.no_grad():
            lp = crit(b["frame2"], layout, flow + eps * v, b["frame3"], b["seg3"])
            lm = crit(b["frame2"], layout, flow - eps * v, b["frame3"], b["seg3"])
        fd = (lp.double() - lm.double()).item() / (2 * eps)
        an = (d_flow * v.double()).sum().item()
        scale = (d_flow.abs() * v.double().abs()).sum().item()      # what the terms of <d_flow, v> add up to without cancellation
        assert abs(fd - an) <= 3e-2 * scale, (trial, fd, an, scale)
        if trial < 2:
            assert abs(fd - an) <= 3e-2 * abs(an), (trial, fd, an)


@pytest.mark.parametrize("layout_kind", ["onehot", "label"])
def test_two_shard_global_step_equals_single_rank_step(layout_kind):
    """Rank r of a 2-rank job processes samples [r*B, (r+1)*B) with global_batch = 2B; NCCL sums the loss vectors and DDP
    sums the gradients.  Emulated on one GPU: the summed shard results must equal the one-rank step on the 2B batch.
    The producer's flow is computed ONCE for the 2B batch and sliced, so that the comparison is about the fused op and the
    data-parallel convention, not about cuDNN picking another convolution algorithm for another batch size."""
    B, H, W, K = 2, 48, 80, 20
    b = _batch(2 * B, H, W, K, seed=11)
    net = _small_net()
    w = (40.0, 20.0, 10.0, 0.5)
    x = _net_input(b)
    src_lay = b["seg2"] if layout_kind == "label" else vlg_b200.one_hot_layout(b["seg2"], K)

    # one rank, 2B samples
    net.zero_grad(set_to_none=True)
    flow, _ = net(x)
    f_full = flow_nhw2(flow)
    f_full.retain_grad()
    crit = vlg_b200.WarpLoss(weights=w)
    crit(b["frame2"], src_lay, f_full, b["frame3"], b["seg3"]).backward()
    vec_full, dflow_full, g_full = crit.last_terms.clone(), f_full.grad.clone(), _grads(net).clone()

    # two ranks, B samples each, divisors of the GLOBAL batch
    vecs, dflows = [], []
    for sl in (slice(0, B), slice(B, 2 * B)):
        f = f_full.detach()[sl].clone().requires_grad_(True)
        crit_g = vlg_b200.WarpLoss(weights=w, global_batch=2 * B)
        crit_g(b["frame2"][sl], src_lay[sl], f, b["frame3"][sl], b["seg3"][sl]).backward()
        vecs.append(crit_g.last_terms.clone())
        dflows.append(f.grad.clone())
    vec_sum = vecs[0] + vecs[1]
    np.testing.assert_allclose(vec_sum[:_cabi.LOSS_NVALID + 1].cpu().numpy(), vec_full[:_cabi.LOSS_NVALID + 1].cpu().numpy(), rtol=1e-6)
    mx, rms = parity_errors(torch.cat(dflows, 0), dflow_full)
    assert mx <= 1e-6 and rms <= 1e-6, (mx, rms)
    # what DDP's gradient all-reduce then adds up: the producer gradients driven by the shard d_flows
    net.zero_grad(set_to_none=True)
    flow2, _ = net(x)
    flow_nhw2(flow2).backward(gradient=torch.cat(dflows, 0))
    # d_flow agrees to 1e-6; the weight gradients are sums over all pixels with heavy cancellation (and cuDNN's backward
    # kernels reduce in their own order), which amplifies that to a few 1e-5 of the largest weight gradient
    mx, rms = parity_errors(_grads(net), g_full)
    assert mx <= 5e-5 and rms <= 5e-5, (mx, rms)
    assert g_full.abs().max().item() > 0


def test_reference_convention_reproduces_trainer_sync():
    """src/trainer.py:248-256,381-386: every rank normalises by its LOCAL batch and sync() averages the scalars.  With
    equal shard sizes the averaged L1 / GD / SSIM / TV terms equal the one-rank means; CE is the mean of the per-shard
    means (all labels valid here, so that is the global mean too)."""
    B, H, W, K = 2, 40, 64, 20
    b = _batch(2 * B, H, W, K, seed=19)
    g = torch.Generator().manual_seed(2)
    flow = (torch.randn(2 * B, H, W, 2, generator=g) * 1.5).to(DEV)
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))
    lay = vlg_b200.one_hot_layout(b["seg2"], K)
    with torch.no_grad():
        crit(b["frame2"], lay, flow, b["frame3"], b["seg3"])
        full = crit.last_terms.clone()
        parts = []
        for sl in (slice(0, B), slice(B, 2 * B)):
            crit(b["frame2"][sl], lay[sl], flow[sl], b["frame3"][sl], b["seg3"][sl])
            parts.append(crit.last_terms.clone())
    avg = (parts[0] + parts[1]) / 2          # what parallel.sync_loss_vector(..., "reference") returns on 2 ranks
    np.testing.assert_allclose(avg[:_cabi.LOSS_TOTAL + 1].cpu().numpy(), full[:_cabi.LOSS_TOTAL + 1].cpu().numpy(), rtol=2e-6)
    # without a process group the helper is the identity (single-GPU runs need no branch)
    assert torch.equal(parallel.sync_loss_vector(full, "reference"), full)


def test_optimiser_step_matches_the_oracle_driven_step():
    """One Adam step of the harness's shape (train.py: producer -> fused op -> backward -> Adam) against the same step
    with the oracle composition in torch CUDA autograd: loss and the op's own gradient (d_flow) to 1e-5.  The producer's
    WEIGHT gradients are sums over all pixels with heavy cancellation, reduced by cuDNN's backward kernels in their own
    order, so two fp32 steps differ by ~1e-5 there whatever the op does: they are judged against the same step in fp64
    (net, oracle and autograd in double) -- the fused op's step must be as close to it as the oracle-driven fp32 step is
    (1e-5, or 1.5x the oracle-driven step's own distance, whichever is larger).  And the loss goes down over a few steps."""
    import copy
    N, H, W, K = 2, 48, 80, 20
    b = _batch(N, H, W, K, seed=23)
    lay = vlg_b200.one_hot_layout(b["seg2"], K)
    x = _net_input(b)
    net = _small_net(seed=1)
    net_ref = _small_net(seed=1)
    net_ref.load_state_dict(net.state_dict())
    net64 = copy.deepcopy(net).double()
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))

    flow, _ = net(x)
    f_op = flow_nhw2(flow)
    f_op.retain_grad()
    loss = crit(b["frame2"], lay, f_op, b["frame3"], b["seg3"])
    loss.backward()

    flow_r, _ = net_ref(x)
    f_ref = flow_nhw2(flow_r)
    f_ref.retain_grad()
    o = TO.warp_loss(b["frame2"], lay, f_ref, b["frame3"], b["seg3"], w_tv=0.5)
    o["total"].backward()
    np.testing.assert_allclose(loss.item(), o["total"].item(), rtol=1e-5)
    mx, rms = parity_errors(f_op.grad, f_ref.grad)
    assert mx <= 1e-5 and rms <= 1e-5, ("d_flow", mx, rms)

    flow_d, _ = net64(x.double())
    o64 = TO.warp_loss(b["frame2"].double(), lay.double(), flow_nhw2(flow_d), b["frame3"].double(), b["seg3"], w_tv=0.5)
    o64["total"].backward()
    g64 = _grads(net64)
    ours, theirs = parity_errors(_grads(net), g64), parity_errors(_grads(net_ref), g64)
    for got, ref, what in zip(ours, theirs, ("normwise-max", "RMS-relative")):
        assert got <= max(1e-5, 1.5 * ref), (what, "fused step vs fp64", got, "oracle-driven fp32 step vs fp64", ref)

    opt = torch.optim.Adam(net.parameters(), lr=2e-3, betas=(0.5, 0.999))      # src/main.py:139-141
    first = None
    for it in range(8):
        opt.zero_grad(set_to_none=True)
        flow, _ = net(x)
        loss = crit(b["frame2"], lay, flow_nhw2(flow), b["frame3"], b["seg3"])
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < first


def test_captured_step_replays_the_eager_step_bitwise():
    """vlg_b200.CapturedStep: the step's public calls (ingest, WarpLoss, backward) recorded once with torch.cuda.graph
    and replayed on NEW data copied into the static buffers give, bit for bit, the loss vector and the gradients of the
    eager call sequence on that data (the kernels are deterministic); and backward() hands its gradient buffers over to
    autograd without a copy (the .grad of a leaf IS the buffer the fused pass wrote)."""
    N, H, W, K = 2, 40, 72, 20
    g = torch.Generator().manual_seed(31)
    mk = lambda: dict(src=torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).to(DEV),
                      tgt=torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).to(DEV),
                      sseg=torch.randint(0, K, (N, H, W), generator=g).to(torch.uint8).to(DEV),
                      tseg=torch.randint(0, K, (N, H, W), generator=g).to(torch.uint8).to(DEV),
                      flow=(torch.randn(N, H, W, 2, generator=g) * 1.5).to(DEV))
    crit = vlg_b200.WarpLoss(weights=(40.0, 20.0, 10.0, 0.5))

    def step(buf):
        src = vlg_b200.ingest(buf["src"], buf["sseg"], n_classes=K, want_label=False, want_one_hot=True)
        tgt = vlg_b200.ingest(buf["tgt"], buf["tseg"], n_classes=K, want_label=True)
        a, b = src["frames"].requires_grad_(True), src["one_hot"].requires_grad_(True)
        f = buf["flow"].detach().requires_grad_(True)
        crit(a, b, f, tgt["frames"], tgt["label"]).backward()
        return crit.last_terms, a.grad, b.grad, f.grad

    static = mk()
    cap = vlg_b200.CapturedStep(lambda: step(static))
    for _ in range(2):
        new = mk()
        want = [t.clone() for t in step(new)]
        for k in static:
            static[k].copy_(new[k])
        got = cap()
        torch.cuda.synchronize()
        for w, x in zip(want, got):
            assert torch.equal(w, x)

    # eager backward: no clone of the gradient buffers
    buf = mk()
    src = vlg_b200.ingest(buf["src"], buf["sseg"], n_classes=K, want_label=False, want_one_hot=True)
    tgt = vlg_b200.ingest(buf["tgt"], buf["tseg"], n_classes=K, want_label=True)
    a, b = src["frames"].requires_grad_(True), src["one_hot"].requires_grad_(True)
    f = buf["flow"].detach().requires_grad_(True)
    total, terms, _ = vlg_b200.warp_loss(a, b, f, tgt["frames"], tgt["label"], crit.cfg)
    ptrs = [t.data_ptr() for t in total.grad_fn.grads]
    total.backward()
    assert [a.grad.data_ptr(), b.grad.data_ptr(), f.grad.data_ptr()] == ptrs
