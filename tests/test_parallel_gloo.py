"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding plus the single
loss-vector all-reduce.  The per-rank compute is the oracle (no GPU here); what is under test is
that sharded + reduced == unsharded for both reduction conventions."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vlg_b200  # noqa: F401  (import shim)
from vlg_b200 import parallel
from oracle import torch_oracle as TO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(N=4, H=12, W=20, K=20, seed=1024):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(N, 3, H, W, generator=g), torch.randn(N, K, H, W, generator=g),
            torch.randn(N, H, W, 2, generator=g) * 1.5, torch.randn(N, 3, H, W, generator=g),
            torch.randint(0, K, (N, H, W), generator=g))


def _local_vector(shard, n_global=None):
    """Oracle loss vector of one shard; `n_global` switches to global-batch divisors."""
    a, b, f, t, l = shard
    out = TO.warp_loss(a, b, f, t, l, w_tv=0.5)
    terms = [out["terms"][k] for k in ("l1", "gd", "ssim", "ce", "tv")]
    scale = 1.0 if n_global is None else a.shape[0] / n_global
    terms = [x * scale for x in terms]
    total = 40 * terms[0] + 20 * (terms[1] + terms[2]) + 10 * terms[3] + 0.5 * terms[4]
    nvalid = torch.tensor(float(l.numel()))
    return torch.stack(terms + [total, nvalid, torch.tensor(0.0)]).float()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    data = _case()
    shard = parallel.shard_batch(data, rank, world)
    n_global = data[0].shape[0]
    v_glob = parallel.sync_loss_vector(_local_vector(shard, n_global), "global")
    v_ref = parallel.sync_loss_vector(_local_vector(shard), "reference")
    if rank == 0:
        q.put((v_glob.numpy(), v_ref.numpy()))
    dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    for n in (1, 5, 16, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_loss_vector_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    v_glob, v_ref = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _local_vector(_case()).numpy()
    # global divisors: per-rank vectors add up to the single-process vector
    np.testing.assert_allclose(v_glob[:6], full[:6], rtol=2e-6)
    assert v_glob[6] == full[6]
    # reference convention (equal shards): mean of local means == global mean as well
    np.testing.assert_allclose(v_ref[:6], full[:6], rtol=2e-6)
    assert v_ref[6] == full[6]


def test_sync_is_identity_without_process_group():
    v = torch.arange(8, dtype=torch.float32)
    assert torch.equal(parallel.sync_loss_vector(v), v)
