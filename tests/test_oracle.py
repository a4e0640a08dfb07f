"""Pins the oracle (oracle/torch_oracle.py, oracle/warp_oracle.c) to the golden fixtures that
tests/golden/make_golden.py produced from the real reference modules (src/loss.py GradientLoss /
SsimLoss; CE / L1 as src/trainer.py:124,130) and torch 2.11.0 CPU grid_sample."""
import numpy as np
import torch

from oracle import c_oracle as CO
from oracle import torch_oracle as TO

TERMS = ["l1", "gd", "ssim", "ce", "tv"]


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def test_torch_oracle_forward_matches_reference_fixture(golden):
    g = golden
    out = TO.warp_loss(_t(g["src_rgb"]), _t(g["src_layout"]), _t(g["flow"]), _t(g["tgt_rgb"]),
                       _t(g["tgt_label"]), w_tv=g["w_tv"], padding_mode=g["padding"])
    assert (_bits(out["grid"].numpy()) == _bits(g["grid"])).all()
    assert (_bits(out["warped_rgb"].numpy()) == _bits(g["warped_rgb"])).all()
    assert (_bits(out["warped_layout"].numpy()) == _bits(g["warped_layout"])).all()
    assert (out["argmax"].numpy() == g["argmax"]).all()
    got = np.array([out["terms"][k].item() for k in TERMS])
    np.testing.assert_allclose(got, g["terms"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(out["total"].item(), g["total"], rtol=2e-6)


def test_torch_oracle_backward_matches_reference_fixture(golden):
    g = golden
    out = TO.warp_loss_fwd_bwd(_t(g["src_rgb"]), _t(g["src_layout"]), _t(g["flow"]), _t(g["tgt_rgb"]),
                               _t(g["tgt_label"]), w_tv=g["w_tv"], padding_mode=g["padding"])
    for k in ("d_src_rgb", "d_src_layout", "d_flow"):
        ref = g[k]
        err = np.abs(out[k].numpy() - ref).max()
        assert err <= 1e-5 * np.abs(ref).max(), (k, err)


def test_torch_oracle_fp64_agrees(golden):
    """fp64 tie-breaker stays within 1e-5 of the fp32 reference terms."""
    g = golden
    out = TO.warp_loss_fwd_bwd(_t(g["src_rgb"]), _t(g["src_layout"]), _t(g["flow"]), _t(g["tgt_rgb"]),
                               _t(g["tgt_label"]), w_tv=g["w_tv"], padding_mode=g["padding"],
                               dtype=torch.float64)
    got = np.array([out["terms"][k].item() for k in TERMS])
    np.testing.assert_allclose(got, g["terms"], rtol=1e-5, atol=1e-7)


def test_c_oracle_indices_and_warp_bitwise(golden):
    g = golden
    assert (_bits(CO.flow_to_grid(g["flow"])) == _bits(g["grid"])).all()
    for name in ("rgb", "layout"):
        out = CO.warp_fwd(g["src_" + name], g["grid"], g["padding"])
        assert (_bits(out) == _bits(g["warped_" + name])).all(), name
    assert (CO.argmax(g["warped_layout"]) == g["argmax"]).all()


def test_c_oracle_loss_terms(golden):
    g = golden
    lab = g["tgt_label"]
    got = np.array([CO.l1(g["warped_rgb"], g["tgt_rgb"]), CO.gd(g["warped_rgb"], g["tgt_rgb"]),
                    CO.ssim(g["warped_rgb"], g["tgt_rgb"]), CO.ce(g["warped_layout"], lab),
                    CO.tv(g["flow"])])
    np.testing.assert_allclose(got, g["terms"], rtol=3e-6, atol=0)


def test_c_oracle_sample_coords_consistent(golden):
    g = golden
    ixy, x0y0, w4 = CO.sample_coords(g["grid"], g["padding"])
    N, H, W, _ = g["grid"].shape
    assert (x0y0 == np.floor(ixy).astype(np.int32)).all()
    if g["padding"] == "border":
        assert ixy[..., 0].min() >= 0 and ixy[..., 0].max() <= W - 1
        assert ixy[..., 1].min() >= 0 and ixy[..., 1].max() <= H - 1
    np.testing.assert_allclose(w4.sum(-1), 1.0, atol=1e-5)


def test_identity_grid_is_not_exact():
    """under the reference grid convention the identity grid misses the
    integer on a sizeable fraction of columns; the oracle must reproduce those off-by-one floors."""
    for W in (256, 512, 1242, 2048):
        grid = CO.base_grid(1, 4, W)
        ixy, x0y0, _ = CO.sample_coords(grid, "border")
        frac_exact = (x0y0[0, 0, :, 0] == np.arange(W)).mean()
        assert 0.6 < frac_exact < 0.95, (W, frac_exact)
        t = TO.base_grid(1, 4, W).numpy()
        assert (_bits(t) == _bits(grid)).all()


def test_argmax_first_max_tiebreak():
    x = np.zeros((1, 3, 1, 2), np.float32)
    x[0, :, 0, 0] = [0.5, 0.5, 0.1]
    x[0, :, 0, 1] = [0.1, 0.7, 0.7]
    assert CO.argmax(x).tolist() == [[[0, 1]]]
    assert TO.argmax_layout(torch.from_numpy(x)).tolist() == [[[0, 1]]]


def test_c_oracle_backward_close_to_torch(golden):
    g = golden
    src = _t(g["src_layout"]).clone().requires_grad_(True)
    grid = _t(g["grid"]).clone().requires_grad_(True)
    out = TO.warp(src, grid, g["padding"])
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(7))
    out.backward(go)
    ds, dg = CO.warp_bwd(g["src_layout"], g["grid"], go.numpy(), g["padding"])
    assert np.abs(ds - src.grad.numpy()).max() <= 1e-5 * np.abs(ds).max()
    assert np.abs(dg - grid.grad.numpy()).max() <= 1e-5 * np.abs(dg).max()


def test_renorm_oracle_follows_trainer_expressions():
    """oracle.renorm_frames restates src/trainer.py:122-123,193-195,200-206,215 -- the tensors are built exactly
    as the trainer builds them ([None,:,None,None] broadcasts of the same constants)."""
    import torch
    from oracle import torch_oracle as TO
    g = torch.Generator().manual_seed(2)
    frame = torch.rand(2, 3, 9, 14, generator=g)
    seg = torch.randint(0, 20, (2, 9, 14), generator=g)
    img_std_arr = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    img_mean_arr = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
    want = (frame - img_mean_arr) / img_std_arr
    assert torch.equal(TO.renorm_frames(frame), want)
    out, lab = TO.renorm_frames(frame, flip=True, labels=seg)
    assert torch.equal(out, torch.flip(want, [3])) and torch.equal(lab, torch.flip(seg, [2]))
    assert torch.equal(TO.renorm_frames(want, denormalize=True), want * img_std_arr + img_mean_arr)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: one JSON line with the
    contract's keys, our arm's workload description, and a cpu_baseline describing the run."""
    import json
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["value"] > 0
    assert d["config"]["workload"].startswith("c1: 2x128x256") and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
