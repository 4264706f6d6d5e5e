"""Host-side logic of the product (operator-surface wiring, engine orchestration, ctypes descriptors and
pointer plumbing) exercised on CPU through tests/fake_backend.py, compared with the oracle.  The fake
replaces only the CUDA library handle; everything above the C ABI is the shipped code."""
import numpy as np
import pytest
import torch

import fake_backend
from conftest import make_scan
from oracle import me_cpu as ome
from oracle.pipeline import DiffCompletionOracle, calibrated_state_dicts


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs() / (b.abs() + b.pow(2).mean().sqrt() + 1e-30)).max().item()


@pytest.fixture(scope="module")
def tiny():
    scan = make_scan(250, 3)                      # (1, 2500, 3)
    sds = calibrated_state_dicts(scan, seed=5)
    g = torch.Generator().manual_seed(11)
    return dict(scan=scan, sds=sds, start=torch.randn(scan.shape, generator=g),
                noise=torch.randn((3,) + tuple(scan.shape), generator=g))


def test_operator_surface_networks_match_oracle(tiny, monkeypatch):
    fake_backend.install(monkeypatch)
    from lidiff_b200.pipeline import DiffCompletion
    scan, sds = tiny["scan"], tiny["sds"]
    pipe = DiffCompletion(state_dicts=sds, denoising_steps=50, device="cpu", hparams={"data": {"num_points": scan.shape[1]}}, engine=False)
    o = DiffCompletionOracle(sds["enc"], sds["diff"], sds["refine"], div_mode="div")     # torch CPU division is a true division
    x = scan + tiny["start"]
    t = torch.tensor([999])
    ref = o.classfree_forward(o.points_to_tensor(x), o.points_to_tensor(scan), o.points_to_tensor(torch.zeros_like(scan)), t)
    got = pipe.classfree_forward(pipe.points_to_tensor(x), pipe.points_to_tensor(scan), pipe.points_to_tensor(torch.zeros_like(scan)), t)
    assert rel_err(got, ref) < 1e-4
    r_ref = o.refine.unet_refine(o.points_to_tensor(scan))
    r_got = pipe.refine_forward(pipe.points_to_tensor(scan))
    assert rel_err(r_got, r_ref) < 1e-4


@pytest.mark.parametrize("emulate_tc", [False, True], ids=["plain", "scatter_split"])
def test_engine_orchestration_matches_oracle(tiny, monkeypatch, emulate_tc):
    h = fake_backend.install(monkeypatch, emulate_tc=emulate_tc)
    from lidiff_b200.engine import DenoiseEngine
    scan, sds = tiny["scan"], tiny["sds"]
    N = scan.shape[1]
    o = DiffCompletionOracle(sds["enc"], sds["diff"], None, div_mode="mul")
    ref = o.completion_loop(scan, o.points_to_tensor(scan + tiny["start"]), o.points_to_tensor(scan),
                            o.points_to_tensor(torch.zeros_like(scan)), tiny["noise"], n_steps=3)
    hist = o.trace["hist"]
    eng = DenoiseEngine(sds["enc"], sds["diff"], device="cpu", n_points=N, denoising_steps=50, div_mode=1)
    out = eng.run(scan, scan + tiny["start"], tiny["noise"][:, 0], n_steps=3)
    d = np.abs(out - ref).max(1)
    print("engine vs oracle after 3 steps: median", np.median(d), "max", d.max())
    assert np.median(d) < 1e-5 and (d > 1e-3).mean() < 0.02
    assert h.launches > 100
    # single step on identical inputs: eps within tolerance
    eng2 = DenoiseEngine(sds["enc"], sds["diff"], device="cpu", n_points=N, denoising_steps=50, div_mode=1)
    eng2.set_condition(scan.reshape(-1, 3))
    xa = (scan + tiny["start"]).reshape(-1, 3).float().contiguous()
    ca = torch.zeros(N, 4)
    ca[:, 1:] = ome.quantize(xa, 0.05, "mul")
    xb, cb, eps = torch.empty_like(xa), torch.empty_like(ca), torch.empty(N, 3)
    eng2.step(0, xa, xb, ca, cb, scan.reshape(-1, 3).double().contiguous(), tiny["noise"][0, 0].contiguous(),
              torch.zeros(N, 3, dtype=torch.float64), eps)
    assert rel_err(eps, hist[0]["eps"][0]) < 1e-4
    assert np.abs(xb.numpy() - hist[0]["x_next"][0].numpy()).max() < 1e-4


def test_pipeline_complete_scan_host_flow(tiny, monkeypatch):
    fake_backend.install(monkeypatch)
    from lidiff_b200.pipeline import DiffCompletion
    scan, sds = tiny["scan"], tiny["sds"]
    pipe = DiffCompletion(state_dicts=sds, denoising_steps=2, device="cpu", hparams={"data": {"num_points": scan.shape[1]}}, engine=True)
    refined, post = pipe.complete_scan(scan, start_noise=tiny["start"], step_noise=tiny["noise"][:2], preprocessed=True)
    assert refined.shape == (post.shape[0] * 6, 3) and np.isfinite(refined).all()
