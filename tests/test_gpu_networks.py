"""GPU parity tests at network / loop level: operator path and fused engine against the CPU oracle.
Tolerance: 1e-3 relative fp32 per element, |a-b| <= 1e-3 (|b| + rms(b))  (BASELINE.json north star,
SURVEY.md 8c-iii); coordinates / indices bit-exact where the inputs are identical."""
import numpy as np
import pytest
import torch

from conftest import make_scan
from oracle import me_cpu as ome
from oracle.pipeline import DiffCompletionOracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs() / (b.abs() + b.pow(2).mean().sqrt() + 1e-30)).max().item()


@pytest.fixture(scope="module", params=[1, 0], ids=["ffma", "auto_tc"])
def setup(request, small_scan, calibrated_sds):
    algo = request.param
    from lidiff_b200.pipeline import DiffCompletion
    g = torch.Generator().manual_seed(1234)
    start = torch.randn(small_scan.shape, generator=g)
    noise = torch.randn((3,) + tuple(small_scan.shape), generator=g)
    oracle = DiffCompletionOracle(calibrated_sds["enc"], calibrated_sds["diff"], calibrated_sds["refine"], denoising_steps=50)
    hp = {"data": {"num_points": small_scan.shape[1]}}
    pipe = DiffCompletion(state_dicts=calibrated_sds, denoising_steps=50, cond_weight=6.0, device=DEV, hparams=hp, engine=False, conv_algo=algo)
    return dict(algo=algo, scan=small_scan, start=start, noise=noise, oracle=oracle, pipe=pipe, sds=calibrated_sds)


def test_operator_path_networks_match_oracle(setup):
    """MinkGlobalEnc, MinkUNetDiff (cond + uncond) and MinkUNet(refine) through the ME operator surface"""
    o, pipe, scan = setup["oracle"], setup["pipe"], setup["scan"]
    x_noisy = scan + setup["start"]
    t = torch.tensor([999])
    ot, oc, ou = o.points_to_tensor(x_noisy), o.points_to_tensor(scan), o.points_to_tensor(torch.zeros_like(scan))
    eps_ref = o.classfree_forward(ot, oc, ou, t)
    with torch.no_grad():
        xt, xc, xu = pipe.points_to_tensor(x_noisy.to(DEV)), pipe.points_to_tensor(scan.to(DEV)), pipe.points_to_tensor(torch.zeros_like(scan).to(DEV))
        assert torch.equal(xt.C.cpu(), ot.C), "TensorField coordinates (torch CUDA round(x/res) vs oracle 'mul' mode)"
        enc = pipe.partial_enc(xc)
        eps = pipe.classfree_forward(xt, xc, xu, t.to(DEV))
    # encoder output of the conditional branch (oracle trace holds the last = uncond run; recompute)
    enc_ref = o.enc.global_enc(oc)
    assert torch.equal(enc.C.cpu(), enc_ref.C)
    e = rel_err(enc.F, enc_ref.F)
    print("encoder rel err", e)
    assert e < TOL
    e = rel_err(eps, eps_ref)
    print("guided eps rel err (operator path)", e)
    assert e < TOL
    post = o.points_to_tensor(scan + 0.05 * setup["start"])
    ref_off = o.refine.unet_refine(post)
    with torch.no_grad():
        off = pipe.refine_forward(pipe.points_to_tensor((scan + 0.05 * setup["start"]).to(DEV)))
    e = rel_err(off, ref_off)
    print("refine offsets rel err", e)
    assert e < TOL


def test_engine_step_matches_oracle(setup):
    """fused engine: geometry bit-exact, per-level features, eps, x_next over 3 steps (1st + 2nd order)"""
    from lidiff_b200.engine import DenoiseEngine
    o, scan, sds = setup["oracle"], setup["scan"], setup["sds"]
    N = scan.shape[1]
    x_feats = scan + setup["start"]
    ot, oc, ou = o.points_to_tensor(x_feats), o.points_to_tensor(scan), o.points_to_tensor(torch.zeros_like(scan))
    o.completion_loop(scan, ot, oc, ou, setup["noise"], n_steps=3)
    hist = o.trace["hist"]

    eng = DenoiseEngine(sds["enc"], sds["diff"], device=DEV, n_points=N, denoising_steps=50, conv_algo=setup["algo"])
    x_init = scan.reshape(-1, 3).to(DEV)
    eng.set_condition(x_init)
    xa = x_feats.reshape(-1, 3).float().to(DEV).contiguous()
    ca = torch.zeros(N, 4, device=DEV)
    ca[:, 1:] = ome.quantize(xa.cpu(), 0.05, "mul").to(DEV)
    x0s = torch.zeros(N, 3, dtype=torch.float64, device=DEV)
    for i in range(3):
        xb, cb, eps_out = torch.empty_like(xa), torch.empty_like(ca), torch.empty(N, 3, device=DEV)
        # run the step from the ORACLE's x_t so that each step is compared on identical inputs
        eng.step(i, xa, xb, ca, cb, x_init, setup["noise"][i][0].to(DEV).contiguous(), x0s, eps_out)
        e = rel_err(eps_out, hist[i]["eps"][0])
        print(f"step {i}: guided eps rel err {e:.3e}; level sizes {eng.geom.sizes()}")
        assert e < TOL
        xe = (xb.double().cpu() - hist[i]["x_next"][0]).abs().max().item()
        print(f"step {i}: max |x_next - oracle| = {xe:.3e}")
        assert xe < 1e-3
        # continue from the oracle's state (identical inputs for the next step)
        xa = hist[i]["x_next"][0].float().to(DEV).contiguous()
        ca = torch.zeros(N, 4, device=DEV)
        ca[:, 1:] = ome.quantize(hist[i]["x_next"][0].float(), 0.05, "mul").to(DEV)
    assert eng.h.read_status() == 0


def test_engine_geometry_bit_exact(setup):
    from lidiff_b200.engine import Geometry
    from lidiff_b200 import _lib
    scan = setup["scan"]
    x = (scan + 0.3 * setup["start"]).float()
    of = ome.TensorField(x[0], torch.cat([torch.zeros(x.shape[1], 1), ome.quantize(x[0], 0.05, "mul")], 1))
    og = of.sparse().geom
    h = _lib.get_handle(DEV)
    N = x.shape[1]
    g = Geometry(h, N)
    coords = torch.cat([torch.zeros(N, 1), ome.quantize(x[0], 0.05, "mul")], 1).to(DEV).contiguous()
    g.build(coords, N)
    sizes = g.sizes()
    for l in range(5):
        ref = og.stride_level(1 << l)
        assert sizes[l] == ref.shape[0]
        assert torch.equal(g.C[l][:sizes[l]].cpu(), torch.from_numpy(ref))
        nbr = g.nbr3[l][:, :sizes[l]].cpu().numpy()
        for k, (i_rows, o_rows) in enumerate(og.kernel_map(1 << l, 3, 1, False)):
            got_o = np.nonzero(nbr[k] >= 0)[0]
            assert np.array_equal(got_o, o_rows) and np.array_equal(nbr[k][got_o], i_rows), (l, k)
    assert torch.equal(g.inv[0].long().cpu(), torch.from_numpy(og.inverse))


def test_pipeline_paths_agree_and_complete_scan_runs(setup):
    """operator path and fused engine agree over a short trajectory; refinement + 6x upsampling run"""
    from lidiff_b200.pipeline import DiffCompletion
    scan, sds = setup["scan"], setup["sds"]
    hp = {"data": {"num_points": scan.shape[1]}}
    g = torch.Generator().manual_seed(7)
    start = torch.randn(scan.shape, generator=g)
    noise = torch.randn((4, 1) + tuple(scan.shape[1:]), generator=g).to(DEV)
    pipe = DiffCompletion(state_dicts=sds, denoising_steps=4, cond_weight=6.0, device=DEV, hparams=hp, engine=True, conv_algo=setup["algo"])
    dscan = scan.to(DEV)
    x_feats = dscan + start.to(DEV)
    a = pipe.completion_loop(dscan, pipe.points_to_tensor(x_feats), pipe.points_to_tensor(dscan),
                             pipe.points_to_tensor(torch.zeros_like(dscan)), noise)
    b = pipe.engine().run(dscan, x_feats, noise)
    d = np.abs(a - b).max(1)
    print(f"operator path vs engine after 4 steps: median {np.median(d):.2e}, 99% {np.quantile(d, 0.99):.2e}, max {d.max():.2e}")
    # Two fp32 paths over a 4-step (250 timesteps per step) trajectory: a point whose coordinate sits within rounding distance of a
    # voxel boundary lands in another voxel on one path, which changes that voxel's mean feature and then the trajectory of its
    # members.  Stated bound: half of the points within 5e-5 m, 90 % within 1 mm, fewer than 1 % further than 1 cm apart.
    # (Per-step agreement on identical inputs is asserted in test_engine_step_matches_oracle.)
    assert np.median(d) < 5e-5 and np.quantile(d, 0.9) < 1e-3 and (d > 1e-2).mean() < 0.01
    refined, post = pipe.complete_scan(scan, start_noise=start, step_noise=noise, preprocessed=True)
    assert refined.shape[0] == 6 * post.shape[0] and np.isfinite(refined).all()
