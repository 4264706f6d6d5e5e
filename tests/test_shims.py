"""Import shims of SURVEY.md 8f-1: pytorch_lightning / natsort / open3d stand-ins the reference's inference script needs."""
import os
import sys

import numpy as np
import pytest
import torch

import lidiff_b200.shims as sh


@pytest.fixture(scope="module", autouse=True)
def _shims_on_path():
    sh.install()
    for m in ("open3d", "natsort", "pytorch_lightning"):          # a real install (none in this image) would be shadowed on purpose
        sys.modules.pop(m, None)
    yield


def test_natsorted_orders_numbers_naturally():
    from natsort import natsorted
    assert natsorted(["10.ply", "9.ply", "000100.ply", "1.ply", "a2", "a10"]) == ["1.ply", "9.ply", "10.ply", "000100.ply", "a2", "a10"]
    assert natsorted(["b3", "b20"], reverse=True) == ["b20", "b3"]


def test_lightning_module_hparams_device_and_partial_state_dict(tmp_path):
    from pytorch_lightning.core.lightning import LightningModule
    import yaml

    class M(LightningModule):
        def __init__(self, hp):
            super().__init__()
            self.save_hyperparameters(hp)
            self.lin = torch.nn.Linear(3, 2)

    m = M({"diff": {"t_steps": 1000}, "data": {"resolution": 0.05}})
    assert m.hparams["diff"]["t_steps"] == 1000 and m.device == torch.device("cpu")
    m.hparams["data"]["max_range"] = 50.0                                   # the reference mutates and dumps it (pipeline:49-56)
    assert yaml.safe_load(yaml.dump(m.hparams))["data"] == {"resolution": 0.05, "max_range": 50.0}
    sd = {"lin.weight": torch.ones(2, 3), "somebody.else": torch.zeros(1)}   # Lightning ckpt dicts hold all sub-models: strict=False
    res = m.load_state_dict(sd, strict=False)
    assert "lin.bias" in res.missing_keys and "somebody.else" in res.unexpected_keys and bool((m.lin.weight == 1).all())
    import pytorch_lightning as pl
    with pytest.raises(NotImplementedError):
        pl.Trainer(gpus=1)


def test_open3d_pointcloud_and_ply_round_trip(tmp_path):
    import open3d as o3d
    g = np.random.default_rng(0)
    pts = g.normal(size=(257, 3)) * [5, 5, 0.01]                               # a thin slab: normals ~ +-z
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(pts)
    assert np.array(pcd.points).shape == (257, 3) and not pcd.has_normals()
    for ascii_ in (False, True):
        f = str(tmp_path / f"a{int(ascii_)}.ply")
        o3d.io.write_point_cloud(f, pcd, write_ascii=ascii_)
        back = np.array(o3d.io.read_point_cloud(f).points)
        assert np.allclose(back, pts, rtol=0, atol=0 if not ascii_ else 1e-8)
    pcd.estimate_normals()
    n = np.asarray(pcd.normals)
    assert pcd.has_normals() and np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-5) and (np.abs(n[:, 2]) > 0.9).mean() > 0.9
    f = str(tmp_path / "n.ply")
    o3d.io.write_point_cloud(f, pcd)
    back = o3d.io.read_point_cloud(f)
    assert back.has_normals() and np.allclose(np.asarray(back.normals), n)
    with pytest.raises(RuntimeError):
        o3d.utility.Vector3dVector(np.zeros((4, 2)))
    from lidiff_b200.synth import read_ply_xyz                                   # the package's own reader agrees with the shim's writer
    assert np.allclose(read_ply_xyz(f), pts)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_open3d_fps_fails_loudly_without_gpu():
    import open3d as o3d
    pcd = o3d.geometry.PointCloud(np.zeros((10, 3)))
    with pytest.raises(RuntimeError):
        pcd.farthest_point_down_sample(3)


@pytest.mark.skipif(not os.path.exists("/root/reference/lidiff/tools/diff_completion_pipeline.py"), reason="reference tree not mounted")
def test_reference_pipeline_script_imports_on_the_shims():
    """the reference's own inference script resolves every import against the shims and defines its classes unchanged"""
    import importlib
    sys.path.insert(0, "/root/reference")
    try:
        for m in [k for k in sys.modules if k == "lidiff" or k.startswith("lidiff.")]:
            sys.modules.pop(m)
        mod = importlib.import_module("lidiff.tools.diff_completion_pipeline")
        from pytorch_lightning.core.lightning import LightningModule
        assert issubclass(mod.DiffCompletion, LightningModule)
        assert mod.o3d.__version__.endswith("lidiff_b200.shim") and callable(mod.natsorted)
        assert mod.minknet.ME.__name__ == "MinkowskiEngine"
    finally:
        sys.path.remove("/root/reference")
        for m in [k for k in sys.modules if k == "lidiff" or k.startswith("lidiff.")]:
            sys.modules.pop(m)


@pytest.mark.skipif(not os.path.exists("/root/reference/lidiff/utils/metrics.py"), reason="reference tree not mounted")
def test_reference_metrics_module_runs_on_the_open3d_shim():
    """/root/reference/lidiff/utils/metrics.py (RMSE, ChamferDistance, PrecisionRecall, CompletionIoU) unchanged on the
    open3d shim; the nearest-neighbour distances it builds on are checked against scipy's exact k-d tree"""
    import importlib
    from scipy.spatial import cKDTree
    import open3d as o3d
    sys.path.insert(0, "/root/reference")
    try:
        for m in [k for k in sys.modules if k == "lidiff" or k.startswith("lidiff.")]:
            sys.modules.pop(m)
        metrics = importlib.import_module("lidiff.utils.metrics")
        metrics.torch = torch                                     # the module uses `torch.Tensor` without importing torch
        g = np.random.default_rng(3)
        gt = g.normal(size=(6000, 3)) * [12, 12, 1.0]
        pred = gt[g.choice(6000, 4000, replace=False)] + g.normal(size=(4000, 3)) * 0.05
        pg, pp = o3d.geometry.PointCloud(gt), o3d.geometry.PointCloud(pred)
        d_pg = np.asarray(pp.compute_point_cloud_distance(pg))
        d_gp = np.asarray(pg.compute_point_cloud_distance(pp))
        assert np.allclose(d_pg, cKDTree(gt).query(pred)[0], rtol=0, atol=1e-9) and np.allclose(d_gp, cKDTree(pred).query(gt)[0], rtol=0, atol=1e-9)
        cd, rm = metrics.ChamferDistance(), metrics.RMSE()
        cd.update(pg, pp); rm.update(pg, pp)
        assert abs(cd.compute()[0] - 0.5 * (d_pg.mean() + d_gp.mean())) < 1e-12 and abs(rm.compute()[0] - d_pg.mean()) < 1e-12
        pr = metrics.PrecisionRecall(0.05, 1.0, 20)
        pr.update(pg, pp)
        p, r, f1, t = pr.compute_at_threshold(0.1)
        assert abs(p - 100.0 * (d_pg < t).mean()) < 1e-9 and abs(r - 100.0 * (d_gp < t).mean()) < 1e-9 and 0 < f1 <= 100
        assert all(0 <= v <= 100.000001 for v in pr.compute_auc())            # percentages, normalised by the perfect predictor
        iou = metrics.CompletionIoU(voxel_sizes=[2.0, 1.0, 0.5])       # (the default 0.1 m grid is a 1000^3 float64 histogram: 8 GB)
        iou.update(pg, pp)
        res = iou.compute()
        assert set(res) == {2.0, 1.0, 0.5} and 0 < res[0.5] <= res[1.0] <= res[2.0] <= 1
        assert not metrics.Metrics3D().prediction_is_empty(pp) and metrics.Metrics3D().prediction_is_empty(np.zeros((0, 3)))
        assert metrics.Metrics3D.convert_to_pcd(pred).__class__ is o3d.geometry.PointCloud
        # viewpoint mask of the training collation (collations.py:44-50): voxel-grid membership at 10 m
        grid = o3d.geometry.VoxelGrid.create_from_point_cloud(pp, voxel_size=10.0)
        inc = np.array(grid.check_if_included(o3d.utility.Vector3dVector(gt)))
        org = pred.min(0) - 5.0
        keys = {tuple(k) for k in np.floor((pred - org) / 10.0).astype(int)}
        assert np.array_equal(inc, np.array([tuple(k) in keys for k in np.floor((gt - org) / 10.0).astype(int)]))
    finally:
        sys.path.remove("/root/reference")
        for m in [k for k in sys.modules if k == "lidiff" or k.startswith("lidiff.")]:
            sys.modules.pop(m)
