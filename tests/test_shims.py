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
