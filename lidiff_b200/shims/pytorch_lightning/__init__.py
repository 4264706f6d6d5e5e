"""Minimal stand-in for pytorch_lightning 1.5: what the reference's inference scripts touch —
`LightningModule` as an nn.Module with `save_hyperparameters` / `.hparams` / `.device`
(/root/reference/lidiff/tools/diff_completion_pipeline.py:7,15-19,69).  Training orchestration (`Trainer`, DDP,
checkpoint callbacks; train.py:88-121) is out of scope and raises."""
from .core.lightning import LightningModule  # noqa: F401


class Trainer:
    def __init__(self, *a, **k):
        raise NotImplementedError("lidiff_b200 shims pytorch_lightning for inference only; training (SURVEY.md 8f-3) is not built")


__version__ = "1.5.10+lidiff_b200.shim"
