"""`pytorch_lightning.core.lightning.LightningModule` for inference: an nn.Module with Lightning's hyper-parameter and
device conveniences (reference use: diff_completion_pipeline.py:15-56,69,119)."""
import torch


class LightningModule(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self._hparams = {}

    def save_hyperparameters(self, *args, **kwargs):
        """Lightning accepts a dict / Namespace / names; the reference passes the checkpoint's `hyper_parameters` dict."""
        for a in args:
            if isinstance(a, dict):
                self._hparams.update(a)
            elif hasattr(a, "__dict__"):
                self._hparams.update(vars(a))
            elif a is not None:
                raise TypeError(f"save_hyperparameters: unsupported argument {type(a).__name__} (shim takes dicts / namespaces)")
        self._hparams.update(kwargs)

    @property
    def hparams(self):
        return self._hparams            # a plain dict: the reference indexes it and yaml.dump()s it

    @property
    def device(self):
        for t in self.parameters():
            return t.device
        for t in self.buffers():
            return t.device
        return torch.device("cpu")

    # hooks Lightning would call; kept so subclasses that define / call them do not break
    def log(self, *a, **k):
        pass

    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
