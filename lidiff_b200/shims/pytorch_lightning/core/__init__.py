from .lightning import LightningModule  # noqa: F401
