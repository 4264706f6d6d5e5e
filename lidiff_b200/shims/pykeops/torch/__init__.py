"""`from pykeops.torch import LazyTensor` -> lidiff_b200.keops"""
from lidiff_b200.keops import LazyTensor  # noqa: F401
