"""Drop-in import shims so the reference's files run unchanged on lidiff_b200."""
import os
import sys


def install():
    """Prepend the shim directory to sys.path (idempotent).  Real installs of the shimmed packages,
    if any, are shadowed on purpose."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
