"""multimodal_neuroimage_b200 -- B200-native attention hot path of
Transconnectome/multimodal_neuroimage (shifted-window and cross-modal attention).

Layout:  csrc/ (CUDA kernels + C ABI, built into libmmn_b200.so), _lib.py (ctypes binding),
ops.py (torch.library ops), geometry.py (integer index maps), modules/ (drop-in mirror of
the reference's modules/ package), install.py (swap the reference's classes for ours).
"""
from . import geometry  # noqa: F401

__all__ = ["geometry", "build", "install"]


def build(force: bool = False, verbose: bool = False) -> str:
    from . import _lib
    return _lib.build(force=force, verbose=verbose)


def install(*args, **kwargs):
    from .install import install as _install
    return _install(*args, **kwargs)
