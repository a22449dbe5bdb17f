"""Swap the reference's attention modules for the CUDA-backed ones, in place.

`install()` registers our `modules.*` files under the reference's module names in
`sys.modules` (`modules.swin_v2_module`, `modules.swinfusion_module`,
`modules.crossmodal_transformer`, `modules.multihead_attention`,
`modules.position_embedding`), so the reference's `model.py` -- which does
`from modules.swin_v2_module import *`, `from modules.swinfusion_module import *` and
`from modules.crossmodal_transformer import TransformerEncoder` (model.py:8,15,18) --
picks them up unchanged; trainer.py and main.py need no edits.  Call it before importing
the reference's `model`.  If `model` is already imported its symbols are rebound too.
"""
from __future__ import annotations

import importlib
import sys
import types

_NAMES = ["swin_v2_module", "swinfusion_module", "crossmodal_transformer", "multihead_attention", "position_embedding"]


def install(rebind_model: bool = True) -> None:
    pkg = sys.modules.get("modules")
    if pkg is None:
        pkg = types.ModuleType("modules")
        pkg.__path__ = []                       # a package with no filesystem search path
        sys.modules["modules"] = pkg
    for name in _NAMES:
        ours = importlib.import_module(f"multimodal_neuroimage_b200.modules.{name}")
        sys.modules[f"modules.{name}"] = ours
        setattr(pkg, name, ours)
    if rebind_model and "model" in sys.modules:
        model = sys.modules["model"]
        for name in ("swin_v2_module", "swinfusion_module"):
            ours = sys.modules[f"modules.{name}"]
            for sym in dir(ours):
                if not sym.startswith("_") and hasattr(model, sym) and isinstance(getattr(ours, sym), type):
                    setattr(model, sym, getattr(ours, sym))
        model.TransformerEncoder = sys.modules["modules.crossmodal_transformer"].TransformerEncoder
