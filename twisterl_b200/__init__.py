"""twisterl_b200 -- B200-native rollout engine behind twisteRL's `twisterl.twisterl` API.

Submodules mirror the reference's PyO3 module (rust/src/python_interface/python_bindings.rs:55-77):
`env`, `nn`, `collector`.  `install_as_twisterl()` registers this package as `twisterl.twisterl`
(and the example crate's `grid_world`), so the reference's Python half (trainers, configs,
checkpoint loading) runs unmodified on top of the CUDA engine.
"""
from __future__ import annotations

import sys
import types

from . import _lib, collector, env, nn
from ._lib import Engine, configure, default_engine

__all__ = ["env", "nn", "collector", "Engine", "configure", "default_engine", "install_as_twisterl"]
__version__ = "0.1.0"


def install_as_twisterl() -> types.ModuleType:
    """Make `from twisterl import twisterl` / `import grid_world` resolve to this engine."""
    mod = types.ModuleType("twisterl.twisterl")
    mod.__doc__ = "twisterl_b200 standing in for the Rust extension twisterl.twisterl"
    mod.env, mod.nn, mod.collector = env, nn, collector
    sys.modules["twisterl.twisterl"] = mod
    for sub in ("env", "nn", "collector"):
        sys.modules[f"twisterl.twisterl.{sub}"] = getattr(mod, sub)
    gw = types.ModuleType("grid_world")
    gw.GridWorld = env.GridWorld
    sys.modules["grid_world"] = gw
    pkg = sys.modules.get("twisterl")
    if pkg is not None:
        pkg.twisterl = mod
    return mod
