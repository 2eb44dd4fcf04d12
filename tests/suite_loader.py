"""The GPU parity suites take their forward precision from a module global.  `clone_suite` executes a suite's source a
second time under another precision and hands back its tests and fixtures, so that ONE `pytest -m gpu` process runs the
suites on the fp32 SIMT path and on the tensor-core paths, every test individually visible in the report."""
from __future__ import annotations

import types
from pathlib import Path

HERE = Path(__file__).resolve().parent


def suite_precision(module_globals, default="fp32"):
    """Precision of the suite module being executed: the one clone_suite planted, else the environment's."""
    import os
    return module_globals.get("__suite_precision__") or os.environ.get("TWISTERL_B200_PRECISION", default)


def clone_suite(name: str, precision: str) -> dict:
    path = HERE / f"{name}.py"
    mod = types.ModuleType(f"{name}__{precision}")
    mod.__file__ = str(path)
    mod.__dict__["__suite_precision__"] = precision
    exec(compile(path.read_text(), str(path), "exec"), mod.__dict__)
    keep = {}
    for k, v in mod.__dict__.items():
        if k.startswith("test_") or k == "pytestmark" or hasattr(v, "_fixture_function_marker") or hasattr(v, "_pytestfixturefunction") \
                or type(v).__name__ == "FixtureFunctionDefinition":
            keep[k] = v
    return keep
