"""Loader for the reference's own CUDA extensions built by oracle/build_ref.py
(oracle/_ref/_*.so).  Test infrastructure only."""
import importlib.util
import os

import torch  # noqa: F401  (libtorch symbols must be loaded first)

_REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle", "_ref")
_cache = {}


def load(name):
    """name in {_raymarching, _gridencoder, _shencoder, _freqencoder}; returns the module or None if not built."""
    if name in _cache:
        return _cache[name]
    path = os.path.join(_REF_DIR, name + ".so")
    mod = None
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def require(name):
    import pytest
    m = load(name)
    if m is None:
        pytest.skip(f"reference extension {name} not built (run python oracle/build_ref.py)")
    return m
