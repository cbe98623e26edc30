"""Test helper: loads the product's host mirror (hoh-ans_b200/host/hohgpu.py) without the package
name being importable (the directory name carries a hyphen)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hoh-ans_b200")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def hohgpu():
    return _load("hohgpu", os.path.join(PKG, "host", "hohgpu.py"))


def builder():
    return _load("hoh_build", os.path.join(PKG, "build.py"))


_gpu = None


def gpu():
    """A process-wide HohGpu on cuda:0 (raises when there is no device: no CPU fallback)."""
    global _gpu
    if _gpu is None:
        _gpu = hohgpu().HohGpu(0)
    return _gpu
