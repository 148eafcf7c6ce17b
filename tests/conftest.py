import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build_library():
    """libnmb200.so is git-ignored, and importing the package dlopens it: build it BEFORE any test module is
    collected, loading build.py by path (importing nanomotif_b200.build would import the package first)."""
    spec = importlib.util.spec_from_file_location("_nmb_build", os.path.join(ROOT, "nanomotif_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except RuntimeError as exc:
        if "nvcc not found" in str(exc):
            import pytest

            pytest.exit(f"libnmb200.so is missing and cannot be built here: {exc}", returncode=2)
        raise


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    _build_library()
