"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* reference package.

Loads MicrobialDarkMatter/nanomotif unchanged from ``/root/reference`` in an image that
lacks its third-party dependencies (polars, epymetheus, pysam, pyfastx, progressbar,
pyinstrument, snakemake, hdbscan, Bio).  It exists for two purposes only:

* ``tests/golden/generate_golden.py`` runs the reference's pure functions to produce the
  committed golden vectors (the reference tree does not travel to the GPU box);
* ``tests/test_oracle_vs_reference.py`` cross-checks ``oracle/restate.py`` against the real
  functions whenever ``/root/reference`` happens to be mounted (skipped otherwise).

``oracle/minipolars.py`` goes one step further for the functions that take a polars FRAME: after
``load_reference()``, ``minipolars.install(nm)`` makes the few polars calls of the scoring path work on numpy
columns, so that ``motif_model_bin`` / ``get_parent_scores`` / ``find_best_candidates`` run unmodified too.

Nothing in ``nanomotif_b200/`` may import this module.

Recipe follows SURVEY.md Appendix A.  Placeholders raise ``AttributeError`` for dunder names
so that pytest/hypothesis module introspection keeps working.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NANOMOTIF_REFERENCE", "/root/reference")

_STUBBED = [
    "polars", "polars.testing", "epymetheus", "epymetheus.epymetheus", "pysam", "pyfastx",
    "progressbar", "pyinstrument", "snakemake", "hdbscan", "Bio", "Bio.SeqIO", "Bio.Seq",
    "Bio.SeqRecord",
]


class _Placeholder:
    """Permissive stand-in: callable, attribute-chaining, usable as a dict value."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _StubDataFrame:  # real class: nanomotif/motif.py:654 subclasses polars.DataFrame
    def __init__(self, *a, **k):
        pass


def _make_stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave like a package so that "import a.b" resolves through sys.modules

    def _getattr(attr, _name=name):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        if _name == "polars" and attr == "DataFrame":
            return _StubDataFrame
        return _Placeholder

    mod.__getattr__ = _getattr  # type: ignore[attr-defined]
    return mod


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nanomotif"))


def load_reference():
    """Return the imported reference package (``nanomotif``) or raise ImportError."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "nanomotif" in sys.modules and getattr(sys.modules["nanomotif"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["nanomotif"]
    for name in _STUBBED:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _make_stub(name)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # last: must never shadow this repo's own packages
    import nanomotif  # noqa: E402  (import-time side effects: np.random.seed(1), random.seed(2403))

    return nanomotif
