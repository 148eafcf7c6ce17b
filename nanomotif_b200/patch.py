"""Install the B200 backend underneath an imported reference package (monkey-patching).

The reference resolves its scoring operators through module globals at call time
(SURVEY.md 8b), so replacing the attributes swaps the backend without editing nanomotif:

    import nanomotif, nanomotif_b200.patch
    nanomotif_b200.patch.install(nanomotif)        # this process

`subseq_indices` is imported BY NAME into nanomotif.find_motifs_bin (find_motifs_bin.py:18), so both
the defining module and that importer are patched.  The pileup frames the reference passes (polars) are
accepted as they are (nanomotif_b200.pileup.PileupTable.from_frame / dataload.rows_from_table).

Worker processes: the reference starts its pool with get_context("spawn") (find_motifs_bin.py:323), and a
spawned worker re-imports nanomotif from scratch -- a patch applied in the parent is gone.  Two spawn-safe
ways to get the backend into every worker:

  * `enable_for_workers()` in the parent before the pool is created: exports NMB200_PATCH=1 and puts
    nanomotif_b200/hooks (a sitecustomize.py) at the front of PYTHONPATH, both of which spawned children
    inherit; the child's sitecustomize calls `install_import_hook()`, which patches nanomotif.utils and
    nanomotif.find_motifs_bin the moment they are imported.
  * or run with the environment prepared by hand: NMB200_PATCH=1 PYTHONPATH=<repo>/nanomotif_b200/hooks:<repo>.
"""
from __future__ import annotations

import importlib.abc
import os
import sys

from . import api

_PATCHED = {
    "utils": {"subseq_indices": api.subseq_indices},
    "find_motifs_bin": {
        "subseq_indices": api.subseq_indices,
        "methylated_motif_occourances": api.methylated_motif_occourances,
        "motif_model_contig": api.motif_model_contig,
        "motif_model_bin": api.motif_model_bin,
        "get_parent_scores": api.get_parent_scores,
    },
}
HOOK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hooks")
ENV_FLAG = "NMB200_PATCH"


def patch_module(mod, short_name: str) -> dict:
    """Patch one reference module (`utils` or `find_motifs_bin`) in place; returns {name: original}."""
    saved = {}
    for name, fn in _PATCHED.get(short_name, {}).items():
        saved[name] = getattr(mod, name)
        setattr(mod, name, fn)
    return saved


def install(nanomotif_pkg) -> dict:
    """Patch `nanomotif_pkg` in place; returns {(module, name): original} for `uninstall`."""
    saved = {}
    for modname in _PATCHED:
        for name, fn in patch_module(getattr(nanomotif_pkg, modname), modname).items():
            saved[(modname, name)] = fn
    return saved


def uninstall(nanomotif_pkg, saved: dict) -> None:
    for (modname, name), fn in saved.items():
        setattr(getattr(nanomotif_pkg, modname), name, fn)


class _PatchOnImport(importlib.abc.MetaPathFinder):
    """Meta-path finder that lets the normal finders locate nanomotif.utils / nanomotif.find_motifs_bin and patches
    the module right after its body has run."""

    TARGETS = {"nanomotif.utils": "utils", "nanomotif.find_motifs_bin": "find_motifs_bin"}

    def find_spec(self, fullname, path=None, target=None):
        short = self.TARGETS.get(fullname)
        if short is None:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None and hasattr(spec.loader, "exec_module"):
                break
        else:
            return None
        original = spec.loader.exec_module

        def exec_module(module, _original=original, _short=short):
            _original(module)
            patch_module(module, _short)

        spec.loader.exec_module = exec_module
        return spec


def install_import_hook() -> None:
    """Patch the reference modules as soon as they are imported in THIS process (idempotent).  Modules that are
    already imported are patched right away."""
    if not any(isinstance(f, _PatchOnImport) for f in sys.meta_path):
        sys.meta_path.insert(0, _PatchOnImport())
    for fullname, short in _PatchOnImport.TARGETS.items():
        if fullname in sys.modules:
            patch_module(sys.modules[fullname], short)


def enable_for_workers() -> None:
    """Make every process spawned from now on install the import hook at start-up (see module docstring)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    parts = [HOOK_DIR, root] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p and p not in (HOOK_DIR, root)]
    os.environ["PYTHONPATH"] = os.pathsep.join(parts)
    os.environ[ENV_FLAG] = "1"
