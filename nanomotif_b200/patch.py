"""Install the B200 backend underneath an imported reference package (monkey-patching).

The reference resolves its scoring operators through module globals at call time
(SURVEY.md 8b), so replacing the attributes swaps the backend without editing nanomotif:

    import nanomotif, nanomotif_b200.patch
    nanomotif_b200.patch.install(nanomotif)        # in every worker process (spawned workers re-import)

`subseq_indices` is imported BY NAME into nanomotif.find_motifs_bin (find_motifs_bin.py:18), so both
the defining module and that importer are patched.  The pileup frames the reference passes (polars) are
accepted as they are (nanomotif_b200.pileup.PileupTable.from_frame).
"""
from __future__ import annotations

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


def install(nanomotif_pkg) -> dict:
    """Patch `nanomotif_pkg` in place; returns {(module, name): original} for `uninstall`."""
    saved = {}
    for modname, names in _PATCHED.items():
        mod = getattr(nanomotif_pkg, modname)
        for name, fn in names.items():
            saved[(modname, name)] = getattr(mod, name)
            setattr(mod, name, fn)
    return saved


def uninstall(nanomotif_pkg, saved: dict) -> None:
    for (modname, name), fn in saved.items():
        setattr(getattr(nanomotif_pkg, modname), name, fn)
