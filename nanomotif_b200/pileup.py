"""Columnar pileup table (host side) and adapters from the frames the reference passes around.

The reference keeps pileups as polars frames with columns contig, position, mod_type, strand,
fraction_mod, Nvalid_cov (nanomotif/dataload.py:86-99).  Here the same columns are plain numpy
arrays; polars / pandas frames and dicts of arrays are accepted wherever a pileup is expected.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

COLUMNS = ("contig", "position", "mod_type", "strand", "fraction_mod", "Nvalid_cov")


def _column(frame, name):
    """Column `name` of a polars / pandas frame, PileupTable or mapping as a numpy array, or None."""
    if isinstance(frame, PileupTable):
        return getattr(frame, name)
    if isinstance(frame, dict):
        v = frame.get(name)
        return None if v is None else np.asarray(v)
    cols = getattr(frame, "columns", None)
    if cols is not None and name not in list(cols):
        return None
    if hasattr(frame, "get_column"):  # polars
        return frame.get_column(name).to_numpy()
    col = frame[name]  # pandas
    return col.to_numpy() if hasattr(col, "to_numpy") else np.asarray(col)


def strand_codes(strand) -> np.ndarray:
    """'+' -> 0, '-' -> 1, anything else (modkit's '.' with --combine-strands) -> 2: such rows take no part, as
    the reference's strand == "+" / strand == "-" filters drop them (find_motifs_bin.py:1311-1314).  Integer input
    is passed through."""
    a = np.asarray(strand)
    if a.dtype.kind in "iub":
        return a.astype(np.uint8, copy=False)
    return np.where(a == "+", 0, np.where(a == "-", 1, 2)).astype(np.uint8)


@dataclass
class PileupTable:
    """Pileup rows as columns.  `contig` holds contig names (object / str array); everything else is
    numeric except `strand` ('+'/'-' strings or 0/1 codes) and `mod_type` (modkit codes as strings)."""

    contig: np.ndarray | None
    position: np.ndarray
    strand: np.ndarray
    fraction_mod: np.ndarray
    mod_type: np.ndarray | None = None
    Nvalid_cov: np.ndarray | None = None
    extra: dict = field(default_factory=dict)

    def __len__(self) -> int:
        return int(len(self.position))

    @property
    def height(self) -> int:
        return len(self)

    def take(self, mask_or_index) -> "PileupTable":
        f = lambda a: None if a is None else np.asarray(a)[mask_or_index]
        return PileupTable(f(self.contig), f(self.position), f(self.strand), f(self.fraction_mod),
                           f(self.mod_type), f(self.Nvalid_cov), {k: f(v) for k, v in self.extra.items()})

    @classmethod
    def from_frame(cls, frame) -> "PileupTable":
        if isinstance(frame, PileupTable):
            return frame
        pos = _column(frame, "position")
        if pos is None:
            raise KeyError("pileup has no 'position' column")
        strand = _column(frame, "strand")
        frac = _column(frame, "fraction_mod")
        if strand is None or frac is None:
            raise KeyError("pileup needs 'strand' and 'fraction_mod' columns")
        return cls(_column(frame, "contig"), np.asarray(pos, dtype=np.int64), np.asarray(strand),
                   np.asarray(frac, dtype=np.float64), _column(frame, "mod_type"), _column(frame, "Nvalid_cov"))
