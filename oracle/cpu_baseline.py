"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU path for the scoring operator, timed.

Runs, per (motif, contig), the very calls the reference makes: ``regex.finditer(overlapped=True)`` +
``np.isin(assume_unique=True)`` on the forward motif and on its reverse complement
(nanomotif/utils.py:44-67, nanomotif/find_motifs_bin.py:1234-1331), summed over the contigs of a bin
like ``motif_model_bin`` (find_motifs_bin.py:1265-1283), over a work list in a ``multiprocessing`` spawn
pool with chunksize 1 -- the reference's own parallelisation (find_motifs_bin.py:323-351).

Two implementations of the per-contig step:
  kind "reference"  the UNMODIFIED functions ``nanomotif.find_motifs_bin.methylated_motif_occourances`` /
                    ``nanomotif.utils.subseq_indices`` / ``Motif.reverse_compliment`` imported from
                    /root/reference through oracle/ref_shim.py -- only where that tree is mounted (this
                    container; the GPU box does not have it)
  kind "port"       oracle/restate.py, the line-by-line restatement pinned against the former

The pileup is pre-split ONCE per worker into the four position arrays per (contig, mod type) instead of
the four polars filters (and the per-contig ``pileup.filter``) the reference runs on every call: this
favours the CPU side.  Only bench.py's cpu_baseline / --impl reference legs and tests
may import this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import restate as O

_G = {}


def _init(bins: dict, kind: str):
    os.environ["OMP_NUM_THREADS"] = "1"
    _G["bins"] = bins  # bin key -> [(contig string, {mod type index: (meth_fwd, non_fwd, meth_rev, non_rev)})]
    _G["kind"] = kind
    if kind == "reference":
        from . import ref_shim

        nm = ref_shim.load_reference()
        _G["Motif"] = nm.motif.Motif
        _G["occ"] = nm.find_motifs_bin.methylated_motif_occourances


def _count_contig_port(seq, split, motif, mod_pos):
    meth_fwd, non_fwd, meth_rev, non_rev = split
    s_motif, s_pos = O.strip_motif(motif, mod_pos)
    a, b = O.methylated_motif_occourances(s_motif, s_pos, seq, meth_fwd, non_fwd)
    rc, rp = O.reverse_complement_motif(s_motif, s_pos)
    c, d = O.methylated_motif_occourances(rc, rp, seq, meth_rev, non_rev)
    return len(a) + len(c), len(b) + len(d)


def _count_contig_reference(seq, split, motif, mod_pos):
    """find_motifs_bin.py:1307-1320 with the real Motif type and the real join."""
    meth_fwd, non_fwd, meth_rev, non_rev = split
    m = _G["Motif"](motif, mod_pos).new_stripped_motif()
    a, b = _G["occ"](m, seq, meth_fwd, non_fwd)
    c, d = _G["occ"](m.reverse_compliment(), seq, meth_rev, non_rev)
    return len(a) + len(c), len(b) + len(d)


def _task(item):
    key, motif, mod_pos, mt = item
    f = _count_contig_reference if _G["kind"] == "reference" else _count_contig_port
    n_mod = n_nomod = 0
    for seq, splits in _G["bins"][key]:  # motif_model_bin: the same model threaded through every contig
        a, b = f(seq, splits[mt], motif, mod_pos)
        n_mod += a
        n_nomod += b
    return n_mod, n_nomod


def presplit(position, strand, mod_type, fraction_mod, n_modtypes: int, low=0.3, high=0.7) -> dict:
    """find_motifs_bin.py:1308-1314 applied once per mod type."""
    position = np.asarray(position, dtype=np.int64)
    out = {}
    for mt in range(n_modtypes):
        sel = mod_type == mt
        p, s, f = position[sel], strand[sel], fraction_mod[sel]
        hi, lo = f >= high, f <= low
        out[mt] = (p[hi & (s == 0)], p[lo & (s == 0)], p[hi & (s == 1)], p[lo & (s == 1)])
    return out


def presplit_bin(ascii_u8, lengths, contig, position, strand, mod_type, fraction_mod, n_modtypes: int,
                 low=0.3, high=0.7) -> list:
    """One bin (contigs concatenated in `ascii_u8`, rows carrying the contig index) -> [(contig string,
    presplit of its rows)]; rows must be grouped by contig (a modkit pileup is)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    bounds = np.searchsorted(contig, np.arange(len(lengths) + 1))
    out = []
    for c, (s, n) in enumerate(zip(starts.tolist(), lengths.tolist())):
        lo_, hi_ = bounds[c], bounds[c + 1]
        out.append((ascii_u8[s:s + n].tobytes().decode("ascii"),
                    presplit(position[lo_:hi_], strand[lo_:hi_], mod_type[lo_:hi_], fraction_mod[lo_:hi_], n_modtypes,
                             low, high)))
    return out


def best_kind() -> str:
    from . import ref_shim

    return "reference" if ref_shim.reference_available() else "port"


class CpuPool:
    """Spawn pool holding the contig strings and the pre-split pileup of every bin in every worker.

    `CpuPool(seq, split)` (one contig, the round-1 form) or `CpuPool(bins={key: presplit_bin(...)})`;
    work items are (motif, mod_pos, mod type) for the former and (key, motif, mod_pos, mod type) for the latter."""

    def __init__(self, seq: str | None = None, split: dict | None = None, workers: int | None = None,
                 bins: dict | None = None, kind: str = "port"):
        self.workers = workers or os.cpu_count() or 1
        self.single = bins is None
        if bins is None:
            bins = {0: [(seq, split)]}
        self.kind = kind
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=(bins, kind))

    def run(self, work: list) -> tuple[list, float]:
        """Score the work list; returns (counts, seconds)."""
        if self.single:
            work = [(0, *w) for w in work]
        t0 = time.perf_counter()
        res = list(self.pool.imap(_task, work, chunksize=1))
        return res, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
