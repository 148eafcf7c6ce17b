"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU path for the scoring operator, timed.

Runs ``oracle.restate.motif_model_contig`` -- i.e. the very calls the reference makes per (motif,
contig): ``regex.finditer(overlapped=True)`` + ``np.isin(assume_unique=True)`` on the forward motif
and on its reverse complement (nanomotif/utils.py:44-67, nanomotif/find_motifs_bin.py:1234-1331) --
over a work list of (motif, mod type) pairs in a ``multiprocessing`` spawn pool with chunksize 1,
the reference's own parallelisation (nanomotif/find_motifs_bin.py:323-351).

The pileup is pre-split ONCE per worker into the four position arrays per mod type instead of the
four polars filters the reference runs on every call: this favours the CPU side.
Only bench.py's cpu_baseline / --impl reference legs and tests may import this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import restate as O

_G = {}


def _init(seq: str, split: dict):
    os.environ["OMP_NUM_THREADS"] = "1"
    _G["seq"] = seq
    _G["split"] = split  # mod type index -> (meth_fwd, non_fwd, meth_rev, non_rev) int64 arrays


def _task(item):
    motif, mod_pos, mt = item
    seq = _G["seq"]
    meth_fwd, non_fwd, meth_rev, non_rev = _G["split"][mt]
    s_motif, s_pos = O.strip_motif(motif, mod_pos)
    a, b = O.methylated_motif_occourances(s_motif, s_pos, seq, meth_fwd, non_fwd)
    rc, rp = O.reverse_complement_motif(s_motif, s_pos)
    c, d = O.methylated_motif_occourances(rc, rp, seq, meth_rev, non_rev)
    return len(a) + len(c), len(b) + len(d)


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


class CpuPool:
    """Spawn pool holding the contig string and the pre-split pileup in every worker."""

    def __init__(self, seq: str, split: dict, workers: int | None = None):
        self.workers = workers or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=(seq, split))
        self.seq_len = len(seq)

    def run(self, work: list) -> tuple[list, float]:
        """Score the (motif, mod_pos, mod type) work list; returns (counts, seconds)."""
        t0 = time.perf_counter()
        res = list(self.pool.imap(_task, work, chunksize=1))
        return res, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
