"""Exhaustive candidate sweep (K8, BASELINE.json configs[4]): the (n_mod, n_nomod) counts that motif_model_bin
(nanomotif/find_motifs_bin.py:1265-1331) would return for EVERY IUPAC motif of length 4..8 and every modified
position, from one histogram pass over the assembly plus a subset-sum transform -- not one scan per motif.

    index = SweepIndex(assembly, pileup, modtype_index)            # one pass: window x offset x class histograms
    index.add(contig_begin, contig_end)                            # a bin, or everything
    n_mod, n_nomod = index.table(k, mod_pos, canonical="A")        # device uint32 [15^(k-1)]: all motifs of length k
    index.counts("GRNGAAGY", 5)                                    # one motif's counts (table lookup)
    index.candidates(k, mod_pos, "A", min_mean=0.7, min_mod=50)    # motifs above a posterior-mean / support bar

Letters of a motif are indexed in the order of nanomotif/constants.py:2 (A T G C R Y S W K M B D H V N); the
modified position's own letter is fixed to the concrete canonical base and left out of the table index (a pileup
row of mod type 'a' only exists under an A, so every other letter there gives zero or a duplicate).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr
from .device import DeviceAssembly, DevicePileup, _stream

IUPAC_ORDER = "ATGCRYSWKMBDHVN"
_BASE_DIGIT = {"A": 0, "T": 1, "G": 2, "C": 3}
MIN_K, MAX_K = 4, 8


def hist_offset(k: int) -> int:
    return sum(j * 2 * 5 ** j for j in range(MIN_K, k))


class SweepIndex:
    """Window histograms of one mod type (nmb_sweep_hist), accumulated over the contig ranges added so far."""

    def __init__(self, assembly: DeviceAssembly, pileup: DevicePileup, modtype_index: int = 0):
        self.asm, self.pileup, self.modtype = assembly, pileup, int(modtype_index)
        if not 0 <= self.modtype < pileup.n_modtypes:
            raise ValueError("mod type index outside the pileup's class records")
        with torch.cuda.device(assembly.device):
            # RAW histogram (nmb_sweep_hist): k = 8 counted row by row, k < 8 only the contig-end windows; raw
            # histograms add (more contig ranges, other GPUs).  `hist` is the finalized copy, made on demand.
            self.raw = torch.zeros(int(lib.nmb_sweep_hist_size()), dtype=torch.int32, device=assembly.device)
        self._hist = None
        self.bip_raw = None
        self._bip = None
        self._tables: dict = {}

    @property
    def hist(self) -> torch.Tensor:
        """The finished histogram of every k = 4..8 (nmb_sweep_finalize applied to a copy of the raw one)."""
        if self._hist is None:
            with torch.cuda.device(self.asm.device):
                h = self.raw.clone()
                check(lib.nmb_sweep_finalize(ptr(h), _stream()), "nmb_sweep_finalize")
            self._hist = h
        return self._hist

    def add(self, contig_begin: int = 0, contig_end: int | None = None) -> "SweepIndex":
        asm = self.asm
        contig_end = asm.n_contigs if contig_end is None else contig_end
        tile_begin, tile_count = asm.tile_span(contig_begin, contig_end)
        cls = self.pileup.class_records
        view = asm.view()
        with torch.cuda.device(asm.device):
            base = ptr(cls) + self.modtype * asm.n_tiles * _lib.CLS_REC_WORDS * 4
            check(lib.nmb_sweep_hist(C.byref(view), base, tile_begin, tile_count, contig_begin, contig_end, ptr(self.raw),
                                     _stream()), "nmb_sweep_hist")
        self._tables.clear()
        self._hist = None
        return self

    # ---- bipartite shapes X{3,4} N{4..8} Y{3,4} over ACGT ----
    def add_bipartite(self, contig_begin: int = 0, contig_end: int | None = None) -> "SweepIndex":
        asm = self.asm
        contig_end = asm.n_contigs if contig_end is None else contig_end
        tile_begin, tile_count = asm.tile_span(contig_begin, contig_end)
        view = asm.view()
        with torch.cuda.device(asm.device):
            if self.bip_raw is None:
                self.bip_raw = torch.zeros(int(lib.nmb_sweep_bipartite_size()), dtype=torch.int32, device=asm.device)
            base = ptr(self.pileup.class_records) + self.modtype * asm.n_tiles * _lib.CLS_REC_WORDS * 4
            check(lib.nmb_sweep_bipartite(C.byref(view), base, tile_begin, tile_count, contig_begin, contig_end,
                                          ptr(self.bip_raw), _stream()), "nmb_sweep_bipartite")
        self._bip = None
        return self

    @property
    def bip(self):
        """The finished bipartite histogram (nmb_sweep_bipartite_finalize applied to a copy of the raw one), or None."""
        if self.bip_raw is None:
            return None
        if self._bip is None:
            with torch.cuda.device(self.asm.device):
                h = self.bip_raw.clone()
                check(lib.nmb_sweep_bipartite_finalize(ptr(h), _stream()), "nmb_sweep_bipartite_finalize")
            self._bip = h
        return self._bip

    @staticmethod
    def bipartite_block(a: int, g: int, b: int) -> tuple[int, int]:
        """(first counter, counters per (offset, class)) of shape X{a} N{g} Y{b} in the bipartite histogram."""
        if a not in (3, 4) or b not in (3, 4) or not 4 <= g <= 8:
            raise ValueError("bipartite shapes are X{3,4} N{4..8} Y{3,4}")
        block = lambda x, y: (x + y) * 2 * 4 ** (x + y)
        per_gap = block(3, 3) + block(3, 4) + block(4, 3) + block(4, 4)
        within = {(3, 3): 0, (3, 4): block(3, 3), (4, 3): block(3, 3) + block(3, 4),
                  (4, 4): block(3, 3) + block(3, 4) + block(4, 3)}[(a, b)]
        return (g - 4) * per_gap + within, 4 ** (a + b)

    def bipartite_table(self, a: int, g: int, b: int, mod_pos: int):
        """(n_mod, n_nomod) device views over all 4^(a+b) motifs of shape X{a} N{g} Y{b} with the modified base at
        motif position mod_pos (inside X or Y).  Entry index: bipartite_index(left, right)."""
        if a <= mod_pos < a + g or not 0 <= mod_pos < a + g + b:
            raise ValueError("the modified position must lie in one of the two concrete parts")
        o = mod_pos if mod_pos < a else mod_pos - g
        first, n = self.bipartite_block(a, g, b)
        base = first + o * 2 * n
        return self.bip[base:base + n], self.bip[base + n:base + 2 * n]

    @staticmethod
    def bipartite_index(left: str, right: str) -> int:
        a, b = len(left), len(right)
        x = lambda part: sum((_BASE_DIGIT[c] >> 1) << i for i, c in enumerate(part))
        y = lambda part: sum((_BASE_DIGIT[c] & 1) << i for i, c in enumerate(part))
        return x(left) | (y(left) << a) | (x(right) << (2 * a)) | (y(right) << (2 * a + b))

    def bipartite_counts(self, left: str, gap: int, right: str, mod_pos: int) -> tuple[int, int]:
        """(n_mod, n_nomod) of the motif left + N*gap + right with the modified base at mod_pos."""
        n_mod, n_nomod = self.bipartite_table(len(left), gap, len(right), mod_pos)
        i = self.bipartite_index(left, right)
        return int(n_mod[i].item()) & 0xFFFFFFFF, int(n_nomod[i].item()) & 0xFFFFFFFF

    def all_reduce(self) -> "SweepIndex":
        """Sum the histograms over the ranks of the default process group (contig-sharded assemblies)."""
        import torch.distributed as dist

        dist.all_reduce(self.raw)  # raw histograms add; the marginalisation runs once, on the sum
        if self.bip_raw is not None:
            dist.all_reduce(self.bip_raw)
        self._tables.clear()
        self._hist = self._bip = None
        return self

    def table(self, k: int, mod_pos: int, canonical: str = "A", keep: bool = False):
        """(n_mod, n_nomod): device uint32-as-int32 tensors of 15^(k-1) entries, index = the motif's letters except the
        modified position as base-15 digits (first letter most significant, IUPAC_ORDER)."""
        if not (MIN_K <= k <= MAX_K and 0 <= mod_pos < k):
            raise ValueError("k must be 4..8 and 0 <= mod_pos < k")
        key = (k, mod_pos, canonical)
        if key in self._tables:
            return self._tables[key]
        d = self.asm.device
        out = []
        with torch.cuda.device(d):
            n5 = 5 ** (k - 1)
            for cls in (0, 1):
                src = self.hist[hist_offset(k) + (mod_pos * 2 + cls) * 5 ** k:][:5 ** k]
                a = torch.empty(n5, dtype=torch.int32, device=d)
                check(lib.nmb_sweep_slice(ptr(src), ptr(a), n5, 5 ** (k - 1 - mod_pos), _BASE_DIGIT[canonical], _stream()),
                      "nmb_sweep_slice")
                for step in range(k - 1):  # last axis first: the big final passes run with a long contiguous inner
                    outer, inner = 5 ** (k - 2 - step), 15 ** step
                    b = torch.empty(outer * 15 * inner, dtype=torch.int32, device=d)
                    check(lib.nmb_sweep_expand(ptr(a), ptr(b), outer, inner, _stream()), "nmb_sweep_expand")
                    a = b
                out.append(a)
        if keep:
            self._tables[key] = tuple(out)
        return tuple(out)

    @staticmethod
    def motif_index(motif_iupac: str, mod_pos: int) -> int:
        idx = 0
        for i, ch in enumerate(motif_iupac):
            if i != mod_pos:
                idx = idx * 15 + IUPAC_ORDER.index(ch)
        return idx

    @staticmethod
    def index_motif(index: int, k: int, mod_pos: int, canonical: str = "A") -> str:
        letters = []
        for i in reversed(range(k)):
            if i == mod_pos:
                letters.append(canonical)
            else:
                letters.append(IUPAC_ORDER[index % 15])
                index //= 15
        return "".join(reversed(letters))

    def counts(self, motif_iupac: str, mod_pos: int, keep: bool = True) -> tuple[int, int]:
        """(n_mod, n_nomod) of one IUPAC motif of length 4..8 -- what motif_model_bin adds to the prior."""
        base = motif_iupac[mod_pos]
        if base not in _BASE_DIGIT:
            raise ValueError("the modified position must hold a concrete base")
        n_mod, n_nomod = self.table(len(motif_iupac), mod_pos, base, keep=keep)
        i = self.motif_index(motif_iupac, mod_pos)
        return int(n_mod[i].item()) & 0xFFFFFFFF, int(n_nomod[i].item()) & 0xFFFFFFFF

    def candidates(self, k: int, mod_pos: int, canonical: str = "A", min_mean: float = 0.7, min_mod: int = 50,
                   canonical_form: bool = True) -> list[tuple[str, int, int]]:
        """[(motif, n_mod, n_nomod)] of length k with posterior mean >= min_mean and n_mod >= min_mod; with
        `canonical_form` motifs that start or end with N are left out (they are the shorter motif)."""
        n_mod, n_nomod = self.table(k, mod_pos, canonical)
        d = self.asm.device
        n = int(n_mod.numel())
        with torch.cuda.device(d):
            n_out = torch.zeros(1, dtype=torch.int64, device=d)
            check(lib.nmb_sweep_filter(ptr(n_mod), ptr(n_nomod), n, float(min_mean), int(min_mod), None, 0, ptr(n_out),
                                       _stream()), "nmb_sweep_filter")
            m = int(n_out.item())
            idx = torch.empty(max(m, 1), dtype=torch.int64, device=d)
            if m:
                check(lib.nmb_sweep_filter(ptr(n_mod), ptr(n_nomod), n, float(min_mean), int(min_mod), ptr(idx), m,
                                           ptr(n_out), _stream()), "nmb_sweep_filter")
            idx = idx[:m]
            a = n_mod[idx].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            b = n_nomod[idx].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            idx = idx.cpu().numpy()
        out = []
        for i, x, y in sorted(zip(idx.tolist(), a.tolist(), b.tolist())):
            s = self.index_motif(i, k, mod_pos, canonical)
            if canonical_form and (s[0] == "N" or s[-1] == "N"):
                continue
            out.append((s, x, y))
        return out
