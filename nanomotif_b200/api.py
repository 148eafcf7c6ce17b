"""Drop-in replacements for the reference's scoring operators, backed by libnmb200 (CUDA, sm_100a).

Same names, argument meaning, return types and error behaviour as the reference callables
(SURVEY.md section 8a/8b):

    subseq_indices                 nanomotif/utils.py:44-67
    methylated_motif_occourances   nanomotif/find_motifs_bin.py:1234-1263
    motif_model_contig             nanomotif/find_motifs_bin.py:1285-1331
    motif_model_bin                nanomotif/find_motifs_bin.py:1265-1283
    get_parent_scores              nanomotif/find_motifs_bin.py:1382-1433

plus the batched forms the GPU needs to be kept busy (`motif_model_bin_many`, `BinScorer`).
There is no CPU fallback: without a CUDA device every function raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr
from .device import (DeviceAssembly, DevicePileup, MotifPrograms, _stream, compact_rows, make_jobs, scan_count,
                     sequence_of)
from .model import BetaBernoulliModel, predictive_evaluation_score
from .motif import Motif, as_motif, tokenize
from .pileup import PileupTable, strand_codes

# ---------------------------------------------------------------------------------------------
# small identity-keyed caches so that repeated calls with the same Python objects (what the
# reference's search loop does) do not re-pack / re-upload
# ---------------------------------------------------------------------------------------------


class _IdCache:
    """LRU of device state keyed by the identity of the caller's objects PLUS a cheap content fingerprint (the
    reference functions are pure: a caller may mutate a dict of contigs or filter a pileup in place between calls),
    bounded by entry count and by device bytes (NMB_CACHE_BYTES, default 16 GiB)."""

    def __init__(self, capacity: int):
        import os

        self.capacity = capacity
        self.byte_budget = int(os.environ.get("NMB_CACHE_BYTES", 16 << 30))
        self._d: OrderedDict = OrderedDict()

    def get(self, key, refs, build, nbytes=lambda v: 0):
        hit = self._d.get(key)
        if hit is not None and all(a is b for a, b in zip(hit[0], refs)):
            self._d.move_to_end(key)
            return hit[1]
        val = build()
        self._d[key] = (refs, val, int(nbytes(val)))  # holding refs keeps the ids stable while cached
        while len(self._d) > 1 and (len(self._d) > self.capacity or sum(e[2] for e in self._d.values()) > self.byte_budget):
            self._d.popitem(last=False)
        return val

    def clear(self):
        self._d.clear()


def _sample(a, k: int = 64) -> tuple:
    """k evenly spaced elements of a column (O(k)): part of the cache fingerprint."""
    n = len(a)
    if n == 0:
        return ()
    idx = np.linspace(0, n - 1, num=min(k, n)).astype(np.int64)
    try:
        return tuple(np.asarray(a)[idx].tolist())
    except Exception:
        return tuple(a[int(i)] for i in idx)


def _pileup_fingerprint(pileup) -> tuple:
    """(rows, samples of position and fraction_mod): catches in-place filtering / edits of a cached pileup."""
    from .pileup import _column

    try:
        pos, frac = _column(pileup, "position"), _column(pileup, "fraction_mod")
    except Exception:
        return (id(pileup),)
    if pos is None or frac is None:
        return (id(pileup),)
    return (len(pos), _sample(pos), _sample(frac))


def _contigs_fingerprint(contigs) -> tuple:
    if isinstance(contigs, str):
        return (len(contigs),)
    return (len(contigs), sum(len(sequence_of(c)) for c in contigs.values()), tuple(contigs.keys())[:4])


def _scorer_bytes(scorer) -> int:
    t = [scorer.assembly.seq_records, scorer.assembly.nonacgt, scorer.pileup.class_records]
    return sum(x.numel() * x.element_size() for x in t)


_seq_cache = _IdCache(8)
_bin_cache = _IdCache(8)


def clear_caches() -> None:
    """Drop the device state cached for the drop-in functions (packed assemblies, class planes)."""
    _seq_cache.clear()
    _bin_cache.clear()


def _single_contig_assembly(seq: str) -> DeviceAssembly:
    return _seq_cache.get(("seq", id(seq), len(seq)), (seq,), lambda: DeviceAssembly.from_sequences({"_": seq}),
                          lambda a: a.seq_records.numel() * 4 + a.nonacgt.numel() * 4)


# ---------------------------------------------------------------------------------------------
# K3-backed position functions
# ---------------------------------------------------------------------------------------------


def _split_literals(toks: list[str]) -> tuple[list[str], dict[int, str]]:
    """Motif tokens with literal non-ACGT letters (e.g. an 'N' typed into a motif string) replaced by '.', and
    {token index: letter} of those literals.  Under the reference's regex semantics (utils.py:61-66) such a letter
    matches the same contig letter and nothing else; it is applied as a letter plane on top of the matcher."""
    lit = {j: t for j, t in enumerate(toks) if len(t) == 1 and t not in "ACGT."}
    for t in lit.values():
        if not ("A" <= t <= "Z"):
            raise ValueError(f"motif token {t!r}: only letters, '.' and [..] classes are supported")
    return ["." if j in lit else t for j, t in enumerate(toks)], lit


def _apply_literals(plane: torch.Tensor, seq: str, literals: dict[int, str], t0: int) -> None:
    """plane (bit p <-> motif token t0 at contig position p) &= 'token j is the literal letter' for every literal."""
    if not literals:
        return
    from .device import _to_device

    d = plane.device
    with torch.cuda.device(d):
        ascii_d = _to_device(np.frombuffer(seq.encode("ascii"), dtype=np.uint8), d)
        for j, letter in literals.items():
            check(lib.nmb_letter_plane(ptr(ascii_d), len(seq), ord(letter), j - t0, ptr(plane), int(plane.numel()),
                                       _stream()), "nmb_letter_plane")


def _match_plane(asm: DeviceAssembly, motif: Motif, align: int, strand: int = 0, seq: str | None = None
                 ) -> tuple[torch.Tensor, int, int]:
    """Device bit-plane of the occurrences of `motif` (regex as given, flanking wildcards allowed).

    Returns (plane, delta, length): bit p of `plane` is set iff the motif occurs with motif
    position `align - delta` at p.  delta is 0 unless `align` points into the flanking wildcards.
    The caller restores the whole-motif-inside-contig rule with `length`.  Literal non-ACGT letters in the motif
    need `seq` (the contig text, single-contig assemblies only).
    """
    toks, literals = _split_literals(tokenize(motif.string))
    length = len(toks)
    if literals and seq is None:
        raise ValueError("a motif with literal non-ACGT letters needs the contig text")
    lead = 0
    while lead < length and toks[lead] == ".":
        lead += 1
    if lead == length:  # nothing but literals and wildcards: every position is a candidate
        plane = torch.full((asm.n_words,), -1, dtype=torch.int32, device=asm.device)
        _apply_literals(plane, seq, literals, align)
        return plane, 0, length
    core = Motif("".join(toks).strip("."), 0)
    core_len = len(tokenize(core.string))
    a = min(max(align - lead, 0), core_len - 1)
    delta = (align - lead) - a
    progs = MotifPrograms([core], asm.device, strip=False, mod_pos_override=a)
    plane = torch.empty(asm.n_words, dtype=torch.int32, device=asm.device)
    view = asm.view()
    check(lib.nmb_match_plane(C.byref(view), ptr(progs.programs), 0, strand, progs.max_len, 0, asm.n_tiles,
                              ptr(plane), _stream()), "nmb_match_plane")
    _apply_literals(plane, seq, literals, align - delta)
    return plane, delta, length


def subseq_indices(subseq: str, seq: str) -> np.ndarray:
    """All (overlapping) 0-based start positions of the regex motif `subseq` in `seq`, ascending int64.

    Supports A C G T, '.', bracket classes, and literal non-ACGT letters (an 'N' matches the contig letter N only --
    regex-literal semantics, utils.py:61-66).
    """
    seq = sequence_of(seq)
    toks = tokenize(subseq)
    length, L = len(toks), len(seq)
    if length == 0:
        raise ValueError("empty motif")
    if all(t == "." for t in toks):  # closed form: every start where the motif fits
        return np.arange(0, max(0, L - length + 1), dtype=np.int64)
    asm = _single_contig_assembly(seq)
    plain, _ = _split_literals(toks)
    lead = next((i for i, t in enumerate(plain) if t != "."), 0)
    # align at the first constrained position (position 0 for motifs made of literals and wildcards only)
    plane, _, length = _match_plane(asm, Motif(subseq, 0), align=lead, seq=seq)
    pos = _compact(plane, 0, L)
    # whole (unstripped) motif inside the contig: 0 <= start and start + length <= L
    pos = pos - lead
    pos = pos[(pos >= 0) & (pos <= L - length)]
    return pos.cpu().numpy()


def _compact(plane: torch.Tensor, pos_begin: int, pos_end: int, mask: torch.Tensor | None = None) -> torch.Tensor:
    d = plane.device
    n_blocks = -(-(pos_end - pos_begin) // _lib.TILE_BP)
    scratch = torch.empty(n_blocks + 2, dtype=torch.int64, device=d)
    n_out = torch.zeros(1, dtype=torch.int64, device=d)
    # pass 1 sizes the output, pass 2 writes it (both inside nmb_compact_positions when capacity > 0)
    check(lib.nmb_compact_positions(ptr(plane), ptr(mask), pos_begin, pos_end, ptr(scratch), None, 0, ptr(n_out),
                                    _stream()), "nmb_compact_positions")
    n = int(n_out.item())
    out = torch.empty(n, dtype=torch.int64, device=d)
    if n:
        check(lib.nmb_compact_positions(ptr(plane), ptr(mask), pos_begin, pos_end, ptr(scratch), ptr(out), n,
                                        ptr(n_out), _stream()), "nmb_compact_positions")
    return out


def _test(plane: torch.Tensor, base: int, limit: int, positions: np.ndarray) -> np.ndarray:
    """np.isin(positions, occurrences) evaluated on the device, in the order given."""
    pos = np.ascontiguousarray(positions, dtype=np.int64)
    if pos.size == 0:
        return np.zeros(0, dtype=bool)
    d = plane.device
    pos_d = torch.from_numpy(pos).to(d)
    flag = torch.empty(pos.size, dtype=torch.uint8, device=d)
    check(lib.nmb_test_positions(ptr(plane), base, limit, ptr(pos_d), pos.size, ptr(flag), _stream()),
          "nmb_test_positions")
    return flag.cpu().numpy().astype(bool)


def methylated_motif_occourances(motif, sequence, methylated_positions, non_methylated_positions) -> tuple:
    """Subsets of the two position arrays that coincide with the modified base of a motif occurrence,
    in the order given (find_motifs_bin.py:1234-1263)."""
    assert len(motif) > 0, "Motif is empty"
    assert len(sequence) > 0, "Sequence is empty"
    assert hasattr(motif, "mod_position") and isinstance(motif, str), "Motif is not a Motif type"
    m = as_motif(motif)
    sequence = sequence_of(sequence)
    meth = np.asarray(methylated_positions)
    nonmeth = np.asarray(non_methylated_positions)
    toks = tokenize(m.string)
    length, L, mp = len(toks), len(sequence), int(m.mod_position)
    if all(t == "." for t in toks):
        ok = lambda p: (p - mp >= 0) & (p - mp <= L - length)
        return meth[ok(meth)] if meth.size else meth, nonmeth[ok(nonmeth)] if nonmeth.size else nonmeth
    asm = _single_contig_assembly(sequence)
    plane, delta, length = _match_plane(asm, m, align=mp, seq=sequence)

    def pick(p):
        if p.size == 0:
            return p
        pi = p.astype(np.int64)
        keep = _test(plane, 0, L, pi - delta)
        keep &= (pi - mp >= 0) & (pi - mp <= L - length)  # whole motif inside the contig
        return p[keep]

    return pick(meth), pick(nonmeth)


# ---------------------------------------------------------------------------------------------
# K2-backed scoring
# ---------------------------------------------------------------------------------------------


class BinScorer:
    """Device context of one (bin, mod_type): the bin's contigs packed once, its pileup as class planes.

    `score(motifs)` evaluates any number of motifs in ONE launch and returns int64 counts
    [n_motifs, 2] = (n_mod, n_nomod) summed over both strands and all contigs -- exactly what
    motif_model_bin adds to the Beta(5,5) prior.
    """

    def __init__(self, pileup, contigs, low_meth_threshold: float, high_meth_threshold: float, device=None):
        table = PileupTable.from_frame(pileup)
        self.assembly = DeviceAssembly.from_sequences(contigs, device)
        asm = self.assembly
        if table.contig is None:
            if asm.n_contigs != 1:
                raise KeyError("pileup has no 'contig' column but several contigs were given")
            cid = np.zeros(len(table), dtype=np.int32)
        else:
            names = np.asarray(table.contig)
            uniq, inv = np.unique(names, return_inverse=True)
            lut = np.fromiter((asm.index.get(str(u), -1) for u in uniq), dtype=np.int32, count=len(uniq))
            cid = lut[inv] if len(uniq) else np.zeros(0, dtype=np.int32)
        self.table = table
        self.contig_id = cid
        # modkit percentages have two decimals: ship 7-byte rows over PCIe when that holds, else float64 rows
        compact = compact_rows(cid, table.position, strand_codes(table.strand), table.fraction_mod, None, asm.n_contigs)
        if compact is not None:
            self.pileup = DevicePileup.from_compact(asm, low=low_meth_threshold, high=high_meth_threshold, **compact)
        else:
            self.pileup = DevicePileup.from_columns(asm, cid, table.position, strand_codes(table.strand),
                                                    table.fraction_mod, low_meth_threshold, high_meth_threshold)

    def _jobs(self, n_motifs: int, per_contig: bool) -> tuple[np.ndarray, int]:
        asm = self.assembly
        jobs = make_jobs(1)
        jobs["motif_begin"], jobs["motif_count"] = 0, n_motifs
        jobs["tile_begin"], jobs["tile_count"] = 0, asm.n_tiles
        jobs["contig_begin"], jobs["contig_end"] = 0, asm.n_contigs
        jobs["group_mode"] = 1 if per_contig else 0
        groups = asm.n_contigs if per_contig else 1
        jobs["n_groups"] = groups
        return jobs, n_motifs * groups

    def counts_by_strand(self, motifs, per_contig: bool = False, motifs_per_item=None) -> torch.Tensor:
        """int64 device tensor [n_motifs, (n_contigs,) 4]: n_mod '+', n_nomod '+', n_mod '-', n_nomod '-'."""
        motifs = list(motifs)
        if not motifs:
            shape = (0, self.assembly.n_contigs, 4) if per_contig else (0, 4)
            return torch.zeros(shape, dtype=torch.int64, device=self.assembly.device)
        progs = MotifPrograms(motifs, self.assembly.device, strip=True)
        jobs, rows = self._jobs(len(motifs), per_contig)
        out = scan_count(self.assembly, self.pileup, progs, jobs, rows, motifs_per_item)
        return out.view(len(motifs), self.assembly.n_contigs, 4) if per_contig else out

    def score(self, motifs) -> np.ndarray:
        c = self.counts_by_strand(motifs).cpu().numpy()
        return np.stack([c[:, 0] + c[:, 2], c[:, 1] + c[:, 3]], axis=1)


class BinContext:
    """One (bin, mod_type) of a MultiBinScorer: a contig range, its tile span and a class-plane set."""

    def __init__(self, owner: "MultiBinScorer", bin_name, mod_type, contig_begin, contig_end, modtype_index):
        self.owner, self.bin_name, self.mod_type = owner, bin_name, mod_type
        self.contig_begin, self.contig_end, self.modtype_index = contig_begin, contig_end, modtype_index
        self.tile_begin, self.tile_count = owner.assembly.tile_span(contig_begin, contig_end)

    def score(self, motifs) -> np.ndarray:
        return self.owner.score_batch([(self, motifs)])[0]


class MultiBinScorer:
    """All bins and mod types of an assembly resident on one GPU; any number of (bin, mod_type, motifs)
    requests are scored by ONE scan launch (one job per request).  This is the data-parallel replacement of
    the reference's process pool over bins (nanomotif/find_motifs_bin.py:330-372)."""

    def __init__(self, pileup, bins: dict, mod_types, low_meth_threshold: float, high_meth_threshold: float,
                 device=None, keep_rows: bool = True, pieces=None):
        """bins: {bin name: {contig name: sequence}}.  pileup: what the reference hands to its workers
        (find_motifs_bin.py:399-427) -- ONE table with contig / position / strand / mod_type / fraction_mod columns
        (pyarrow Table, polars or pandas frame, PileupTable, dict of arrays), or the partitioned form
        {(bin, mod_type): table} / a list of tables -- or dataload.DeviceRows (rows already parsed on the device).
        String columns cross PCIe as their Arrow buffers and are resolved to ids on the device
        (dataload.rows_from_table): no per-row host work.

        pieces (sharding.ShardedMultiBinScorer, a contig cut into position ranges over several ranks): [(contig name
        of the piece in `bins`, contig name the pileup rows carry, a, b, shift)] -- rows of that contig with
        a <= position < b are joined to the piece at position - shift, its other rows are dropped."""
        from .dataload import DeviceRows, rows_from_table
        from .sharding import remap_split_rows

        contigs, self._ranges = {}, {}
        for b, cs in bins.items():
            begin = len(contigs)
            for name, seq in cs.items():
                if name in contigs:
                    raise ValueError(f"contig {name} is assigned to more than one bin")
                contigs[name] = seq
            self._ranges[b] = (begin, len(contigs))
        self.assembly = DeviceAssembly.from_sequences(contigs, device)
        self.mod_types = list(mod_types)
        d = self.assembly.device
        if isinstance(pileup, DeviceRows) or hasattr(pileup, "columns") or hasattr(pileup, "column_names") \
                or isinstance(pileup, PileupTable) or (isinstance(pileup, dict) and "position" in pileup):
            tables = [pileup]
        elif isinstance(pileup, dict):
            tables = list(pileup.values())  # the reference's partitioned_pileup: {(bin, mod_type): frame}
        else:
            tables = list(pileup)
        self.pileup = DevicePileup(self.assembly, len(self.mod_types), low_meth_threshold, high_meth_threshold).clear()
        self.rows, cache = [], {}
        for t in tables:
            rows = t if isinstance(t, DeviceRows) else rows_from_table(t, self.assembly.names, self.mod_types, d, cache,
                                                                             with_coverage=False)
            cid, mt = rows.contig_id, rows.mod_type
            if isinstance(t, DeviceRows) and (list(t.contig_names) != self.assembly.names or
                                              [str(m) for m in t.mod_types] != [str(m) for m in self.mod_types]):
                with torch.cuda.device(d):  # ids of the rows' contig / mod-type tables -> ids of this assembly / list
                    c_lut = np.fromiter((self.assembly.index.get(n, -1) for n in t.contig_names), dtype=np.int32,
                                        count=len(t.contig_names))
                    names = [str(m) for m in self.mod_types]
                    m_lut = np.full(256, 255, dtype=np.uint8)
                    for i, name in enumerate(t.mod_types):
                        if str(name) in names:
                            m_lut[i] = names.index(str(name))
                    cid = torch.from_numpy(np.append(c_lut, np.int32(-1))).to(d)[rows.contig_id.long()]  # -1 -> -1
                    mt = torch.from_numpy(m_lut).to(d)[rows.mod_type.long()]
            position = rows.position
            if pieces:
                index = self.assembly.index
                with torch.cuda.device(d):
                    cid, position = remap_split_rows(cid, position, [(index[carried], index[own], a, b, shift)
                                                                     for own, carried, a, b, shift in pieces])
            self.pileup.add_columns(cid, position, rows.strand, rows.fraction_mod, mt, sync=False)
            if keep_rows:  # device columns for the window step (growth.WindowPool) -- ids of THIS assembly
                self.rows.append(DeviceRows(self.assembly.names, self.mod_types, d, contig_id=cid, position=position,
                                            strand=rows.strand, mod_type=mt, fraction_mod=rows.fraction_mod,
                                            Nvalid_cov=rows.Nvalid_cov))
        torch.cuda.current_stream(d).synchronize()
        self.pileup.check_unique()
        first = self.rows[0] if self.rows else None
        self.contig_id = first.contig_id if first is not None else None
        self.mod_type_id = first.mod_type if first is not None else None
        self.table = tables[0] if len(tables) == 1 and isinstance(tables[0], PileupTable) else None

    @classmethod
    def from_device(cls, assembly: DeviceAssembly, pileup: DevicePileup, ranges: dict, mod_types, rows=None) -> "MultiBinScorer":
        """Scorer over state that already lives on the device: ranges = {bin name: (contig_begin, contig_end)};
        rows = optional list of dataload.DeviceRows (contig ids of `assembly`, mod types of `mod_types`) for the
        window step (growth.prepare_searches)."""
        self = cls.__new__(cls)
        self.assembly, self.pileup, self._ranges, self.mod_types = assembly, pileup, dict(ranges), list(mod_types)
        self.rows, self.contig_id, self.mod_type_id, self.table = list(rows or []), None, None, None
        return self

    def context(self, bin_name, mod_type) -> BinContext:
        begin, end = self._ranges[bin_name]
        return BinContext(self, bin_name, mod_type, begin, end, self.mod_types.index(mod_type))

    def score_batch_device(self, requests, out: torch.Tensor | None = None) -> torch.Tensor:
        """requests: [(BinContext or None, motifs)] -> int64 device tensor [total motifs, 4] (n_mod '+', n_nomod '+',
        n_mod '-', n_nomod '-') in request order; ONE scan launch, nothing synchronised.  Rows of requests whose
        context is None (bins that live on another rank, sharding.ShardedMultiBinScorer) stay zero.  `out` is
        accumulated into when given."""
        requests = [(ctx, list(motifs)) for ctx, motifs in requests]
        total = sum(len(ms) for _, ms in requests)
        d = self.assembly.device
        if out is None:
            with torch.cuda.device(d):
                out = torch.zeros((total, 4), dtype=torch.int64, device=d)
        live, base = [], 0
        for ctx, ms in requests:
            if ctx is not None and ms and ctx.tile_count > 0:
                live.append((ctx, ms, base))
            base += len(ms)
        if not live:
            return out
        progs = MotifPrograms([m for _, ms, _ in live for m in ms], d, strip=True)
        jobs = make_jobs(len(live))
        counts = np.fromiter((len(ms) for _, ms, _ in live), dtype=np.int64, count=len(live))
        jobs["motif_count"] = counts
        jobs["motif_begin"] = np.cumsum(counts) - counts
        jobs["out_base"] = np.fromiter((row for _, _, row in live), dtype=np.int64, count=len(live))
        for field in ("modtype_index", "tile_begin", "tile_count", "contig_begin", "contig_end"):  # whole columns at once
            jobs["modtype" if field == "modtype_index" else field] = np.fromiter(
                (getattr(ctx, field) for ctx, _, _ in live), dtype=np.int64, count=len(live))
        jobs["n_groups"] = 1  # group_mode 0: one posterior per motif
        return scan_count(self.assembly, self.pileup, progs, jobs, total, out=out)

    def score_batch(self, requests) -> list:
        """requests: [(BinContext, motifs)] -> [int64 array [n_motifs, 2] = (n_mod, n_nomod)] per request."""
        requests = [(ctx, list(motifs)) for ctx, motifs in requests]
        if not any(ms for _, ms in requests):
            return [np.zeros((0, 2), dtype=np.int64) for _ in requests]
        return split_counts(self.score_batch_device(requests).cpu().numpy(), requests)


def split_counts(c: np.ndarray, requests) -> list:
    """[total motifs, 4] strand-wise counts -> per request [n_motifs, 2] = (n_mod, n_nomod)."""
    counts = np.stack([c[:, 0] + c[:, 2], c[:, 1] + c[:, 3]], axis=1)
    out, at = [], 0
    for _, ms in requests:
        out.append(counts[at:at + len(ms)])
        at += len(ms)
    return out


def _scorer_for(pileup, contigs, low, high) -> BinScorer:
    key = ("bin", id(pileup), id(contigs), float(low), float(high), _pileup_fingerprint(pileup), _contigs_fingerprint(contigs))
    return _bin_cache.get(key, (pileup, contigs), lambda: BinScorer(pileup, contigs, low, high), _scorer_bytes)


def motif_model_bin(pileup, contigs, motif, model, low_meth_threshold, high_meth_threshold):
    """Posterior of one motif over all contigs of a bin; mutates and returns `model`
    (find_motifs_bin.py:1265-1283)."""
    if len(contigs) == 0:
        return model
    n_mod, n_nomod = _scorer_for(pileup, contigs, low_meth_threshold, high_meth_threshold).score([motif])[0]
    model.update(int(n_mod), int(n_nomod))
    return model


def motif_model_bin_many(pileup, contigs, motifs, low_meth_threshold, high_meth_threshold, model_factory=BetaBernoulliModel):
    """Batched motif_model_bin: one launch for all motifs, one fresh model per motif."""
    motifs = list(motifs)
    if len(contigs) == 0 or not motifs:
        return [model_factory() for _ in motifs]
    counts = _scorer_for(pileup, contigs, low_meth_threshold, high_meth_threshold).score(motifs)
    models = []
    for n_mod, n_nomod in counts:
        mdl = model_factory()
        mdl.update(int(n_mod), int(n_nomod))
        models.append(mdl)
    return models


def motif_model_contig(pileup, contig: str, prior, motif, low_meth_threshold=0.3, high_meth_threshold=0.7,
                       save_motif_positions=False):
    """Posterior update of one motif on one contig; `pileup` holds that contig's rows
    (find_motifs_bin.py:1285-1331).  Mutates and returns `prior`."""
    contig = sequence_of(contig)
    table = PileupTable.from_frame(pileup)
    m = as_motif(motif).new_stripped_motif()
    key = ("contig", id(pileup), id(contig), float(low_meth_threshold), float(high_meth_threshold),
           _pileup_fingerprint(pileup), len(contig))

    def build():
        t = PileupTable(None, table.position, table.strand, table.fraction_mod)
        return BinScorer(t, {"_": contig}, low_meth_threshold, high_meth_threshold)

    scorer = _bin_cache.get(key, (pileup, contig), build, _scorer_bytes)
    c = scorer.counts_by_strand([m]).cpu().numpy()[0]
    prior.update(int(c[0] + c[2]), int(c[1] + c[3]))
    if not save_motif_positions:
        return prior
    # position lists in pileup order (find_motifs_bin.py:1308-1329)
    strand = strand_codes(table.strand)
    high = table.fraction_mod >= high_meth_threshold
    low = table.fraction_mod <= low_meth_threshold
    pos = table.position
    fwd = methylated_motif_occourances(m, contig, pos[high & (strand == 0)], pos[low & (strand == 0)])
    rev = methylated_motif_occourances(m.reverse_compliment(), contig, pos[high & (strand == 1)],
                                       pos[low & (strand == 1)])
    return prior, {"index_meth_fwd": fwd[0], "index_nonmeth_fwd": fwd[1],
                   "index_meth_rev": rev[0], "index_nonmeth_rev": rev[1]}


def get_parent_scores(motif, pileup, contigs, low_meth_threshold, high_meth_threshold):
    """Score a motif against every parent (one constrained position reset to '.');
    one batched launch instead of 1 + K sequential scans (find_motifs_bin.py:1382-1433)."""
    m = as_motif(motif)
    split = m.split()
    parents, positions = [], []
    for i, base in enumerate(split):
        if i == m.mod_position or base in (".", "N"):
            continue
        toks = list(split)
        toks[i] = "."
        parents.append(Motif("".join(toks), m.mod_position))
        positions.append(i)
    models = motif_model_bin_many(pileup, contigs, [m] + parents, low_meth_threshold, high_meth_threshold)
    child = models[0]
    out = {}
    for parent, i, pm in zip(parents, positions, models[1:]):
        out[parent] = dict(motif_position=i, parent_model=pm, child_model=child,
                           score=predictive_evaluation_score(child, pm))
    return out
