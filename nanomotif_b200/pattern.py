"""Contig x motif methylation-pattern table (K5) -- the operator `nanomotif detect_contamination` /
`include_contigs` obtain from the external Rust package ``epymetheus`` (call site
nanomotif/main.py:167-178; output schema nanomotif/main.py:157-161).

The Rust source is not part of the reference tree, so the semantics are fixed by the written spec in
DESIGN.md (section 4, K5) and SURVEY.md 8c; parity is "exact against oracle/restate.py::
methylation_pattern, unpinned against epimetheus-py 0.7.5".

For every motif string ``<IUPAC>_<mod_type>_<mod_position>`` and contig: occurrences of the motif on
'+' and of its reverse complement on '-' (IUPAC letters with SET semantics) are joined to the pileup
rows of that mod type at the modified base; rows need n_valid_cov >= min_valid_read_coverage and
n_valid_cov / (n_valid_cov + n_diff) >= min_valid_cov_to_diff_fraction.  Per (contig, motif):
n_motif_obs = kept occurrences, mean_read_cov = mean n_valid_cov, methylation_value = median of
n_mod / n_valid_cov (Median) or sum n_mod / sum n_valid_cov (WeightedMean).  Pairs without
observations are omitted.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr
from .device import DeviceAssembly, MotifPrograms, _stream, _to_device
from .motif import Motif
from .pileup import PileupTable, strand_codes

COLUMNS = ["contig", "motif", "mod_type", "mod_position", "methylation_value", "mean_read_cov", "n_motif_obs"]


class MethylationOutput(enum.Enum):
    Median = "median"
    WeightedMean = "weighted_mean"


def parse_motif_spec(spec: str) -> tuple[str, str, int]:
    """'GATC_a_1' -> ('GATC', 'a', 1)  (nanomotif/main.py:136-140 builds these strings)."""
    iupac, mod_type, pos = spec.rsplit("_", 2)
    return iupac, mod_type, int(pos)


def _planes(asm: DeviceAssembly, regex_motif: Motif):
    progs = MotifPrograms([regex_motif], asm.device, strip=False)
    view = asm.view()
    out = []
    for strand in (0, 1):
        plane = torch.empty(asm.n_words, dtype=torch.int32, device=asm.device)
        check(lib.nmb_match_plane(C.byref(view), ptr(progs.programs), 0, strand, progs.max_len, 0, asm.n_tiles,
                                  ptr(plane), _stream()), "nmb_match_plane")
        out.append(plane)
    return out


def methylation_pattern(pileup, assembly, motifs, threads: int = 1, min_valid_read_coverage: int = 3,
                        batch_size: int = 1000, min_valid_cov_to_diff_fraction: float = 0.8, output: str | None = None,
                        allow_assembly_pileup_mismatch: bool = True,
                        output_type: MethylationOutput = MethylationOutput.Median, device=None):
    """Same keyword arguments as ``epymetheus.methylation_pattern``.  `pileup`: path to a bedMethyl file or a
    table with n_mod / n_diff in ``extra``; `assembly`: FASTA path or {name: sequence}.  Returns a pandas
    DataFrame with the columns of nanomotif/main.py:157-161 (and writes it as TSV to `output`).
    `threads` and `batch_size` are accepted for signature compatibility (the GPU needs neither)."""
    import pandas as pd

    from . import dataload

    contigs = dataload.load_fasta(assembly) if isinstance(assembly, str) else {k: (v if isinstance(v, str) else v.sequence) for k, v in assembly.items()}
    table = dataload.load_pileup(pileup, with_counts=True) if isinstance(pileup, str) else PileupTable.from_frame(pileup)
    if "n_mod" not in table.extra or "n_diff" not in table.extra:
        raise KeyError("methylation_pattern needs the n_mod and n_diff pileup columns")
    asm = DeviceAssembly.from_sequences(contigs, device)
    d = asm.device
    names = np.asarray(table.contig).astype(str)
    uniq, inv = np.unique(names, return_inverse=True)
    lut = np.fromiter((asm.index.get(u, -1) for u in uniq), dtype=np.int64, count=len(uniq))
    cid_all = lut[inv] if len(uniq) else np.zeros(0, dtype=np.int64)
    if not allow_assembly_pileup_mismatch and (cid_all < 0).any():
        raise ValueError("pileup contains contigs that are absent from the assembly")
    pos_all = np.asarray(table.position, dtype=np.int64)
    cov_all = np.asarray(table.Nvalid_cov, dtype=np.int64)
    nmod_all = np.asarray(table.extra["n_mod"], dtype=np.int64)
    diff_all = np.asarray(table.extra["n_diff"], dtype=np.int64)
    strand_all = strand_codes(table.strand)
    mt_all = np.asarray(table.mod_type).astype(str)
    in_range = (cid_all >= 0) & (pos_all >= 0) & (pos_all < asm.lengths[np.maximum(cid_all, 0)])
    ok = in_range & (cov_all >= min_valid_read_coverage)
    with np.errstate(divide="ignore", invalid="ignore"):
        ok &= (cov_all / (cov_all + diff_all)) >= min_valid_cov_to_diff_fraction

    rows_by_mod: dict[str, tuple] = {}
    out_rows = []
    with torch.cuda.device(d):
        for spec in motifs:
            iupac, mod_type, mod_pos = parse_motif_spec(spec)
            if mod_type not in rows_by_mod:
                sel = np.flatnonzero(ok & (mt_all == mod_type))
                sel = sel[np.lexsort((pos_all[sel], cid_all[sel]))]  # the kernels want contig-sorted rows
                rows_by_mod[mod_type] = (
                    _to_device(asm.starts[cid_all[sel]] + pos_all[sel], d), _to_device(strand_all[sel], d),
                    _to_device(cid_all[sel].astype(np.int32), d), _to_device(nmod_all[sel].astype(np.int32), d),
                    _to_device(cov_all[sel].astype(np.int32), d), len(sel))
            gpos, st, cid, nmod, cov, n_rows = rows_by_mod[mod_type]
            fwd, rev = _planes(asm, Motif(iupac, mod_pos).from_iupac())
            nc = asm.n_contigs
            stats = torch.empty((nc, 3), dtype=torch.int64, device=d)
            offsets = torch.empty(nc + 1, dtype=torch.int64, device=d)
            cursor = torch.empty(nc, dtype=torch.int32, device=d)
            check(lib.nmb_pattern_stats(ptr(gpos), ptr(st), ptr(cid), ptr(nmod), ptr(cov), n_rows, ptr(fwd), ptr(rev), nc,
                                        ptr(stats), ptr(offsets), ptr(cursor), _stream()), "nmb_pattern_stats")
            s = stats.cpu().numpy()
            n_obs = s[:, 0]
            if output_type == MethylationOutput.Median:
                total = int(n_obs.sum())
                fractions = torch.empty(max(1, total), dtype=torch.float64, device=d)
                median = torch.empty(nc, dtype=torch.float64, device=d)
                check(lib.nmb_pattern_median(ptr(gpos), ptr(st), ptr(cid), ptr(nmod), ptr(cov), n_rows, ptr(fwd), ptr(rev),
                                             nc, ptr(offsets), ptr(cursor), ptr(fractions), ptr(median), _stream()),
                      "nmb_pattern_median")
                value = median.cpu().numpy()
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    value = s[:, 1] / s[:, 2]
            for c in np.flatnonzero(n_obs > 0):
                out_rows.append((asm.names[c], iupac, mod_type, mod_pos, float(value[c]), float(s[c, 2] / s[c, 0]),
                                 int(n_obs[c])))
    df = pd.DataFrame(out_rows, columns=COLUMNS).astype({"mod_position": np.int8, "n_motif_obs": np.int32,
                                                         "methylation_value": np.float64, "mean_read_cov": np.float64})
    if output is not None:
        df.to_csv(output, sep="\t", index=False)
    return df
