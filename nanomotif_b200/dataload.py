"""Loaders and pileup filters of the scoring path (mirror of nanomotif/dataload.py and fasta.py).

    load_pileup                              nanomotif/dataload.py:72-100
    filter_pileup                            nanomotif/dataload.py:191-200
    filter_pileup_minimummod_frequency       nanomotif/dataload.py:202-226
    filter_pileup_adjacency_filter           nanomotif/dataload.py:228-247
    load_fasta                               nanomotif/fasta.py:35-49

The text is parsed on the host (pyarrow CSV reader); the filters run as CUDA kernels on the columnar
arrays and return a filtered `PileupTable`.  Row order: the input order is kept (the reference returns
the adjacency-filtered rows grouped by (contig, strand) and sorted by position; nothing downstream
depends on that order).
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib, ptr
from .device import _require_cuda, _stream, _to_device
from .pileup import PileupTable, strand_codes

# bedMethyl columns kept by the reference (1-based 1,2,4,6,10,11; dataload.py:84)
_BED_COLUMNS = {0: "contig", 1: "position", 3: "mod_type", 5: "strand", 9: "Nvalid_cov", 10: "percent_modified",
                11: "n_mod", 16: "n_diff"}


def load_pileup(path: str, with_counts: bool = False) -> PileupTable:
    """modkit pileup (18-column bedMethyl, tab separated, optionally gzip) -> PileupTable with
    fraction_mod = column 11 / 100 (dataload.py:85).  with_counts keeps n_mod (column 12) and n_diff
    (column 17) for the methylation-pattern table."""
    import pyarrow as pa
    import pyarrow.csv as pacsv

    names = [f"column_{i + 1}" for i in range(18)]
    types = {"column_1": pa.string(), "column_2": pa.int64(), "column_4": pa.string(), "column_6": pa.string(),
             "column_10": pa.int64(), "column_11": pa.float64(), "column_12": pa.int64(), "column_17": pa.int64()}
    keep = ["column_1", "column_2", "column_4", "column_6", "column_10", "column_11"]
    if with_counts:
        keep += ["column_12", "column_17"]
    table = pacsv.read_csv(
        path,
        read_options=pacsv.ReadOptions(column_names=names),
        parse_options=pacsv.ParseOptions(delimiter="\t"),
        convert_options=pacsv.ConvertOptions(column_types=types, include_columns=keep, null_values=["NA", "null"],
                                             strings_can_be_null=True),
    )
    if table.num_rows == 0:
        raise SystemExit("Pileup is empty after initial load")  # the reference prints and sys.exit(1)s
    col = lambda n: table.column(n).to_numpy(zero_copy_only=False)
    extra = {}
    if with_counts:
        extra = {"n_mod": col("column_12").astype(np.int64), "n_diff": col("column_17").astype(np.int64)}
    return PileupTable(col("column_1").astype(object), col("column_2").astype(np.int64), col("column_6").astype(object),
                       col("column_11").astype(np.float64) / 100, col("column_4").astype(object),
                       col("column_10").astype(np.int64), extra)


def load_fasta(path: str, trim_names: bool = False, trim_character: str = " ") -> dict[str, str]:
    """FASTA (optionally gzip) -> {name: upper-case sequence} (fasta.py:35-49 + seq.py:55)."""
    import gzip

    opener = gzip.open if str(path).endswith(".gz") else open
    out, name, parts = {}, None, []
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    out[name] = "".join(parts).upper()
                name = line[1:].strip()
                name = name.split(trim_character)[0] if trim_names else name.split()[0]
                parts = []
            else:
                parts.append(line.strip())
    if name is not None:
        out[name] = "".join(parts).upper()
    return out


def _codes(values) -> tuple[np.ndarray, np.ndarray]:
    """Dense int32 codes (order of first appearance is irrelevant) and the unique values."""
    uniq, inv = np.unique(np.asarray(values).astype(str), return_inverse=True)
    return inv.astype(np.int32), uniq


def _mask_to_host(keep: torch.Tensor) -> np.ndarray:
    return keep.cpu().numpy().astype(bool)


def filter_pileup(pileup, min_modtype_fraction: float = 0.3, min_coverage: int = 5, device=None) -> PileupTable:
    """Keep positions with Nvalid_cov > min_coverage (dataload.py:191-200; min_modtype_fraction is
    accepted and ignored exactly like the reference)."""
    t = PileupTable.from_frame(pileup)
    d = _require_cuda(device)
    n = len(t)
    with torch.cuda.device(d):
        cov = _to_device(np.asarray(t.Nvalid_cov, dtype=np.int64), d)
        keep = torch.empty(n, dtype=torch.uint8, device=d)
        check(lib.nmb_filter_coverage(ptr(cov), n, int(min_coverage), ptr(keep), _stream()), "nmb_filter_coverage")
    return t.take(_mask_to_host(keep))


def filter_pileup_minimummod_frequency(pileup, methylation_threshold: float = 0.7, min_mod_frequency: float = 0.0001,
                                       min_mods_pr_contig: int = 50, device=None) -> PileupTable:
    """Keep contig_mods with enough methylated positions (dataload.py:202-226)."""
    t = PileupTable.from_frame(pileup)
    d = _require_cuda(device)
    n = len(t)
    cid, contigs = _codes(t.contig)
    mid, mods = _codes(t.mod_type)
    group = (cid.astype(np.int64) * len(mods) + mid).astype(np.int32)
    n_groups = max(1, len(contigs) * len(mods))
    with torch.cuda.device(d):
        g = _to_device(group, d)
        fr = _to_device(np.asarray(t.fraction_mod, dtype=np.float64), d)
        counts = torch.empty(2 * n_groups, dtype=torch.int64, device=d)
        keep = torch.empty(n, dtype=torch.uint8, device=d)
        check(lib.nmb_filter_min_mod_frequency(ptr(g), ptr(fr), n, n_groups, float(methylation_threshold),
                                               float(min_mod_frequency), int(min_mods_pr_contig), ptr(counts),
                                               ptr(keep), _stream()), "nmb_filter_min_mod_frequency")
    return t.take(_mask_to_host(keep))


def filter_pileup_adjacency_filter(pileup, methylation_threshold: float = 0.7, adjacency_distance: int = 8,
                                   device=None) -> PileupTable:
    """Drop a methylated position when a larger fraction exists within +-adjacency_distance on the
    same (contig, strand), over all mod types (dataload.py:228-247)."""
    t = PileupTable.from_frame(pileup)
    d = _require_cuda(device)
    n = len(t)
    cid, _ = _codes(t.contig) if t.contig is not None else (np.zeros(n, dtype=np.int32), None)
    pos = np.asarray(t.position, dtype=np.int64)
    order = None
    if n > 1 and not np.all((cid[1:] > cid[:-1]) | ((cid[1:] == cid[:-1]) & (pos[1:] >= pos[:-1]))):
        order = np.lexsort((pos, cid))  # the kernel needs (contig, position) order
    sel = (lambda a: a[order]) if order is not None else (lambda a: a)
    with torch.cuda.device(d):
        c_d = _to_device(sel(cid), d)
        p_d = _to_device(sel(pos), d)
        s_d = _to_device(sel(strand_codes(t.strand)), d)
        f_d = _to_device(sel(np.asarray(t.fraction_mod, dtype=np.float64)), d)
        keep = torch.empty(n, dtype=torch.uint8, device=d)
        flag = torch.zeros(1, dtype=torch.int32, device=d)
        check(lib.nmb_filter_adjacency(ptr(c_d), ptr(p_d), ptr(s_d), ptr(f_d), n, float(methylation_threshold),
                                       int(adjacency_distance), ptr(keep), ptr(flag), _stream()), "nmb_filter_adjacency")
        if int(flag.item()):
            raise RuntimeError("nmb_filter_adjacency: rows were not sorted by (contig, position)")
    k = _mask_to_host(keep)
    if order is not None:
        inv = np.empty(n, dtype=bool)
        inv[order] = k
        k = inv
    return t.take(k)
