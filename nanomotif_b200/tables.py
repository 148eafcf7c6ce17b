"""Post-scan table products: the files and matrices users consume, built from the
count tables the kernels return.

    motif_type                    nanomotif/utils.py:15-34
    write_motif_formatted         nanomotif/motif.py:899-926   (bin-motifs.tsv)
    contig_methylation_frame      the frame epymetheus.methylation_pattern returns (nanomotif/main.py:157-161)
                                  after the coverage filter of nanomotif/main.py:192-193
    bin_feature_matrix            add_bin + impute_contig_methylation_within_bin + create_matrix
                                  (nanomotif/binnary/data_processing.py:174-213,255-269), straight from the dense
                                  [motif, contig] arrays of pattern.pattern_table -- no frame in between

Host code (numpy / pandas): these tables are O(contigs x motifs) and the consumers (PCA, classifiers, TSV)
are host libraries.
"""
from __future__ import annotations

import re

import numpy as np

# nanomotif/constants.py:14-20
COMPLEMENT = {"A": "T", "T": "A", "G": "C", "C": "G", "N": "N", "R": "Y", "Y": "R", "S": "S", "W": "W", "K": "M", "M": "K",
              "B": "V", "D": "H", "H": "D", "V": "B", ".": ".", "[": "]", "]": "["}


def reverse_compliment(seq: str) -> str:
    """nanomotif/seq.py:645-647 (the reference's spelling)."""
    return "".join(COMPLEMENT[b] for b in reversed(seq))


def has_n_character_stretches_of_length_m(sequence: str, n: int, m: int, character: str = "N") -> bool:
    """nanomotif/utils.py:15-24: at least n runs of `character` of length >= m."""
    return len(re.findall(rf"({character}){{{m},}}", sequence)) >= n


def motif_type(motif_str: str) -> str:
    """nanomotif/utils.py:26-34."""
    if has_n_character_stretches_of_length_m(motif_str, 2, 2, "N"):
        return "ambiguous"
    if re.search(r"(N){3,}", motif_str):
        return "bipartite"
    if reverse_compliment(motif_str) == motif_str:
        return "palindrome"
    return "non-palindrome"


_COMPLEMENT_COLS = ("motif_iupac_complement", "mod_position_iupac_complement", "n_mod_complement", "n_nomod_complement")


def motif_formatted_frame(records):
    """The frame write_motif_formatted writes (motif.py:899-923).  `records`: a pandas DataFrame or a list of dicts
    with reference, motif_iupac, mod_position_iupac, mod_type, n_mod, n_nomod (+ the four *_complement columns)."""
    import pandas as pd

    df = records if isinstance(records, pd.DataFrame) else pd.DataFrame(list(records))
    out = df[["reference", "motif_iupac", "mod_position_iupac", "mod_type", "n_mod", "n_nomod"]].rename(
        columns={"motif_iupac": "motif", "mod_position_iupac": "mod_position"})
    out["motif_type"] = [motif_type(m) for m in out["motif"]]
    if all(c in df.columns for c in _COMPLEMENT_COLS):
        out["motif_complement"] = df["motif_iupac_complement"].values
        out["mod_position_complement"] = df["mod_position_iupac_complement"].values
        out["n_mod_complement"] = df["n_mod_complement"].values
        out["n_nomod_complement"] = df["n_nomod_complement"].values
    return out.sort_values(["reference", "mod_type", "motif"], kind="stable").reset_index(drop=True)


def write_motif_formatted(records, file_path: str) -> None:
    motif_formatted_frame(records).to_csv(file_path, sep="\t", index=False)


def bin_feature_matrix(stats: np.ndarray, value: np.ndarray, contig_names, motif_mods, contig_bin: dict,
                       methylation_threshold: float = 24.0):
    """(contig_names, matrix float64 [n_rows, n_features], feature names) of the binned contigs, from the dense
    K5 arrays stats [n_motifs, n_contigs, 3] / value [n_motifs, n_contigs] (pattern.pattern_table).

    Follows the reference pipeline cell by cell: keep cells with n_motif_obs * mean_read_cov >=
    methylation_threshold (main.py:192-193), attach bins and drop unbinned contigs (data_processing.py:174-192),
    per (bin, motif_mod) mean = sum(value * n_obs) / sum(n_obs) (:193-199), a contig with at least one kept cell
    gets every motif_mod its bin has, own value else the bin mean (:201-211); pivot contig x motif_mod with the
    features sorted, missing -> 0, rows in (bin, contig) order (:255-269)."""
    stats = np.asarray(stats)
    value = np.asarray(value, dtype=np.float64)
    names = np.asarray(contig_names, dtype=object)
    motif_mods = np.asarray(motif_mods, dtype=object)
    n_obs = stats[:, :, 0].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean_cov = stats[:, :, 2] / n_obs
        keep = (stats[:, :, 0] > 0) & (n_obs * mean_cov >= methylation_threshold)
    bins = np.array([contig_bin.get(n) for n in names], dtype=object)
    binned = np.array([b is not None for b in bins])
    keep &= binned[None, :]
    bin_names, bin_of = np.unique(bins[binned].astype(str), return_inverse=True)
    bin_idx = np.full(len(names), -1, dtype=np.int64)
    bin_idx[binned] = bin_of
    nb, nm = len(bin_names), len(motif_mods)
    # per (bin, motif_mod): weighted mean over the kept cells
    w = np.where(keep, n_obs, 0.0)
    num = np.zeros((nm, nb))
    den = np.zeros((nm, nb))
    cols = np.flatnonzero(binned)
    np.add.at(num.T, bin_idx[cols], (np.where(keep, value, 0.0) * w)[:, cols].T)
    np.add.at(den.T, bin_idx[cols], w[:, cols].T)
    bin_has = den > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        bin_mean = num / den
    rows = np.flatnonzero(keep.any(axis=0))  # contigs with at least one kept cell
    rows = rows[np.lexsort((names[rows].astype(str), bins[rows].astype(str)))]
    feat_order = np.argsort(motif_mods.astype(str), kind="stable")
    feat_used = feat_order[bin_has[feat_order].any(axis=1)]  # motif_mods that appear in the imputed frame
    b = bin_idx[rows]
    own = keep[:, rows]
    cell = np.where(own, value[:, rows], np.where(bin_has[:, b], bin_mean[:, b], 0.0))
    return names[rows], np.ascontiguousarray(cell[feat_used].T), motif_mods[feat_used]


def bin_feature_matrix_device(stats, value, contig_names, motif_mods, contig_bin: dict, methylation_threshold: float = 24.0):
    """bin_feature_matrix on the DEVICE (K9: nmb_bin_means + nmb_bin_matrix), straight from the K5 outputs as device
    tensors -- stats int64 [n_motifs, n_contigs, 3], value float64 [n_motifs, n_contigs] or None (weighted mean) --
    as pattern.pattern_table(..., on_device=True) returns them.  Returns (contig names, matrix as a device tensor
    [n_rows, n_features] float64, feature names); same cells, bit for bit, as bin_feature_matrix.  Only the two string
    sorts (rows by (bin, contig name), features by motif_mod) run on the host, on two small flag arrays."""
    import torch

    from ._lib import check, lib, ptr
    from .device import _stream, _to_device

    d = stats.device
    names = np.asarray(contig_names, dtype=object)
    motif_mods = np.asarray(motif_mods, dtype=object)
    nm, nc = int(stats.shape[0]), int(stats.shape[1])
    bins = np.array([contig_bin.get(n) for n in names], dtype=object)
    binned = np.array([b is not None for b in bins], dtype=bool)
    bin_names, bin_of = np.unique(bins[binned].astype(str), return_inverse=True)
    bin_idx = np.full(nc, -1, dtype=np.int32)
    bin_idx[binned] = bin_of
    nb = len(bin_names)
    order = np.flatnonzero(binned)
    order = order[np.argsort(bin_idx[order], kind="stable")].astype(np.int32)  # grouped by bin, contigs ascending inside
    bin_off = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(np.bincount(bin_idx[binned], minlength=nb), out=bin_off[1:])
    with torch.cuda.device(d):
        stats = stats.contiguous()
        value = None if value is None else value.contiguous()
        keep = torch.empty((nm, nc), dtype=torch.uint8, device=d)
        bin_mean = torch.empty((nm, max(nb, 1)), dtype=torch.float64, device=d)
        bin_has = torch.zeros((nm, max(nb, 1)), dtype=torch.uint8, device=d)
        contig_has = torch.empty(nc, dtype=torch.uint8, device=d)
        off_d, ord_d, bidx_d = _to_device(bin_off, d), _to_device(order if len(order) else np.zeros(1, np.int32), d), _to_device(bin_idx, d)
        check(lib.nmb_bin_means(ptr(stats), ptr(value), nm, nc, ptr(off_d), ptr(ord_d), nb, float(methylation_threshold),
                                ptr(keep), ptr(bin_mean), ptr(bin_has), ptr(contig_has), _stream()), "nmb_bin_means")
        rows = np.flatnonzero(contig_has.cpu().numpy())
        rows = rows[np.lexsort((names[rows].astype(str), bins[rows].astype(str)))].astype(np.int32)
        feat_order = np.argsort(motif_mods.astype(str), kind="stable")
        has_any = (bin_has[:, :nb] != 0).any(dim=1).cpu().numpy().astype(bool) if nb else np.zeros(nm, dtype=bool)
        feats = feat_order[has_any[feat_order]].astype(np.int32)
        matrix = torch.empty((len(rows), len(feats)), dtype=torch.float64, device=d)
        if len(rows) and len(feats):
            rows_d, feats_d = _to_device(rows, d), _to_device(feats, d)
            check(lib.nmb_bin_matrix(ptr(stats), ptr(value), ptr(keep), ptr(bin_mean), ptr(bin_has), ptr(bidx_d), nc, max(nb, 1),
                                     ptr(rows_d), len(rows), ptr(feats_d), len(feats), ptr(matrix), _stream()),
                  "nmb_bin_matrix")
    return names[rows], matrix, motif_mods[feats]
