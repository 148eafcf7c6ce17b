#!/usr/bin/env python
"""Generate tests/golden/binmodel_vectors.json by RUNNING the reference's own frame-taking functions.

    python tests/golden/generate_binmodel_golden.py        (needs /root/reference)

motif_model_contig / motif_model_bin / get_parent_scores (find_motifs_bin.py:1265-1331, 1382-1433) and the loader
filters filter_pileup / filter_pileup_minimummod_frequency (dataload.py:191-226) take polars frames, and polars is not
in this image.  oracle/minipolars.py implements the handful of polars calls those functions make on numpy columns, so
the UNMODIFIED reference functions run here (imported through oracle/ref_shim.py).  Recorded: their results on a seeded
synthetic bin (inputs are regenerated from the seed by nanomotif_b200.synth, so only the seed travels) --
(n_mod, n_nomod) per motif for the bin and per contig, the four position lists of save_motif_positions, the parent
table of get_parent_scores, and the rows the two filters keep.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import minipolars as mp  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

nm = load_reference()
mp.install(nm)
fmb, Motif, B = nm.find_motifs_bin, nm.motif.Motif, nm.model.BetaBernoulliModel

SPEC = dict(seed=77, contig_lengths=[40000, 17000, 900, 61], gc=0.45, n_rate=2e-4, depth=12, mod_types=["a", "m"],
            planted=[["GATC", 1, "a"], ["CC[AT]GG", 1, "m"], ["GCAC......GTT", 2, "a"]], low=0.3, high=0.7)
MOTIFS = {"a": [("GATC", 1), ("A", 0), ("....GATC..", 5), ("GCAC......GTT", 2), ("G[AG].GAAG[CT]", 5), ("[AG]A[CT]", 1),
                ("A" + "." * 30 + "T", 0), ("TCGA", 3), ("AA", 0), ("AA", 1)],
          "m": [("CC[AT]GG", 1), ("C", 0), ("GC[ACT]C", 3), ("CG", 0)]}
THRESHOLDS = [(0.3, 0.7), (0.0, 1.0), (0.5, 0.5)]


def build_inputs(spec=SPEC):
    from nanomotif_b200 import synth

    rng = np.random.default_rng(spec["seed"])
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "mod_type", "fraction_mod", "Nvalid_cov")}
    for i, L in enumerate(spec["contig_lengths"]):
        seq = synth.random_sequence(rng, L, spec["gc"], spec["n_rate"])
        name = f"contig_{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=spec["depth"], mod_types=tuple(spec["mod_types"]),
                               planted=[tuple(x) for x in spec["planted"]])
        n = len(p["position"])
        cols["contig"].append(np.full(n, name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["mod_type"].append(np.array(spec["mod_types"], dtype=object)[p["mod_type"]])
        cols["fraction_mod"].append(p["fraction_mod"])
        cols["Nvalid_cov"].append(p["Nvalid_cov"])
    return contigs, {k: np.concatenate(v) for k, v in cols.items()}


def digest(a) -> dict:
    """A position list as (length, sha1 of its int64 bytes, first values): order and content are pinned, the file stays
    small (the list of every methylated 'A' of a contig is long)."""
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int64)
    return dict(n=int(a.size), sha1=hashlib.sha1(a.tobytes()).hexdigest(), head=a[:6].tolist())


def main():
    contigs, pile = build_inputs()
    frame = mp.DataFrame(pile)
    seqs = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}
    out = dict(spec=SPEC, reference_version=getattr(nm, "__version__", "1.1.2"), n_rows=int(frame.height),
               checksum=int(pile["position"].sum()), bin=[], contig=[], parents=[], filters=[])
    for mt, motifs in MOTIFS.items():
        sub = frame.filter(mp.col("mod_type") == mt)
        for low, high in THRESHOLDS:
            for s, p in motifs:
                m = fmb.motif_model_bin(sub, seqs, Motif(s, p), B(), low, high)
                out["bin"].append(dict(mod_type=mt, motif=s, mod_pos=p, low=low, high=high, counts=[int(x) for x in m.get_raw_counts()]))
        for s, p in motifs[:5]:
            for name in contigs:
                pc = sub.filter(mp.col("contig") == name)
                m, pos = fmb.motif_model_contig(pc, contigs[name], B(), Motif(s, p), save_motif_positions=True)
                out["contig"].append(dict(mod_type=mt, motif=s, mod_pos=p, contig=name, counts=[int(x) for x in m.get_raw_counts()],
                                          positions={k: digest(v) for k, v in pos.items()}))
    for mt, s, p in (("a", "GATC", 1), ("a", "..GCAC......GTT.", 4), ("m", "CC[AT]GG", 1), ("a", "G[AG].GAAG[CT]", 5)):
        sub = frame.filter(mp.col("mod_type") == mt)
        res = fmb.get_parent_scores(Motif(s, p), sub, seqs, 0.3, 0.7)
        out["parents"].append(dict(mod_type=mt, motif=s, mod_pos=p, parents=[
            dict(parent=k.string, mod_pos=int(k.mod_position), motif_position=int(v["motif_position"]),
                 parent_counts=[int(x) for x in v["parent_model"].get_raw_counts()],
                 child_counts=[int(x) for x in v["child_model"].get_raw_counts()], score=float(v["score"])) for k, v in res.items()]))
    # the loader filters: row ids that survive (a row id column rides along)
    frame_id = mp.DataFrame(dict(pile, row=np.arange(frame.height)))
    for cov in (5, 11, 40):
        kept = nm.dataload.filter_pileup(frame_id, min_coverage=cov)
        out["filters"].append(dict(kind="coverage", min_coverage=cov, n_kept=int(kept.height), row_sum=int(kept["row"].to_numpy().sum())))
    for thr, freq, mods in ((0.7, 0.0001, 50), (0.7, 0.01, 5), (0.9, 0.002, 0), (0.7, 0.5, 0)):
        kept = nm.dataload.filter_pileup_minimummod_frequency(frame_id, thr, freq, mods)
        groups = sorted({f"{c}_{m}" for c, m in zip(kept["contig"].to_list(), kept["mod_type"].to_list())})
        out["filters"].append(dict(kind="min_mod_frequency", methylation_threshold=thr, min_mod_frequency=freq,
                                   min_mods_pr_contig=mods, n_kept=int(kept.height), row_sum=int(kept["row"].to_numpy().sum()),
                                   groups=groups, columns=kept.columns))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "binmodel_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, os.path.getsize(path), "bytes;", len(out["bin"]), "bin vectors,", len(out["contig"]), "contig vectors")


if __name__ == "__main__":
    main()
