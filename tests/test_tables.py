"""Post-scan table products against the oracle's pandas restatement of the reference pipeline and
the reference's own expectations.  CPU only."""
import numpy as np
import pandas as pd

from nanomotif_b200 import tables
from oracle import restate as O


def test_motif_type_reference_cases():
    # nanomotif/utils.py:26-34; examples from the shipped bin-motifs tables (nanomotif/datasets/*bin-motifs.tsv)
    cases = {"GATC": "palindrome", "CCWGG": "palindrome", "GCACNNNNNNGTT": "bipartite", "AACNNNNNNGTGC": "bipartite",
             "GRNGAAGY": "non-palindrome", "ACNNGTNNAC": "ambiguous", "CCGG": "palindrome", "GANTC": "palindrome",
             "ATNNNNNAT": "bipartite", "RGATCY": "palindrome", "A": "non-palindrome"}
    for m, want in cases.items():
        assert tables.motif_type(m) == want == O.motif_type(m), m
    assert tables.reverse_compliment("GRNGAAGY") == "RCTTCNYC"


def test_write_motif_formatted(tmp_path):
    recs = [dict(reference="bin2", motif_iupac="GATC", mod_position_iupac=1, mod_type="a", n_mod=900, n_nomod=10, junk=1),
            dict(reference="bin1", motif_iupac="GCACNNNNNNGTT", mod_position_iupac=2, mod_type="a", n_mod=50, n_nomod=3, junk=2),
            dict(reference="bin1", motif_iupac="CCWGG", mod_position_iupac=1, mod_type="m", n_mod=70, n_nomod=1, junk=3),
            dict(reference="bin1", motif_iupac="AACNNNNNNGTGC", mod_position_iupac=1, mod_type="a", n_mod=48, n_nomod=5, junk=4)]
    path = tmp_path / "bin-motifs.tsv"
    tables.write_motif_formatted(recs, str(path))
    got = pd.read_csv(path, sep="\t")
    # motif.py:899-926: columns, renames, motif_type, sort by reference / mod_type / motif
    assert list(got.columns) == ["reference", "motif", "mod_position", "mod_type", "n_mod", "n_nomod", "motif_type"]
    assert got["motif"].tolist() == ["AACNNNNNNGTGC", "GCACNNNNNNGTT", "CCWGG", "GATC"]
    assert got["motif_type"].tolist() == ["bipartite", "bipartite", "palindrome", "palindrome"]
    comp = [dict(r, motif_iupac_complement=tables.reverse_compliment(r["motif_iupac"]), mod_position_iupac_complement=0,
                 n_mod_complement=1, n_nomod_complement=2) for r in recs]
    got = tables.motif_formatted_frame(comp)
    assert list(got.columns)[-4:] == ["motif_complement", "mod_position_complement", "n_mod_complement", "n_nomod_complement"]
    assert got["motif_complement"].tolist() == ["GCACNNNNNNGTT", "AACNNNNNNGTGC", "CCWGG", "GATC"]


def test_bin_feature_matrix_equals_reference_pipeline():
    rng = np.random.default_rng(4)
    n_motifs, n_contigs = 9, 40
    names = [f"contig_{i}" for i in range(n_contigs)]
    specs = [("GATC", "a", 1), ("CCWGG", "m", 1), ("GRNGAAGY", "a", 5), ("CCGG", "21839", 0), ("GATC", "m", 3),
             ("ACNNGTNNAC", "a", 0), ("TTAA", "a", 3), ("GCACNNNNNNGTT", "a", 2), ("AAAA", "a", 0)]
    motif_mods = [f"{m}_{t}_{p}" for m, t, p in specs]
    obs = rng.integers(0, 40, size=(n_motifs, n_contigs)) * (rng.random((n_motifs, n_contigs)) < 0.6)
    obs[8] = 0  # a motif nobody observes
    cov_sum = obs * rng.integers(1, 30, size=obs.shape)
    stats = np.stack([obs, (cov_sum * rng.random(obs.shape)).astype(np.int64), cov_sum], axis=2).astype(np.int64)
    value = np.where(obs > 0, rng.random(obs.shape), np.nan)
    contig_bin = {n: f"bin{(i * 7) % 5}" for i, n in enumerate(names) if i % 6 != 5}  # some contigs unbinned
    mi, ci = np.nonzero(obs > 0)
    frame = pd.DataFrame({"contig": np.array(names, dtype=object)[ci], "motif": [specs[m][0] for m in mi],
                          "mod_type": [specs[m][1] for m in mi], "mod_position": [specs[m][2] for m in mi],
                          "methylation_value": value[mi, ci], "mean_read_cov": stats[mi, ci, 2] / stats[mi, ci, 0],
                          "n_motif_obs": stats[mi, ci, 0].astype(np.int32)})
    for thr in (24.0, 0.0, 200.0):
        want_c, want_m, want_f = O.binnary_matrix(frame, contig_bin, thr)
        got_c, got_m, got_f = tables.bin_feature_matrix(stats, value, names, motif_mods, contig_bin, thr)
        assert got_c.tolist() == want_c.tolist() and got_f.tolist() == want_f.tolist()
        np.testing.assert_allclose(got_m, want_m, rtol=1e-13, atol=0)
        assert got_m.shape[0] > 5 and got_m.shape[1] >= 6
