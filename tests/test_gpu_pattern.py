"""K5 (contig x motif methylation-pattern table) against the written-spec oracle.  Parity vs the external
Rust reference is unpinned (its source is not in the tree); vs the spec it is exact / 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O


def _data(seed=9):
    from nanomotif_b200 import synth
    from nanomotif_b200.pileup import PileupTable

    rng = np.random.default_rng(seed)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod", "mod_type", "Nvalid_cov", "n_mod", "n_diff")}
    for i, L in enumerate((40000, 3000, 150000, 700)):
        seq = synth.random_sequence(rng, L, 0.5, 1e-4)
        name = f"contig_{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=8, with_counts=True)
        n = len(p["position"])
        cols["contig"].append(np.full(n, name, dtype=object))
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["mod_type"].append(np.array(synth.MOD_TYPES, dtype=object)[p["mod_type"]])
        for k in ("position", "fraction_mod", "Nvalid_cov", "n_mod", "n_diff"):
            cols[k].append(p[k])
    c = {k: np.concatenate(v) for k, v in cols.items()}
    # a contig of the pileup that is not in the assembly (allow_assembly_pileup_mismatch=True ignores it)
    c["contig"][:5] = "ghost"
    t = PileupTable(c["contig"], c["position"], c["strand"], c["fraction_mod"], c["mod_type"], c["Nvalid_cov"],
                    {"n_mod": c["n_mod"], "n_diff": c["n_diff"]})
    return contigs, t


@pytest.mark.parametrize("weighted", [False, True])
def test_methylation_pattern_matches_spec(weighted):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200.pattern import COLUMNS, MethylationOutput, methylation_pattern

    contigs, t = _data()
    motifs = ["GATC_a_1", "GATC_m_3", "CCWGG_m_1", "GRNGAAGY_a_5", "A_a_0", "CCGG_21839_0", "ACGTACGTAC_a_0"]
    df = methylation_pattern(t, contigs, motifs, output_type=MethylationOutput.WeightedMean if weighted else MethylationOutput.Median)
    assert list(df.columns) == COLUMNS  # nanomotif/main.py:157-161
    assert df["mod_position"].dtype == np.int8 and df["n_motif_obs"].dtype == np.int32
    want = O.methylation_pattern(contigs, t.contig, t.position, t.strand, t.mod_type, t.extra["n_mod"], t.Nvalid_cov,
                                 t.extra["n_diff"], motifs, weighted_mean=weighted)
    got = {(r.contig, r.motif, r.mod_type): r for r in df.itertuples()}
    assert len(got) == len(want) and len(want) > 10
    for name, motif, mt, mp, value, cov, n in want:
        r = got[(name, motif, mt)]
        assert r.mod_position == mp and r.n_motif_obs == n
        assert r.mean_read_cov == pytest.approx(cov, rel=1e-13)
        assert r.methylation_value == pytest.approx(value, rel=1e-12, abs=1e-15)
    # the reference's own test only pins columns and the (motifs x contigs) row count (tests/binnary/test_utils.py:56-59)
    two = methylation_pattern(t, contigs, ["GATC_m_3", "GATC_a_1"])
    assert list(two.columns) == COLUMNS and two.shape[1] == 7


def test_large_segment_median_uses_radix_select():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200.pattern import methylation_pattern

    contigs, t = _data(seed=10)
    df = methylation_pattern(t, contigs, ["A_a_0", "T_a_0"])  # thousands of observations per contig
    want = O.methylation_pattern(contigs, t.contig, t.position, t.strand, t.mod_type, t.extra["n_mod"], t.Nvalid_cov,
                                 t.extra["n_diff"], ["A_a_0", "T_a_0"])
    got = {(r.contig, r.motif): r for r in df.itertuples()}
    assert max(w[6] for w in want) > 1000
    for name, motif, mt, mp, value, cov, n in want:
        r = got[(name, motif)]
        assert r.n_motif_obs == n and r.methylation_value == pytest.approx(value, rel=1e-12, abs=1e-15)


def test_tile_driven_equals_row_driven_and_file_input(tmp_path):
    """Two independent implementations of the join (tile-driven nmb_pattern_scan behind methylation_pattern, row-driven
    nmb_pattern_stats / nmb_pattern_median) agree cell by cell; a bedMethyl FILE parsed on the device gives the same table."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200.pattern import methylation_pattern, methylation_pattern_rows

    contigs, t = _data(seed=12)
    specs = ["GATC_a_1", "CCWGG_m_1", "A_a_0", "GRNGAAGY_a_5", "C" + "N" * 33 + "AT_a_34", "T_m_0"]
    df = methylation_pattern(t, contigs, specs)
    names = list(contigs)
    for spec in specs:
        stats, med = methylation_pattern_rows(t, contigs, spec)
        sub = df[(df["motif"] + "_" + df["mod_type"] + "_" + df["mod_position"].astype(str)) == spec]
        assert len(sub) == int((stats[:, 0] > 0).sum())
        for r in sub.itertuples():
            c = names.index(r.contig)
            assert r.n_motif_obs == stats[c, 0] and r.mean_read_cov == stats[c, 2] / stats[c, 0]
            assert r.methylation_value == med[c]  # both take the exact median of the same multiset
    # the same pileup as a bedMethyl file
    keep = t.contig != "ghost"
    path = tmp_path / "p.bed"
    with open(path, "w") as f:
        for i in np.flatnonzero(keep):
            pos, cov = int(t.position[i]), int(t.Nvalid_cov[i])
            row = [t.contig[i], pos, pos + 1, t.mod_type[i], cov, t.strand[i], pos, pos + 1, "255,0,0", cov,
                   f"{t.fraction_mod[i] * 100:.2f}", int(t.extra["n_mod"][i]), 0, 0, 0, 0, int(t.extra["n_diff"][i]), 0]
            f.write("\t".join(str(x) for x in row) + "\n")
    fa = tmp_path / "a.fasta"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in contigs.items()))
    out = tmp_path / "table.tsv"
    df2 = methylation_pattern(str(path), str(fa), specs, output=str(out))
    assert df2.equals(df) and out.read_text().startswith("contig\tmotif\tmod_type")
