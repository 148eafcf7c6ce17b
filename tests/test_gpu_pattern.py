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


@pytest.mark.parametrize("weighted", [False, True])
def test_binnary_matrix_on_device_from_k5_arrays(weighted):
    """SURVEY 8f rank 3 on the GPU: the feature matrix of detect_contamination / include_contigs built by K9
    (nmb_bin_means + nmb_bin_matrix) straight from the K5 device arrays -- bit-identical to the host restatement
    (tables.bin_feature_matrix) and equal to the oracle's pandas pipeline (main.py:192-205,
    binnary/data_processing.py:174-213,255-269) run on the methylation_pattern frame."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandas as pd

    from nanomotif_b200 import synth, tables
    from nanomotif_b200.device import DeviceAssembly, _to_device
    from nanomotif_b200.motif import Motif
    from nanomotif_b200.pattern import PatternIndex, pattern_table

    rng = np.random.default_rng(31)
    contigs, cols = {}, {k: [] for k in ("contig_id", "position", "strand", "Nvalid_cov", "n_mod", "n_diff")}
    for i in range(26):
        seq = synth.random_sequence(rng, int(rng.integers(1500, 30000)), 0.5, 1e-4)
        contigs[f"contig_{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=int(rng.integers(2, 12)), mod_types=("a",), with_counts=True)
        cols["contig_id"].append(np.full(len(p["position"]), i, dtype=np.int32))
        for k in ("position", "strand", "Nvalid_cov", "n_mod", "n_diff"):
            cols[k].append(p[k])
    c = {k: np.concatenate(v) for k, v in cols.items()}
    asm = DeviceAssembly.from_sequences(contigs)
    d = asm.device
    rows = {k: _to_device(v, d) for k, v in c.items()}
    index = PatternIndex(asm, rows, 0)
    specs = [("GATC", 1), ("A", 0), ("GRNGAAGY", 5), ("TTAA", 3), ("CAYNNNNRTG", 1), ("ACGTACGTACGT", 0), ("AC", 0)]
    motifs = [Motif(s, p).from_iupac() for s, p in specs]
    motif_mods = [f"{s}_a_{p}" for s, p in specs]
    st_h, val_h = pattern_table(index, motifs, median=not weighted)
    st_d, val_d = pattern_table(index, motifs, median=not weighted, on_device=True, batch=3)
    np.testing.assert_array_equal(st_d.cpu().numpy(), st_h)
    if not weighted:
        np.testing.assert_array_equal(val_d.cpu().numpy(), val_h)
    names = list(contigs)
    contig_bin = {n: f"bin{(i * 5) % 4}" for i, n in enumerate(names) if i % 7 != 3}  # some contigs stay unbinned
    if weighted:
        with np.errstate(divide="ignore", invalid="ignore"):
            val_h = st_h[:, :, 1] / st_h[:, :, 2]
    mi, ci = np.nonzero(st_h[:, :, 0] > 0)
    frame = pd.DataFrame({"contig": np.array(names, dtype=object)[ci], "motif": [specs[m][0] for m in mi], "mod_type": "a",
                          "mod_position": [specs[m][1] for m in mi], "methylation_value": val_h[mi, ci],
                          "mean_read_cov": st_h[mi, ci, 2] / st_h[mi, ci, 0], "n_motif_obs": st_h[mi, ci, 0].astype(np.int32)})
    for thr in (24.0, 0.0, 60.0):
        got_c, got_m, got_f = tables.bin_feature_matrix_device(st_d, val_d, names, motif_mods, contig_bin, thr)
        host_c, host_m, host_f = tables.bin_feature_matrix(st_h, val_h, names, motif_mods, contig_bin, thr)
        assert got_m.is_cuda and got_c.tolist() == host_c.tolist() and got_f.tolist() == host_f.tolist()
        np.testing.assert_array_equal(got_m.cpu().numpy(), host_m)  # same sums in the same order: bit-identical
        want_c, want_m, want_f = O.binnary_matrix(frame, contig_bin, thr)
        assert got_c.tolist() == want_c.tolist() and got_f.tolist() == want_f.tolist()
        np.testing.assert_allclose(got_m.cpu().numpy(), want_m, rtol=1e-13, atol=0)
    assert got_m.shape[0] >= 1
