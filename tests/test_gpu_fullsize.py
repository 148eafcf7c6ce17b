"""Size-independent properties of the scan kernel at BASELINE.json's full size (cfg 3: 1.5 Gbp, 17 k contigs,
300 bins) where the CPU oracle would take hours: checksums against the class planes, linearity over allowed
sets, per-contig / per-bin / whole-assembly consistency, and agreement of the two kernels that share the
matcher (K2 counts vs K3 match planes).  All comparisons are exact integers."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200 import synth

    dev = torch.device("cuda", 0)
    asm, pile = synth.device_workload(dev, 1_500_000_000, 17000, seed=11)
    return dict(torch=torch, dev=dev, asm=asm, pile=pile)


def _popcount(torch, t):
    """Number of set bits of an int32 tensor (test helper, torch ops)."""
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=t.device)
    return int(lut[t.contiguous().view(torch.uint8).long()].sum().item())


def _scan(W, motifs, group_mode=0, contig_group=None, n_groups=1, mpi=None):
    import nanomotif_b200 as nmb
    from nanomotif_b200.device import MotifPrograms, make_jobs, scan_count

    asm = W["asm"]
    progs = MotifPrograms([nmb.Motif(s, p) for s, p in motifs], W["dev"])
    jobs = make_jobs(1)
    jobs["motif_count"], jobs["tile_count"], jobs["contig_end"] = len(motifs), asm.n_tiles, asm.n_contigs
    jobs["group_mode"], jobs["n_groups"] = group_mode, n_groups
    out = scan_count(asm, W["pile"], progs, jobs, len(motifs) * n_groups, motifs_per_item=mpi, contig_group=contig_group)
    return out.view(len(motifs), n_groups, 4).cpu().numpy()


def test_single_base_counts_equal_plane_popcounts(W):
    from nanomotif_b200 import _lib

    torch, asm = W["torch"], W["asm"]
    cls = W["pile"].class_records.view(asm.n_tiles, 4, _lib.TILE_WORDS)
    got = _scan(W, [("A", 0)])[0, 0]
    want = [_popcount(torch, cls[:, k]) for k in range(4)]  # every A ('+') / T ('-') site is an occurrence
    assert got.tolist() == want and min(want) > 5e7
    assert _scan(W, [("G", 0)])[0, 0].tolist() == [0, 0, 0, 0]  # the planes hold no G/C sites


def test_linearity_over_allowed_sets(W):
    c = _scan(W, [("GA.C", 1), ("GAAC", 1), ("GATC", 1), ("GAGC", 1), ("GACC", 1), ("G[AT]TC", 1), ("GTTC", 1),
                  ("GA[CGT]C", 1), ("CA......TG", 1), ("CA...[AC]..TG", 1), ("CA...[GT]..TG", 1)], mpi=4)[:, 0]
    np.testing.assert_array_equal(c[0], c[1] + c[2] + c[3] + c[4])      # wildcard = sum over the four bases
    np.testing.assert_array_equal(c[5][[0, 1]], (c[2] + c[6])[[0, 1]])  # '+' strand: [AT] at a non-modified position
    np.testing.assert_array_equal(c[7], c[2] + c[3] + c[4])              # three-letter class
    np.testing.assert_array_equal(c[8], c[9] + c[10])                    # complementary classes inside a gap
    assert c[2].min() > 1e5


def test_group_modes_are_consistent(W):
    torch, asm = W["torch"], W["asm"]
    motifs = [("GATC", 1), ("CC[AT]GG", 2), ("A", 0), ("GCAC......GTT", 2)]
    whole = _scan(W, motifs)[:, 0]
    rng = np.random.default_rng(0)
    bins = rng.integers(-1, 300, size=asm.n_contigs).astype(np.int32)  # -1 = contig left out
    by_bin = _scan(W, motifs, 2, torch.from_numpy(bins).to(W["dev"]), 300)
    by_contig = _scan(W, motifs[:2], 1, None, asm.n_contigs)
    np.testing.assert_array_equal(by_contig.sum(axis=1), whole[:2])
    for b in (0, 7, 299):
        np.testing.assert_array_equal(by_contig[:, bins == b].sum(axis=1), by_bin[:2, b])
    np.testing.assert_array_equal(by_contig[:, bins >= 0].sum(axis=1), by_bin[:2].sum(axis=1))
    left_out = _scan(W, motifs, 2, torch.from_numpy(np.where(bins < 0, 0, -1).astype(np.int32)).to(W["dev"]), 1)[:, 0]
    np.testing.assert_array_equal(by_bin.sum(axis=1) + left_out, whole)


def test_counts_agree_with_match_planes(W):
    """K2 (fused scan+join+reduce) vs K3 (match plane) + an explicit AND/popcount with the class planes."""
    import nanomotif_b200 as nmb
    from nanomotif_b200 import _lib
    from nanomotif_b200._lib import check, lib, ptr
    from nanomotif_b200.device import MotifPrograms, _stream

    torch, asm = W["torch"], W["asm"]
    cls = W["pile"].class_records.view(asm.n_tiles, 4, _lib.TILE_WORDS)
    for s, p in (("GATC", 1), ("T[AG]A....C", 2)):
        progs = MotifPrograms([nmb.Motif(s, p)], W["dev"])
        view = asm.view()
        want = []
        for strand in (0, 1):
            plane = torch.empty(asm.n_words, dtype=torch.int32, device=W["dev"])
            check(lib.nmb_match_plane(C.byref(view), ptr(progs.programs), 0, strand, progs.max_len, 0, asm.n_tiles,
                                      ptr(plane), _stream()))
            # match planes are in natural word order, class planes lane-interleaved (NMB_WORD_SLOT)
            m = plane.view(asm.n_tiles, _lib.TILE_WORDS)[:, torch.from_numpy(_lib.SLOT_WORD).to(W["dev"])]
            want += [_popcount(torch, m & cls[:, 2 * strand]), _popcount(torch, m & cls[:, 2 * strand + 1])]
        assert _scan(W, [(s, p)])[0, 0].tolist() == want


def test_pattern_table_at_cfg4_size():
    """BASELINE.json configs[3]: 50 000 contigs, 1.5 Gbp, read-level pileup (0.75e9 rows).  The CPU oracle cannot run
    at this size; checked through properties: K5 n_motif_obs == K2 per-contig counts over planes built from the same
    valid rows (two different kernels, two different joins), column sums against torch reductions of the rows, and
    exact medians of a few contigs recomputed from their rows."""
    import time

    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from nanomotif_b200 import synth
    from nanomotif_b200.device import DevicePileup, MotifPrograms, make_jobs, scan_count
    from nanomotif_b200.pattern import PatternIndex, pattern_table

    dev = torch.device("cuda", 0)
    asm, rows = synth.device_pattern_workload(dev, 1_500_000_000, 50_000, seed=4)
    nc = asm.n_contigs
    n_rows = int(rows["position"].numel())
    assert n_rows > 7e8
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index = PatternIndex(asm, rows, 0, 3, 0.8)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    cov, diff = rows["Nvalid_cov"], rows["n_diff"]
    ok = (cov >= 3) & ((cov.double() / (cov + diff).double()) >= 0.8)
    assert index.n_valid_rows == int(ok.sum()) and 0.5 * n_rows < index.n_valid_rows < n_rows

    specs = [("GATC", 1), ("A", 0), ("CC[AT]GG", 2), ("G[AG].GAAG[CT]", 5), ("GCAC......GTT", 2), ("T", 0),
             ("[AG]AA[CT]", 1), ("TTAA", 3)]
    motifs = [nmb.Motif(s, p) for s, p in specs]
    t0 = time.perf_counter()
    stats, med = pattern_table(index, motifs, median=True)
    t_table = time.perf_counter() - t0
    print(f"cfg4: index {t_index * 1e3:.0f} ms, {len(motifs)} motifs x {nc} contigs (median) {t_table * 1e3:.0f} ms")

    # (1) K2 on class planes made of the same valid rows: n_mod('+') + n_mod('-') per contig == n_motif_obs
    frac = torch.where(ok, 1.0, 0.5).double()
    pile = DevicePileup.from_columns(asm, rows["contig_id"], rows["position"], rows["strand"], frac, 0.3, 0.7, None, 1)
    del frac
    jobs = make_jobs(1)
    jobs["motif_count"], jobs["tile_count"], jobs["contig_end"] = len(motifs), asm.n_tiles, nc
    jobs["group_mode"], jobs["n_groups"] = 1, nc
    k2 = scan_count(asm, pile, MotifPrograms(motifs, dev), jobs, len(motifs) * nc).view(len(motifs), nc, 4).cpu().numpy()
    np.testing.assert_array_equal(stats[:, :, 0], k2[:, :, 0] + k2[:, :, 2])
    assert stats[0, :, 0].sum() > 1e6
    # (2) motif A: every valid '+' row is an occurrence, and T on '-' through the reverse complement
    cid = rows["contig_id"].long()
    for col, k in (("n_mod", 1), ("Nvalid_cov", 2)):
        want = torch.zeros(nc, dtype=torch.int64, device=dev).index_add_(0, cid[ok], rows[col][ok]).cpu().numpy()
        np.testing.assert_array_equal(stats[1, :, k], want)
    # motif T: '+' occurrences at T have no '+' rows, its reverse complement A has no '-' rows
    assert stats[5].sum() == 0 and np.isnan(med[5]).all()
    # (3) exact medians of a few contigs from their rows
    fr = rows["n_mod"].double() / cov.double()
    for c in (0, 17, 4242, nc - 1):
        sel = ok & (rows["contig_id"] == c)
        assert med[1, c] == float(np.median(fr[sel].cpu().numpy()))


def test_cfg2_size_counts_equal_oracle():
    """BASELINE.json configs[1] at full size against the ORACLE (not another kernel): one 4.6 Mbp contig, a depth-100
    pileup of three mod types, 36 motifs (the planted ones, short dense ones, gapped and degenerate ones, seeded random
    ones) -- every (n_mod, n_nomod) pair equals oracle.restate.motif_model_bin."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from nanomotif_b200 import synth
    from oracle import restate as O

    rng = np.random.default_rng(2)
    seq = synth.random_sequence(rng, 4_600_000, 0.508, 1e-6)
    pile = synth.synth_pileup(seq, rng, depth=100)
    text = seq.tobytes().decode()
    contigs = {"contig_0": text}
    fixed = {"a": [("GATC", 1), ("A", 0), ("GCAC......GTT", 2), ("AAC......GTGC", 1), ("G[AG].GAAG[CT]", 5), ("CA", 1),
                   ("T.A", 2), ("A" + "." * 40 + "C", 0)],
             "m": [("CC[AT]GG", 1), ("C", 0), ("GC.GC", 1), ("[AG]C[CT]", 1)],
             "21839": [("CCGG", 0), ("C.G", 0), ("GCGC", 1)]}
    mrng = np.random.default_rng(99)
    names = np.array(synth.MOD_TYPES, dtype=object)
    table = {"contig": np.full(len(pile["position"]), "contig_0", dtype=object), "position": pile["position"],
             "strand": np.where(pile["strand"] == 0, "+", "-").astype(object), "mod_type": names[pile["mod_type"]],
             "fraction_mod": pile["fraction_mod"]}
    scorer = nmb.MultiBinScorer(table, {"bin": contigs}, synth.MOD_TYPES, 0.3, 0.7)
    n = 0
    for mt, name in enumerate(synth.MOD_TYPES):
        motifs = fixed[name] + synth.random_motifs(mrng, 7, synth.CANONICAL[name])
        got = scorer.context("bin", name).score([nmb.Motif(m, p) for m, p in motifs])
        sel = pile["mod_type"] == mt
        for (m, p), g in zip(motifs, got.tolist()):
            want = O.motif_model_bin(table["contig"][sel], table["position"][sel], table["strand"][sel],
                                     table["fraction_mod"][sel], contigs, m, p, fast=True)
            assert tuple(g) == tuple(want), (name, m, p)
            n += 1
    assert n >= 30
