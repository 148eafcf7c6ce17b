"""Parity of the CUDA scan / join / reduce path (K1-K3) against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O  # the checker, never the thing under test


@pytest.fixture(scope="module")
def nmb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb

    return nmb


def _strand_str(codes):
    return np.where(np.asarray(codes) == 0, "+", "-")


MOTIFS = [("GATC", 1), ("A", 0), ("CC[AT]GG", 1), ("G[AG].GAAG[CT]", 5), ("GCAC......GTT", 2), ("AAC......GTGC", 1),
          ("[ACG]A[CT]", 1), ("C", 0), ("TTAA", 3), ("A..............................T", 0),
          ("A........................................................C", 0), ("ACGT", 0), ("ACG[AT]", 2),
          ("C" + "." * 31 + "A", 32), ("C" + "." * 30 + "A", 31), ("G" + "." * 40 + "A" + "." * 10 + "C", 41),
          ("T" + "." * 60 + "A", 61), ("A" + "." * 60 + "G", 0), ("[CG]" + "." * 33 + "A..T", 34)]


def test_subseq_indices_reference_kat(nmb):
    # /root/reference/tests/test_fasta.py:95-109
    seq = "AATTAAATTAAGTAAAT"
    assert nmb.subseq_indices("AATT", seq).tolist() == [0, 5]
    assert nmb.subseq_indices("AA.T", seq).tolist() == [0, 4, 5, 9, 13]
    # SURVEY 8c golden (regex-literal semantics on non-ACGT letters)
    seq = "ACGTNACGTRACGTACNT"
    assert nmb.subseq_indices("ACGT", seq).tolist() == [0, 5, 10]
    assert nmb.subseq_indices("AC.T", seq).tolist() == [0, 5, 10, 14]
    assert nmb.subseq_indices("A[CG]GT", seq).tolist() == [0, 5, 10]


def test_methylated_motif_occourances_reference_kat(nmb):
    # /root/reference/tests/test_motif_find.py:14-39
    m = nmb.Motif("ACG", 0)
    seq = "TACGGACGCCACG"
    a, b = nmb.methylated_motif_occourances(m, seq, np.array([1, 5]), np.array([10]))
    assert a.tolist() == [1, 5] and b.tolist() == [10]
    a, b = nmb.methylated_motif_occourances(m, seq, np.array([]), np.array([1, 10]))
    assert a.tolist() == [] and b.tolist() == [1, 10]


@pytest.mark.parametrize("length,n_rate", [(1000, 0.0), (70000, 0.0), (200001, 2e-5), (65536 - 64, 0.0), (65536, 1e-4)])
def test_subseq_indices_random(nmb, length, n_rate):
    from nanomotif_b200 import synth

    rng = np.random.default_rng(length)
    seq = synth.random_sequence(rng, length, 0.5, n_rate).tobytes().decode()
    for motif, _ in MOTIFS + [(".A.", 0), ("..GATC...", 0), ("A.", 0)]:
        got = nmb.subseq_indices(motif, seq)
        want = O.subseq_indices(motif, seq)
        assert got.dtype == np.int64
        np.testing.assert_array_equal(got, want, err_msg=motif)


def test_motif_model_bin_counts(nmb):
    from nanomotif_b200 import synth

    rng = np.random.default_rng(11)
    lengths = [300, 5000, 65536 * 2 + 17, 256, 40000, 65472, 100000]
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate(lengths):
        seq = synth.random_sequence(rng, L, 0.45, 1e-4 if i % 2 else 0.0)
        name = f"contig_{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=20, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(_strand_str(p["strand"]))
        cols["fraction_mod"].append(p["fraction_mod"])
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    motifs = [nmb.Motif(s, p) for s, p in MOTIFS] + [nmb.Motif("....GATC....", 5)]
    models = nmb.motif_model_bin_many(pile, contigs, motifs, 0.3, 0.7)
    for m, mdl in zip(motifs, models):
        want = O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], contigs,
                                 m.string, m.mod_position, 0.3, 0.7, fast=True)
        assert (mdl._alpha - 5, mdl._beta - 5) == want, (m, mdl, want)
    # single-motif entry point mutates and returns the model it was given
    mdl = nmb.BetaBernoulliModel()
    out = nmb.motif_model_bin(pile, contigs, motifs[0], mdl, 0.3, 0.7)
    assert out is mdl and (mdl._alpha, mdl._beta) == (models[0]._alpha, models[0]._beta)
    # per-contig rows agree with the per-contig oracle
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    per = scorer.counts_by_strand(motifs[:4], per_contig=True).cpu().numpy()
    for mi, m in enumerate(motifs[:4]):
        for ci, (name, seq) in enumerate(contigs.items()):
            sel = pile["contig"] == name
            a, b, d = O.motif_model_contig(pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel], seq,
                                           m.string, m.mod_position, 0.3, 0.7, fast=True)
            assert per[mi, ci].tolist() == [len(d["index_meth_fwd"]), len(d["index_nonmeth_fwd"]),
                                            len(d["index_meth_rev"]), len(d["index_nonmeth_rev"])]


def test_motif_model_contig_positions(nmb):
    from nanomotif_b200 import synth

    rng = np.random.default_rng(5)
    seq = synth.random_sequence(rng, 30000, 0.5, 1e-4)
    p = synth.synth_pileup(seq, rng, depth=15, mod_types=("a",))
    pile = dict(position=p["position"], strand=_strand_str(p["strand"]), fraction_mod=p["fraction_mod"])
    contig = seq.tobytes().decode()
    for s, mp in MOTIFS[:6]:
        mdl, data = nmb.motif_model_contig(pile, contig, nmb.BetaBernoulliModel(), nmb.Motif(s, mp), 0.3, 0.7,
                                           save_motif_positions=True)
        a, b, want = O.motif_model_contig(pile["position"], pile["strand"], pile["fraction_mod"], contig, s, mp)
        assert (mdl._alpha - 5, mdl._beta - 5) == (a, b)
        for k in want:
            np.testing.assert_array_equal(data[k], want[k], err_msg=f"{s} {k}")


@pytest.mark.parametrize("low,high", [(0.3, 0.7), (0.25, 0.75), (0.295, 0.705), (0.5, 0.5001), (0.0, 1.0)])
def test_compact_rows_give_identical_class_planes(nmb, low, high):
    """7-byte rows + integer threshold keys == float64 rows + the reference's float tests, bit for bit."""
    import torch

    from nanomotif_b200 import synth
    from nanomotif_b200.device import DeviceAssembly, DevicePileup, compact_rows, threshold_keys

    rng = np.random.default_rng(8)
    seqs, cols = {}, {k: [] for k in ("cid", "position", "strand", "mod_type", "fraction_mod")}
    for i, L in enumerate((70000, 900, 30000)):
        seq = synth.random_sequence(rng, L, 0.5)
        seqs[f"c{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=int(rng.integers(3, 200)))
        cols["cid"].append(np.full(len(p["position"]), i if i != 1 else -1, dtype=np.int32))  # contig 1 unknown
        for k in ("position", "strand", "mod_type", "fraction_mod"):
            cols[k].append(p[k])
    c = {k: np.concatenate(v) for k, v in cols.items()}
    perm = rng.permutation(len(c["cid"]))  # not grouped by contig: compact_rows has to group
    c = {k: v[perm] for k, v in c.items()}
    asm = DeviceAssembly.from_sequences(seqs)
    a = DevicePileup.from_columns(asm, c["cid"], c["position"], c["strand"], c["fraction_mod"], low, high, c["mod_type"], 3)
    rows = compact_rows(c["cid"], c["position"], c["strand"], c["fraction_mod"], c["mod_type"], asm.n_contigs)
    assert rows is not None and rows["percent_x100"].dtype == np.uint16
    b = DevicePileup.from_compact(asm, low=low, high=high, n_modtypes=3, **rows)
    assert torch.equal(a.class_records, b.class_records) and int(a.class_records.ne(0).sum()) > 1000
    k_lo, k_hi = threshold_keys(low, high)
    grid = (np.arange(10001) / 100.0) / 100.0
    assert np.array_equal(grid >= high, np.arange(10001) >= k_hi) and np.array_equal(grid <= low, np.arange(10001) <= k_lo)
    # fractions off the two-decimal grid cannot be compacted: the float64 path stays in charge
    assert compact_rows(c["cid"], c["position"], c["strand"], c["fraction_mod"] + 1e-9, c["mod_type"], 3) is None


def test_contig_edges_poly_a(nmb):
    """Inter-contig padding is stored as code 0 (= 'A') and kept out by clipping the finished chains
    (scan.cuh LaneEdge): poly-A / poly-T contigs whose ends sit on, before and after chunk and tile borders,
    scored with motifs made of A / T only -- any occurrence leaking into the padding would be counted."""
    rng = np.random.default_rng(21)
    lengths = [1, 2, 31, 63, 64, 65, 450, 511, 512, 513, 540, 1023, 1024, 1025, 16384, 16383, 65536 - 513, 65536 - 1,
               65536, 65537, 3 * 512 + 7, 70000, 33, 5]
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate(lengths):
        seq = np.full(L, ord("A") if i % 3 else ord("T"), dtype=np.uint8)
        flip = rng.random(L) < 0.02  # a few other letters so that not everything matches
        seq[flip] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(flip.sum()))]
        name = f"edge_{i}"
        contigs[name] = seq.tobytes().decode()
        for strand in "+-":
            cols["contig"].append(np.full(L, name, dtype=object))
            cols["position"].append(np.arange(L, dtype=np.int64))
            cols["strand"].append(np.full(L, strand, dtype=object))
            cols["fraction_mod"].append(rng.choice([0.05, 0.5, 0.95], size=L))
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    specs = [("A", 0), ("T", 0), ("AA", 1), ("AAAA", 0), ("AAAA", 3), ("TTTT", 1), ("A" * 31, 15), ("T" * 32, 31),
             ("A......A", 0), ("A" + "." * 30 + "A", 31), ("T" + "." * 30 + "A", 0), ("A" + "." * 60 + "A", 0),
             ("A" + "." * 60 + "A", 61), ("T" + "." * 45 + "TT", 46), ("A" * 40, 39), ("[AT]" * 33, 32)]
    motifs = [nmb.Motif(s, p) for s, p in specs]
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    per = scorer.counts_by_strand(motifs, per_contig=True).cpu().numpy()
    for mi, m in enumerate(motifs):
        for ci, (name, seq) in enumerate(contigs.items()):
            sel = pile["contig"] == name
            a, b, d = O.motif_model_contig(pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel], seq,
                                           m.string, m.mod_position, 0.3, 0.7, fast=True)
            want = [len(d["index_meth_fwd"]), len(d["index_nonmeth_fwd"]), len(d["index_meth_rev"]), len(d["index_nonmeth_rev"])]
            assert per[mi, ci].tolist() == want, (m, name, len(seq))
    assert per.sum() > 100000
    # K3 shares the clip: positions on the same contigs
    for s, _ in specs[:12]:
        for name in ("edge_8", "edge_9", "edge_17", "edge_19"):
            np.testing.assert_array_equal(nmb.subseq_indices(s, contigs[name]), O.subseq_indices(s, contigs[name]), err_msg=s)


def test_streamed_host_blocks_equal_one_shot(nmb):
    """pipeline.score_host_blocks (copies overlapped with class-plane builds and scans) == DevicePileup.from_compact +
    one scan launch, for blocks per mod type, blocks in scrambled order and a block mixing two mod types."""
    import torch

    from nanomotif_b200 import synth
    from nanomotif_b200.device import DeviceAssembly, DevicePileup, MotifPrograms, compact_rows, make_jobs, scan_count
    from nanomotif_b200.motif import pack_motifs
    from nanomotif_b200.pipeline import HostBlock, blocks_by_modtype, score_host_blocks

    rng = np.random.default_rng(31)
    lens = [70000, 900, 140000, 5000]
    seqs = [synth.random_sequence(rng, L, 0.5, 1e-4 if i == 2 else 0.0) for i, L in enumerate(lens)]
    cols = {k: [] for k in ("cid", "position", "strand", "mod_type", "fraction_mod")}
    for i, seq in enumerate(seqs):
        p = synth.synth_pileup(seq, rng, depth=30)
        cols["cid"].append(np.full(len(p["position"]), i, dtype=np.int32))
        for k in ("position", "strand", "mod_type", "fraction_mod"):
            cols[k].append(p[k])
    c = {k: np.concatenate(v) for k, v in cols.items()}
    rows = compact_rows(c["cid"], c["position"], c["strand"], c["fraction_mod"], c["mod_type"], len(lens))
    ascii_u8 = np.concatenate(seqs)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    names = [f"c{i}" for i in range(len(lens))]
    work = [(s, p, mt) for mt, base in enumerate("ACC") for s, p in synth.random_motifs(np.random.default_rng(5 + mt), 40, base)]
    packed = pack_motifs([nmb.Motif(s, p) for s, p, _ in work])
    jobs = make_jobs(4)  # mod types 0, 1, 2 over the whole bin + a per-contig job of mod type 1
    for j, (mt, mode) in enumerate(((0, 0), (1, 0), (2, 0), (1, 1))):
        jobs[j]["motif_begin"], jobs[j]["motif_count"], jobs[j]["modtype"] = 40 * mt, 40, mt
        jobs[j]["contig_end"], jobs[j]["group_mode"] = len(lens), mode
        jobs[j]["n_groups"] = len(lens) if mode else 1
    jobs["out_base"] = np.concatenate([[0], np.cumsum(jobs["motif_count"] * jobs["n_groups"])[:-1]])
    n_out = int((jobs["motif_count"] * jobs["n_groups"]).sum())

    dev = torch.device("cuda", 0)
    asm = DeviceAssembly(names, lens, ascii_u8, off, dev)
    pile = DevicePileup.from_compact(asm, low=0.3, high=0.7, n_modtypes=3, **rows)
    j1 = jobs.copy()
    j1["tile_count"] = asm.n_tiles
    want = scan_count(asm, pile, MotifPrograms(packed, dev), j1, n_out).cpu()
    assert int(want.sum()) > 1000

    per_mt = blocks_by_modtype(rows["position"], rows["flags"], rows["percent_x100"], rows["contig_row_off"], 3)
    assert sum(len(b.position) for b in per_mt) == len(rows["position"])
    mixed = blocks_by_modtype(rows["position"], rows["flags"] & 1 | (np.minimum(rows["flags"] >> 1, 1) << 1),
                              rows["percent_x100"], rows["contig_row_off"], 2)
    # block "1" of `mixed` holds the rows of mod types 1 and 2: restore their flags
    sel = np.flatnonzero((rows["flags"] >> 1) >= 1)
    mixed[1] = HostBlock(mixed[1].position, rows["flags"][sel], mixed[1].percent_x100, mixed[1].contig_row_off, (1, 2))
    from nanomotif_b200.pipeline import blocks_by_position

    by_pos = blocks_by_position(rows["position"], rows["flags"], rows["percent_x100"], rows["contig_row_off"], lens, 3)
    assert len(by_pos) >= 3 and all(b.tiles is not None for b in by_pos)
    assert sum(len(b.position) for b in by_pos) == len(rows["position"])
    halves = blocks_by_position(rows["position"], rows["flags"], rows["percent_x100"], rows["contig_row_off"], lens, 3, (0.5, 0.5))
    for blocks in (per_mt, per_mt[::-1], [per_mt[0], mixed[1]], [mixed[1], per_mt[0]], by_pos, by_pos[::-1], halves):
        got = score_host_blocks(names, lens, ascii_u8, off, blocks, packed, jobs, n_out, low=0.3, high=0.7, n_modtypes=3,
                                device=dev)
        assert torch.equal(got, want)


@pytest.mark.parametrize("seed", [1, 2])
def test_random_layouts_against_oracle(nmb, seed):
    """Differential fuzz of K2 (per-contig counts) and K5-style joins' matcher: hundreds of short contigs per tile (every
    warp straddles contigs), non-ACGT letters in the interior and right at contig ends, motifs of every length class
    (one- and two-word halos, gaps, bracket classes) -- the three matcher paths (plain, edge clip, non-ACGT) and their
    combinations inside one warp."""
    rng = np.random.default_rng(1000 + seed)
    n_contigs = 120
    lengths = np.concatenate([rng.integers(1, 80, 30), rng.integers(400, 700, 40), rng.integers(1000, 4000, 40),
                              rng.integers(20000, 70000, 10)])
    rng.shuffle(lengths)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate(lengths[:n_contigs]):
        L = int(L)
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L, p=[0.4, 0.1, 0.1, 0.4]).copy()
        kind = i % 4
        if kind == 1 and L > 3:      # N at the very ends
            seq[0] = seq[-1] = ord("N")
        elif kind == 2 and L > 50:   # IUPAC letters in the interior
            for s in rng.integers(0, L - 5, 3):
                seq[s:s + int(rng.integers(1, 5))] = ord(rng.choice(list("NRYK")))
        name = f"f{i}"
        contigs[name] = seq.tobytes().decode()
        for strand in "+-":
            keep = rng.random(L) < 0.7
            pos = np.flatnonzero(keep).astype(np.int64)
            cols["contig"].append(np.full(len(pos), name, dtype=object))
            cols["position"].append(pos)
            cols["strand"].append(np.full(len(pos), strand, dtype=object))
            cols["fraction_mod"].append(rng.choice([0.0, 0.3, 0.31, 0.69, 0.7, 1.0], size=len(pos)))
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    specs = [("A", 0), ("TA", 1), ("CA[AT]", 1), ("A.A", 2), ("T" + "." * 29 + "A", 30), ("A" + "." * 30 + "T", 0),
             ("AC" + "." * 40 + "[AG]T", 1), ("T" + "." * 60 + "A", 61), ("A" + "." * 55 + "CA", 0), ("[ACT]A[AGT]", 1),
             ("AAAAAAAAT", 3), ("TTTTTTTTTTTA", 11), ("ATATATATATATAT", 6), ("GA.C", 1), ("A....T....A", 5)]
    motifs = [nmb.Motif(s, p) for s, p in specs]
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    per = scorer.counts_by_strand(motifs, per_contig=True).cpu().numpy()
    total = 0
    for mi, m in enumerate(motifs):
        for ci, (name, seq) in enumerate(contigs.items()):
            sel = pile["contig"] == name
            a, b, d = O.motif_model_contig(pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel], seq,
                                           m.string, m.mod_position, 0.3, 0.7, fast=True)
            want = [len(d["index_meth_fwd"]), len(d["index_nonmeth_fwd"]), len(d["index_meth_rev"]), len(d["index_nonmeth_rev"])]
            assert per[mi, ci].tolist() == want, (m, name, len(seq))
            total += sum(want)
    assert total > 50000
    # whole-bin sums through the other group mode
    models = nmb.motif_model_bin_many(pile, contigs, motifs, 0.3, 0.7)
    for mi, mdl in enumerate(models):
        assert (mdl._alpha - 5, mdl._beta - 5) == (int(per[mi, :, [0, 2]].sum()), int(per[mi, :, [1, 3]].sum()))


def test_family_sharing_equals_general_path_and_oracle(nmb):
    """K2's family path (children of one search expansion share the parent's chains) against the general path and the
    oracle: search-shaped work lists (synth.frontier_worklist) plus hand-made families -- extra position left / right of
    the parent, same position with different bases, bracket classes, members beyond 31 positions (no family), short
    contigs and non-ACGT letters (those warps take the general path) -- for several motifs-per-item block sizes."""
    import torch

    from nanomotif_b200 import device as D, synth

    rng = np.random.default_rng(23)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate((150000, 700, 66000, 40, 9000, 131072 - 64)):
        seq = synth.random_sequence(rng, L, 0.45, 3e-5 if L > 1000 else 0.0)
        contigs[f"c{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=15, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), f"c{i}", dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(_strand_str(p["strand"]))
        cols["fraction_mod"].append(p["fraction_mod"])
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    lists = []
    for seed in (1, 2, 3):
        for kids in synth.frontier_worklist(seed, "A", rounds=24, width=4):
            lists.append(kids)
    lists += [
        [("GA", 1), ("CA", 1), ("TA", 1), ("AA", 1)],                       # parent = the modified base alone
        [("A.G", 0), ("A.C", 0), ("A..T", 0), ("T.A", 2)],                   # extras on both sides
        [("GA.C", 1), ("GA[CT]C", 1), ("GA[AG]C", 1)],                       # a class as the extra; 'GA.C' is the parent itself
        [("GATC", 1), ("GATG", 1), ("CATC", 1), ("GAT.C", 1)],               # mixed: runs break and restart
        [("A" + "." * 30 + "C", 0), ("A" + "." * 30 + "G", 0)],              # extra at offset 31: still a family
        [("A" + "." * 31 + "C", 0), ("A" + "." * 31 + "G", 0)],              # offset 32: general path
        [("C" + "." * 30 + "AG", 31), ("T" + "." * 30 + "AG", 31)],          # extra 31 to the left
        [("G" + "." * 28 + "A.C", 29), ("G" + "." * 28 + "A.T", 29)],        # parent spans 30 positions
        [("GATC", 1)], [("GATC", 1), ("GATC", 1)],                           # single; duplicates (no extra position)
    ]
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    flat = [nmb.Motif(m, p) for kids in lists for m, p in kids]
    want = np.array([O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], contigs,
                                       m.string, m.mod_position, fast=True) for m in flat])
    old, old_bal = D.FAMILIES, D.BALANCED
    try:
        for families, balanced in ((True, True), (False, True), (False, False)):  # families / dynamic / static item split
            D.FAMILIES, D.BALANCED = families, balanced
            for mpi in (None, 4, 3, 32):
                got = scorer.counts_by_strand(flat, motifs_per_item=mpi).cpu().numpy()
                np.testing.assert_array_equal(np.stack([got[:, 0] + got[:, 2], got[:, 1] + got[:, 3]], axis=1), want,
                                              err_msg=f"families={families} balanced={balanced} mpi={mpi}")
            per = scorer.counts_by_strand(flat[:40], per_contig=True).cpu().numpy()  # group mode 1 through the same path
            assert per.sum() == scorer.counts_by_strand(flat[:40]).cpu().numpy().sum()
    finally:
        D.FAMILIES, D.BALANCED = old, old_bal
    assert int(want.sum()) > 10000
