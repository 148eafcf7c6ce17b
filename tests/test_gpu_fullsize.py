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
