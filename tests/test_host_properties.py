"""Property tests (hypothesis) of the host logic that decides WHAT the kernels are asked to compute: motif records,
the integer images of the float thresholds, compact pileup rows, the shard plan.  No device."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from nanomotif_b200 import _lib, sharding
from nanomotif_b200 import motif as M
from nanomotif_b200.device import compact_rows, percent_keys, plan_layout, threshold_keys

# the same examples on every run and no timing-dependent health checks: a CI box under load must not turn these red
STABLE = dict(deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))

TOKENS = ["A", "C", "G", "T", ".", ".", "[AC]", "[AG]", "[AT]", "[CG]", "[CT]", "[GT]", "[ACG]", "[ACT]", "[AGT]", "[CGT]"]
motif_st = st.builds(lambda toks, lead, trail, frac: ("." * lead + "".join(toks) + "." * trail, frac),
                     st.lists(st.sampled_from(TOKENS), min_size=1, max_size=20), st.integers(0, 5), st.integers(0, 5),
                     st.floats(0, 0.999))


def _with_mod_pos(spec):
    s, frac = spec
    toks = M.tokenize(s)
    return M.Motif(s, int(frac * len(toks)))


@settings(max_examples=300, **STABLE)
@given(st.lists(motif_st, min_size=1, max_size=12))
def test_batched_motif_packing_equals_the_one_by_one_path(specs):
    """pack_motifs packs bracket-free motifs with one table lookup for the whole batch; the records (and the refusals)
    must be those of the motif-by-motif path, and the masks must be the reference's one_hot rows of the STRIPPED motif."""
    motifs = [_with_mod_pos(s) for s in specs]

    def one(m):
        out = np.zeros(1, dtype=_lib.MOTIF_DTYPE)
        M._pack_one(out, 0, m, True, None)
        return out

    want, errors = [], []
    for m in motifs:
        try:
            want.append(one(m))
        except ValueError as exc:
            errors.append(str(exc))
    if errors:
        with pytest.raises(ValueError):
            M.pack_motifs(motifs)
        return
    got = M.pack_motifs(motifs)  # one native call for the batch (nmb_pack_motifs)
    assert got.tobytes() == np.concatenate(want).tobytes()
    assert M._pack_motifs_numpy(motifs).tobytes() == got.tobytes()  # the numpy implementation it replaced
    assert M.pack_motifs(motifs, strip=False, mod_pos_override=0).tobytes() == M._pack_motifs_numpy(motifs, False, 0).tobytes()
    for m, rec in zip(motifs, got):
        s = m.new_stripped_motif()
        oh = s.one_hot()
        assert rec["len"] == len(oh) and rec["mod_pos"] == s.mod_position
        assert rec["allowed"][:len(oh)].tolist() == [int(r[0] + 2 * r[1] + 4 * r[2] + 8 * r[3]) for r in oh]
        assert not rec["allowed"][len(oh):].any()


@settings(max_examples=300, **STABLE)
@given(st.lists(st.tuples(st.text(alphabet="ACGT.[]N", min_size=0, max_size=70), st.integers(-2, 70)), min_size=1, max_size=6),
       st.booleans(), st.sampled_from([None, 0, 3]))
def test_native_packer_refuses_exactly_what_the_python_path_refuses(specs, strip, override):
    """Arbitrary strings over the motif alphabet plus junk (unbalanced and empty brackets, N, too long, nothing but
    wildcards, mod positions outside the motif): the native batch packer and the numpy / one-by-one path either return
    the same records or raise the same kind of error (the same message for a single motif)."""
    motifs = [M.Motif(s, p) for s, p in specs]
    results = []
    for f in (M.pack_motifs, M._pack_motifs_numpy):
        try:
            results.append(("ok", f(motifs, strip, override).tobytes()))
        except (ValueError, TypeError) as exc:
            results.append((type(exc).__name__, str(exc)))
    if results[0][0] == "ok" or len(motifs) == 1:
        assert results[0] == results[1]
    else:
        assert results[0][0] == results[1][0]


@settings(max_examples=200, **STABLE)
@given(st.floats(-0.2, 1.2), st.floats(-0.2, 1.2))
def test_threshold_keys_are_the_float_tests_on_the_percent_grid(low, high):
    """fraction_mod = fl(fl(k/100)/100) for modkit's two-decimal percentages; `>= high` / `<= low` in float64
    (find_motifs_bin.py:1308-1309) must equal the integer tests the compact rows use, for EVERY grid value."""
    k_low, k_high = threshold_keys(low, high)
    k = np.arange(10001)
    frac = (k.astype(np.float64) / 100.0) / 100.0
    assert np.array_equal(frac >= high, k >= k_high) and np.array_equal(frac <= low, k <= k_low)
    assert percent_keys(frac).tolist() == k.tolist()


@settings(max_examples=100, **STABLE)
@given(st.integers(1, 400), st.integers(0, 2**32 - 1), st.integers(1, 6))
def test_compact_rows_keep_every_usable_row_grouped_by_contig(n, seed, n_contigs):
    rng = np.random.default_rng(seed)
    cid = rng.integers(-1, n_contigs, n)
    pos = rng.integers(0, 100000, n)
    strand = rng.integers(0, 3, n).astype(np.uint8)  # 2 = '.', takes no part
    key = rng.integers(0, 10001, n)
    frac = (key.astype(np.float64) / 100.0) / 100.0
    mt = rng.integers(0, 3, n).astype(np.uint8)
    rows = compact_rows(cid, pos, strand, frac, mt, n_contigs)
    keep = (cid >= 0) & (strand <= 1)
    assert len(rows["position"]) == int(keep.sum()) == int(rows["contig_row_off"][-1])
    assert rows["contig_row_off"].tolist() == np.concatenate([[0], np.cumsum(np.bincount(cid[keep], minlength=n_contigs))]).tolist()
    want = sorted(zip(cid[keep].tolist(), pos[keep].tolist(), strand[keep].tolist(), mt[keep].tolist(), key[keep].tolist()))
    got = []
    for c in range(n_contigs):
        a, b = rows["contig_row_off"][c], rows["contig_row_off"][c + 1]
        got += [(c, int(p), int(f & 1), int(f >> 1), int(q)) for p, f, q in
                zip(rows["position"][a:b], rows["flags"][a:b], rows["percent_x100"][a:b])]
    assert sorted(got) == want
    assert compact_rows(cid, pos, strand, frac + 1e-7, mt, n_contigs) is None  # off the grid: float64 rows instead


@settings(max_examples=200, **STABLE)
@given(st.lists(st.lists(st.integers(1, 300000), min_size=1, max_size=6), min_size=1, max_size=10), st.integers(1, 8))
def test_shard_plan_covers_every_base_pair_exactly_once(bin_lengths, world):
    bins = {f"b{i}": {f"b{i}_c{j}": n for j, n in enumerate(lens)} for i, lens in enumerate(bin_lengths)}
    owner, per_rank, split, bin_ranks = sharding.ShardedMultiBinScorer.plan(bins, world)
    assert len(per_rank) == world
    covered = {}
    for r, local in enumerate(per_rank):
        for b, cs in local.items():
            assert r in bin_ranks[b]
            for name, v in cs.items():
                if isinstance(v, sharding.ContigPiece):
                    assert v.name in bins[b] and v.length == bins[b][v.name] and 0 <= v.a < v.b <= v.length
                    assert v.lo == max(0, v.a - sharding.HALO_BP) and v.hi == min(v.length, v.b + sharding.HALO_BP)
                    covered.setdefault(v.name, []).append((v.a, v.b))
                else:
                    assert name in bins[b]
                    covered.setdefault(name, []).append((0, v))
    flat = {n: L for cs in bins.values() for n, L in cs.items()}
    assert set(covered) == set(flat)
    for name, ranges in covered.items():
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == flat[name]
        assert all(x[1] == y[0] for x, y in zip(ranges, ranges[1:]))  # no gap, no overlap
    for i, (name, L) in enumerate(flat.items()):
        assert (owner[i] >= 0) == (len(covered[name]) == 1)
    assert split == {b for b, r in bin_ranks.items() if len(r) > 1}
    total = sum(flat.values())
    if world > 1 and total > 8 * 1024 * world:  # no rank is asked to hold much more than its share + one unit
        loads = [sum((v.b - v.a) if isinstance(v, sharding.ContigPiece) else v for cs in local.values() for v in cs.values())
                 for local in per_rank]
        biggest_unit = max(max((p[1] - p[0]) for p in ranges) for ranges in covered.values())
        whole_bins = [sum(cs.values()) for b, cs in bins.items() if b not in split]
        assert max(loads) <= total / world + max([biggest_unit] + whole_bins)


@settings(max_examples=100, **STABLE)
@given(st.lists(st.integers(0, 200000), min_size=0, max_size=40))
def test_layout_keeps_contigs_apart_and_chunk_aligned(lengths):
    starts, n_tiles = plan_layout(lengths)
    lens = np.asarray(lengths, dtype=np.int64)
    assert (starts % _lib.CHUNK_BP == 0).all()
    if len(lens) > 1:
        assert ((starts[1:] - (starts[:-1] + lens[:-1])) >= _lib.MIN_GAP_BP).all()
    if len(lens):
        assert n_tiles * _lib.TILE_BP >= starts[-1] + lens[-1] + _lib.MIN_GAP_BP
