"""Drop-in boundary from REFERENCE-SHAPED host tables (what find_motifs_bin.py:399-427 hands to workers): Arrow / pandas /
numpy columns with string contig, strand and mod_type -> device rows (nmb_lookup_strings) -> counts, against the oracle."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O  # the checker, never the thing under test

MOD_TYPES = ("a", "m", "21839")
MOTIFS = {"a": [("GATC", 1), ("A", 0), ("G[AG].GAAG[CT]", 5), ("GCAC......GTT", 2)],
          "m": [("CC[AT]GG", 1), ("C", 0), ("GC.GC", 1)], "21839": [("CCGG", 0), ("C..G", 0)]}


@pytest.fixture(scope="module")
def nmb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb

    return nmb


@pytest.fixture(scope="module")
def data():
    """Two bins of three / two contigs with odd names, a pileup of three mod types in modkit order."""
    from nanomotif_b200 import synth

    rng = np.random.default_rng(21)
    bins = {"bin_A": {}, "bin_B": {}}
    cols = {k: [] for k in ("contig", "position", "strand", "mod_type", "fraction_mod", "Nvalid_cov")}
    for b, names in (("bin_A", ("contig_1", "contig_10", "k141_7 flag")), ("bin_B", ("c", "contig_100"))):
        for name in names:
            seq = synth.random_sequence(rng, int(rng.integers(3000, 90000)), 0.5, 2e-5)
            bins[b][name] = seq.tobytes().decode()
            p = synth.synth_pileup(seq, rng, depth=20)
            n = len(p["position"])
            cols["contig"].append(np.full(n, name, dtype=object))
            cols["position"].append(p["position"])
            cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
            cols["mod_type"].append(np.array(MOD_TYPES, dtype=object)[p["mod_type"]])
            cols["fraction_mod"].append(p["fraction_mod"])
            cols["Nvalid_cov"].append(p["Nvalid_cov"])
    return bins, {k: np.concatenate(v) for k, v in cols.items()}


def _oracle(bins, pile, bin_name, mt, motif, pos):
    sel = pile["mod_type"] == mt
    return O.motif_model_bin(pile["contig"][sel], pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel],
                             bins[bin_name], motif, pos, fast=True)


def _check(nmb, scorer, bins, pile):
    reqs, want = [], []
    for b in bins:
        for mt in MOD_TYPES:
            ms = [nmb.Motif(m, p) for m, p in MOTIFS[mt]]
            reqs.append((scorer.context(b, mt), ms))
            want.append([_oracle(bins, pile, b, mt, m, p) for m, p in MOTIFS[mt]])
    got = scorer.score_batch(reqs)
    for g, w, (ctx, ms) in zip(got, want, reqs):
        assert g.tolist() == [list(x) for x in w], (ctx.bin_name, ctx.mod_type)


def _arrow(pile, kind):
    import pyarrow as pa

    def strings(a):
        if kind == "string":
            return pa.array(a, type=pa.string())
        if kind == "large_string":
            return pa.array(a, type=pa.large_string())
        if kind == "dictionary":
            return pa.array(a, type=pa.string()).dictionary_encode()
        return pa.chunked_array([pa.array(a[:1000], type=pa.large_string()), pa.array(a[1000:], type=pa.large_string())])

    return pa.table({"contig": strings(pile["contig"]), "position": pa.array(pile["position"]),
                     "mod_type": strings(pile["mod_type"]), "strand": strings(pile["strand"]),
                     "fraction_mod": pa.array(pile["fraction_mod"]), "Nvalid_cov": pa.array(pile["Nvalid_cov"])})


@pytest.mark.parametrize("kind", ["string", "large_string", "dictionary", "chunked"])
def test_arrow_table_counts_equal_oracle(nmb, data, kind):
    bins, pile = data
    scorer = nmb.MultiBinScorer(_arrow(pile, kind), bins, MOD_TYPES, 0.3, 0.7)
    _check(nmb, scorer, bins, pile)


def test_sliced_arrow_table(nmb, data):
    """A zero-copy slice keeps the parent's buffers: offsets do not start at 0."""
    bins, pile = data
    n = len(pile["position"])
    lo, hi = n // 3, n - 17
    sub = {k: v[lo:hi] for k, v in pile.items()}
    scorer = nmb.MultiBinScorer(_arrow(pile, "large_string").slice(lo, hi - lo), bins, MOD_TYPES, 0.3, 0.7)
    _check(nmb, scorer, bins, sub)


def test_numpy_and_pandas_frames(nmb, data):
    import pandas as pd

    bins, pile = data
    _check(nmb, nmb.MultiBinScorer(pile, bins, MOD_TYPES, 0.3, 0.7), bins, pile)
    _check(nmb, nmb.MultiBinScorer(pd.DataFrame(pile), bins, MOD_TYPES, 0.3, 0.7), bins, pile)
    fixed = dict(pile, contig=pile["contig"].astype(str), strand=pile["strand"].astype(str))  # '<U..' columns
    _check(nmb, nmb.MultiBinScorer(fixed, bins, MOD_TYPES, 0.3, 0.7), bins, pile)


def test_partitioned_pileup(nmb, data):
    """The reference's partitioned_pileup: {(bin, mod_type): frame} (find_motifs_bin.py:416)."""
    bins, pile = data
    parts = {}
    for b, cs in bins.items():
        in_bin = np.isin(pile["contig"], list(cs))
        for mt in MOD_TYPES:
            sel = in_bin & (pile["mod_type"] == mt)
            parts[(b, mt)] = _arrow({k: v[sel] for k, v in pile.items()}, "large_string")
    _check(nmb, nmb.MultiBinScorer(parts, bins, MOD_TYPES, 0.3, 0.7), bins, pile)


def test_rows_round_trip_and_unknowns(nmb, data):
    """rows_from_table restores every column; unknown contigs / mod types / strands are flagged, and dropped from the
    counts exactly like the reference's strand == '+' / '-' filters drop a '.' row."""
    from nanomotif_b200.dataload import rows_from_table

    bins, pile = data
    names = [n for cs in bins.values() for n in cs]
    odd = {k: v.copy() for k, v in pile.items()}
    odd["contig"][5] = "not_in_assembly"
    odd["mod_type"][6] = "h"
    dot = np.flatnonzero(pile["fraction_mod"] >= 0.7)[:50]
    odd["strand"][dot] = "."
    rows = rows_from_table(_arrow(odd, "large_string"), names, MOD_TYPES)
    t = rows.to_table()
    np.testing.assert_array_equal(t.position, odd["position"])
    np.testing.assert_array_equal(t.fraction_mod, odd["fraction_mod"])
    np.testing.assert_array_equal(t.Nvalid_cov, odd["Nvalid_cov"])
    assert t.strand.tolist() == odd["strand"].tolist()
    want_contig = odd["contig"].copy()
    want_contig[5] = "?"
    assert t.contig.tolist() == want_contig.tolist()
    want_mt = odd["mod_type"].copy()
    want_mt[6] = "?"
    assert t.mod_type.tolist() == want_mt.tolist()
    keep = (odd["strand"] != ".") & (odd["contig"] != "not_in_assembly") & (odd["mod_type"] != "h")
    kept = {k: v[keep] for k, v in odd.items()}
    _check(nmb, nmb.MultiBinScorer(_arrow(odd, "string"), bins, MOD_TYPES, 0.3, 0.7), bins, kept)
    # the single-bin scorer agrees (host path: strand_codes / compact_rows)
    sel = (odd["mod_type"] == "a") & np.isin(odd["contig"], list(bins["bin_A"]))
    one = nmb.BinScorer({k: v[sel] for k, v in odd.items()}, bins["bin_A"], 0.3, 0.7)
    want = [list(_oracle(bins, kept, "bin_A", "a", m, p)) for m, p in MOTIFS["a"]]
    assert one.score([nmb.Motif(m, p) for m, p in MOTIFS["a"]]).tolist() == want


def test_duplicate_rows_are_reported(nmb, data):
    bins, pile = data
    dup = {k: np.concatenate([v, v[:100]]) for k, v in pile.items()}
    with pytest.warns(RuntimeWarning, match="repeat a"):
        nmb.MultiBinScorer(dup, bins, MOD_TYPES, 0.3, 0.7)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        nmb.MultiBinScorer(pile, bins, MOD_TYPES, 0.3, 0.7)


def test_drop_in_cache_sees_mutated_inputs(nmb, data):
    """The reference functions are pure; the cached device state behind the drop-in names must not go stale when the
    caller mutates the SAME objects between calls (contigs dict edited, pileup filtered in place)."""
    bins, pile = data
    contigs = dict(bins["bin_A"])
    sel = (pile["mod_type"] == "a") & np.isin(pile["contig"], list(contigs))
    p = {k: v[sel] for k, v in pile.items()}
    m = nmb.Motif("GATC", 1)

    def got():
        mdl = nmb.motif_model_bin(p, contigs, m, nmb.BetaBernoulliModel(), 0.3, 0.7)
        return mdl._alpha - 5, mdl._beta - 5

    def want():
        return tuple(O.motif_model_bin(p["contig"], p["position"], p["strand"], p["fraction_mod"], contigs, "GATC", 1, fast=True))

    first = got()
    assert first == want() and got() == first  # second call: cache hit
    del contigs[next(iter(contigs))]           # same dict object, one contig fewer
    assert got() == want() != first
    keep = p["position"] % 2 == 0              # same dict-of-arrays pileup, filtered in place
    for k in p:
        p[k] = p[k][keep]
    assert got() == want()
    nmb.clear_caches()
    assert got() == want()


def test_staged_copies_and_narrowing(nmb):
    """The pinned stager behind every large host -> device copy: plain, gathered from many pieces, and narrowing
    int64 -> int32 with an overflow flag; odd sizes around the 4 MB slot size."""
    import torch

    from nanomotif_b200 import device as D

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    for n in (D.STAGE_MIN_BYTES // 8 + 5, 3 * D.STAGE_SLOT_BYTES // 8 + 17, 5_000_001):  # below STAGE_MIN_BYTES: plain copy
        a = rng.integers(0, 2**31 - 1, size=n, dtype=np.int64)
        assert np.array_equal(D._to_device(a, dev).cpu().numpy(), a)                      # nmb_stager_copy
        narrow = D._to_device_narrow(a, dev)                                               # nmb_stager_copy_narrow
        assert narrow is not None and narrow.dtype == torch.int32
        assert np.array_equal(narrow.cpu().numpy(), a.astype(np.int32))
        a[n // 2] = 2**31  # does not fit
        assert D._to_device_narrow(a, dev) is None
        a[n // 2] = -5     # negative values fit
        assert np.array_equal(D._to_device_narrow(a, dev).cpu().numpy(), a.astype(np.int32))
    u8 = rng.integers(0, 256, size=9_000_001, dtype=np.uint8)
    assert np.array_equal(D._to_device(u8, dev).cpu().numpy(), u8)
    seqs = {f"c{i}": "".join(rng.choice(list("ACGTN"), size=int(rng.integers(1, 900_000)))) for i in range(23)}
    asm = D.DeviceAssembly.from_sequences(seqs)                                            # nmb_stager_gather
    assert asm.total_bp == sum(len(s) for s in seqs.values()) >= D.STAGE_MIN_BYTES
    old = D.STAGE_MIN_BYTES
    try:
        D.STAGE_MIN_BYTES = 1 << 60  # the plain path: "".join + one cudaMemcpy
        ref = D.DeviceAssembly.from_sequences(seqs)
    finally:
        D.STAGE_MIN_BYTES = old
    assert torch.equal(asm.seq_records, ref.seq_records) and torch.equal(asm.nonacgt, ref.nonacgt)
