"""Host-side logic of nanomotif_b200 (no GPU): Motif type, motif compilation, posterior / scores, layout
planning, pileup adapters, synthetic generators -- checked against the golden vectors of the reference."""
import json
import os

import numpy as np
import pytest

import nanomotif_b200 as nmb
from nanomotif_b200 import _lib, model, motif as M, synth
from nanomotif_b200.device import choose_motifs_per_item, make_jobs, plan_layout
from nanomotif_b200.pileup import PileupTable, strand_codes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def test_motif_type_matches_reference(G):
    for c in G["motif"]:
        m = nmb.Motif(c["motif"], c["mod_pos"])
        st = m.new_stripped_motif()
        assert [st.string, st.mod_position] == c["stripped"]
        rc = st.reverse_compliment()
        assert [rc.string, rc.mod_position] == c["rc_of_stripped"]
        assert m.one_hot().tolist() == c["one_hot"]
        assert m.split() == c["split"] and m.length() == c["length"]
        assert m.iupac() == c["iupac"]
    for c in G["from_iupac"]:
        assert nmb.Motif(c["iupac"], 0).from_iupac().string == c["regex"]
    a, b = nmb.Motif("GATC", 1), nmb.Motif("GATC", 1)
    assert a == b and hash(a) == hash(b) and not (a == nmb.Motif("GATC", 2)) and repr(a) == "Motif('GATC', pos=1)"


def test_pack_motifs():
    rec = M.pack_motifs([nmb.Motif("....G[AG].GAAG[CT]....", 9), nmb.Motif("GATC", 1)])
    assert rec.dtype.itemsize == 64
    assert rec["len"].tolist() == [8, 4] and rec["mod_pos"].tolist() == [5, 1]
    assert rec["allowed"][0, :8].tolist() == [4, 5, 15, 4, 1, 1, 4, 10]  # bit0=A bit1=T bit2=G bit3=C
    assert rec["allowed"][1, :4].tolist() == [4, 1, 2, 8]
    with pytest.raises(ValueError):
        M.pack_motifs([nmb.Motif(".....", 2)])
    with pytest.raises(ValueError):
        M.pack_motifs([nmb.Motif("A" * 63, 0)])
    with pytest.raises(ValueError):
        M.pack_motifs([nmb.Motif("ANT", 0)])
    with pytest.raises(TypeError):
        M.pack_motifs(["GATC"])


def test_scores_match_reference(G):
    for c in G["scores"]:
        nxt, cur = nmb.BetaBernoulliModel(), nmb.BetaBernoulliModel()
        nxt.update(*c["next"])
        cur.update(*c["cur"])
        assert nmb.predictive_evaluation_score(nxt, cur) == pytest.approx(c["score"], rel=1e-12, abs=1e-15)
        assert model.priority(nxt, cur) == pytest.approx(c["priority"], rel=1e-12, abs=1e-15)
        assert nxt.mean() == pytest.approx(c["mean_next"], rel=1e-15)
        assert nxt.variance() == pytest.approx(c["variance_next"], rel=1e-13)
        assert nxt.get_raw_counts() == tuple(c["next"])
    nx = np.array([c["next"] for c in G["scores"]]) + 5
    cu = np.array([c["cur"] for c in G["scores"]]) + 5
    got = model.predictive_evaluation_scores(nx[:, 0], nx[:, 1], cu[:, 0], cu[:, 1])
    np.testing.assert_allclose(got, [c["score"] for c in G["scores"]], rtol=1e-12, atol=1e-15)


def test_model_pickle_roundtrip():
    import pickle

    m = nmb.BetaBernoulliModel()
    m.update(7, 9)
    m2 = pickle.loads(pickle.dumps(m))
    assert (m2._alpha, m2._beta, m2._alpha_prior, m2._beta_prior) == (12, 14, 5, 5)


def test_plan_layout():
    starts, n_tiles = plan_layout([300, 5000, 65536 * 2 + 17, 512, 1])
    assert (starts % _lib.CHUNK_BP == 0).all()
    lens = np.array([300, 5000, 65536 * 2 + 17, 512, 1])
    assert ((starts[1:] - (starts[:-1] + lens[:-1])) >= _lib.MIN_GAP_BP).all()
    assert n_tiles * _lib.TILE_BP >= starts[-1] + lens[-1] + _lib.MIN_GAP_BP
    assert plan_layout([])[1] == 1 and len(plan_layout([])[0]) == 0
    # the closed form equals the contig-by-contig recurrence (cursor rounded up to a chunk after every contig + gap)
    rng = np.random.default_rng(3)
    lens = np.concatenate([rng.integers(0, 2000, 300), rng.integers(1, 300000, 50), [0, 448, 449, 511, 512, 513]])
    rng.shuffle(lens)
    cur, want = 0, []
    for n in lens.tolist():
        want.append(cur)
        cur = -(-(cur + n + _lib.MIN_GAP_BP) // _lib.CHUNK_BP) * _lib.CHUNK_BP
    starts, n_tiles = plan_layout(lens)
    assert starts.tolist() == want and starts.dtype == np.int64 and n_tiles == max(1, -(-cur // _lib.TILE_BP))


def test_jobs_and_items():
    jobs = make_jobs(2)
    assert jobs.dtype.itemsize == 48
    jobs["motif_count"] = [1000, 3]
    jobs["tile_count"] = [71, 2]
    assert choose_motifs_per_item(jobs, 148) == 29
    jobs["motif_count"] = [1, 1]
    assert choose_motifs_per_item(jobs, 148) == 1


def test_pileup_adapters():
    d = dict(contig=np.array(["a", "b"]), position=[3, 4], strand=np.array(["+", "-"]), fraction_mod=[0.1, 0.9],
             mod_type=np.array(["a", "a"]), Nvalid_cov=[10, 20])
    t = PileupTable.from_frame(d)
    assert t.position.dtype == np.int64 and t.fraction_mod.dtype == np.float64 and len(t) == 2
    assert strand_codes(t.strand).tolist() == [0, 1]
    import pandas as pd

    t2 = PileupTable.from_frame(pd.DataFrame(d))
    assert t2.position.tolist() == [3, 4] and t2.contig.tolist() == ["a", "b"]
    assert t.take(np.array([False, True])).position.tolist() == [4]
    with pytest.raises(KeyError):
        PileupTable.from_frame({"position": [1]})


def test_synth_pileup_properties():
    rng = np.random.default_rng(0)
    seq = synth.random_sequence(rng, 20000, 0.5, 1e-4)
    p = synth.synth_pileup(seq, rng, depth=30)
    key = p["position"] * 8 + p["strand"] * 4 + p["mod_type"]
    assert len(np.unique(key)) == len(key)  # rows unique per (position, strand, mod_type)
    assert (np.diff(p["position"]) >= 0).all()
    # 'a' rows sit on A ('+') or T ('-')
    a = p["mod_type"] == 0
    assert set(seq[p["position"][a & (p["strand"] == 0)]].tolist()) == {ord("A")}
    assert set(seq[p["position"][a & (p["strand"] == 1)]].tolist()) == {ord("T")}
    # two-decimal percentages: integer compare on round(100*percent) is exact
    pct = np.round(p["fraction_mod"] * 100, 2)
    assert np.array_equal(p["fraction_mod"] >= 0.7, np.round(pct * 100) >= 7000)
    motifs = synth.random_motifs(np.random.default_rng(1), 50, "A")
    for s, mp in motifs:
        toks = M.tokenize(s)
        assert toks[mp] == "A" and toks[0] != "." and toks[-1] != "." and 4 <= len(toks) <= 21


def test_motifs_per_item():
    from nanomotif_b200.device import choose_motifs_per_item, make_jobs

    sms = 148
    j = make_jobs(3)  # bench.py cfg2: 3 mod types x 1000 motifs x 71 tiles
    j["motif_count"], j["tile_count"] = 1000, 71
    assert choose_motifs_per_item(j, sms) == 32
    j = make_jobs(1)  # a search step: few motifs, many items wanted
    j["motif_count"], j["tile_count"] = 4, 71
    assert choose_motifs_per_item(j, sms) == 1
    j["motif_count"], j["tile_count"] = 640, 71
    assert choose_motifs_per_item(j, sms) == 19
    j["motif_count"], j["tile_count"] = 0, 0
    assert choose_motifs_per_item(j, sms) == 1


def test_bgzf_block_table_and_pileup_blocks():
    """Host-side pieces of the ingest path: the BGZF header walk and the per-mod-type split of compact rows."""
    import gzip
    import struct
    import zlib

    from nanomotif_b200 import dataload
    from nanomotif_b200.pipeline import blocks_by_modtype

    payload = [b"contig_1\t0\t1\ta\n" * 3000, b"", b"x" * 70000]
    z = bytearray()
    for chunk in payload + [b""]:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = co.compress(chunk[:0xff00]) + co.flush()
        z += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(raw) + 8 - 1)
        z += raw + struct.pack("<II", zlib.crc32(chunk[:0xff00]), len(chunk[:0xff00]))
    blocks = dataload.bgzf_blocks(bytes(z))
    assert blocks is not None and len(blocks["in_len"]) == 4
    assert blocks["out_len"].tolist() == [len(payload[0]), 0, 0xff00, 0] and blocks["total"] == len(payload[0]) + 0xff00
    assert blocks["out_off"].tolist() == [0, len(payload[0]), len(payload[0]), len(payload[0]) + 0xff00]
    for off, n, want, crc in zip(blocks["in_off"], blocks["in_len"], [payload[0], b"", payload[2][:0xff00], b""], blocks["crc"]):
        data = zlib.decompress(bytes(z[off:off + n]), -15)
        assert data == want and zlib.crc32(data) == crc
    assert dataload.bgzf_blocks(gzip.compress(b"plain gzip member")) is None
    assert dataload.bgzf_blocks(b"contig_1\t0\t1\n") is None
    assert dataload.bgzf_blocks(bytes(z[:-5])) is None  # truncated file

    rng = np.random.default_rng(0)
    n_contigs, n = 3, 1000
    cid = np.sort(rng.integers(0, n_contigs, n))
    flags = (rng.integers(0, 2, n) | (rng.integers(0, 3, n) << 1)).astype(np.uint8)
    pos = rng.integers(0, 10**6, n).astype(np.int32)
    key = rng.integers(0, 10001, n).astype(np.uint16)
    off = np.concatenate([[0], np.cumsum(np.bincount(cid, minlength=n_contigs))]).astype(np.int64)
    blocks = blocks_by_modtype(pos, flags, key, off, 3)
    assert [b.modtypes for b in blocks] == [(0,), (1,), (2,)] and sum(len(b.position) for b in blocks) == n
    for t, b in enumerate(blocks):
        sel = (flags >> 1) == t
        np.testing.assert_array_equal(b.position, pos[sel])
        np.testing.assert_array_equal(b.percent_x100, key[sel])
        np.testing.assert_array_equal(np.diff(b.contig_row_off), np.bincount(cid[sel], minlength=n_contigs))


def test_blocks_by_position_cut_on_tile_borders():
    from nanomotif_b200 import _lib
    from nanomotif_b200.device import plan_layout
    from nanomotif_b200.pipeline import blocks_by_position

    rng = np.random.default_rng(1)
    lengths = [200000, 50000, 300000]
    pos, cid = [], []
    for c, L in enumerate(lengths):
        p = np.sort(rng.choice(L, size=L // 3, replace=False))
        pos.append(p)
        cid.append(np.full(len(p), c))
    pos, cid = np.concatenate(pos).astype(np.int32), np.concatenate(cid)
    flags = (rng.integers(0, 2, len(pos)) | (rng.integers(0, 3, len(pos)) << 1)).astype(np.uint8)
    key = rng.integers(0, 10001, len(pos)).astype(np.uint16)
    off = np.concatenate([[0], np.cumsum(np.bincount(cid, minlength=3))]).astype(np.int64)
    starts, n_tiles = plan_layout(np.array(lengths))
    blocks = blocks_by_position(pos, flags, key, off, lengths, 3)
    assert 2 <= len(blocks) <= 4 and sum(len(b.position) for b in blocks) == len(pos)
    covered = []
    at = 0
    for b in blocks:
        n = len(b.position)
        tiles = (starts[cid[at:at + n]] + pos[at:at + n]) // _lib.TILE_BP
        t0, tn = b.tiles
        assert tiles.min() == t0 and tiles.max() == t0 + tn - 1 and b.modtypes == (0, 1, 2)
        np.testing.assert_array_equal(np.diff(b.contig_row_off), np.bincount(cid[at:at + n], minlength=3))
        covered.append((t0, t0 + tn))
        at += n
    assert all(a[1] <= b[0] for a, b in zip(covered[:-1], covered[1:]))  # disjoint tile ranges, ascending


def test_column_kl_is_scipy_entropy_bit_for_bit():
    from scipy.stats import entropy

    from nanomotif_b200.growth import column_kl

    rng = np.random.default_rng(9)
    for _ in range(200):
        w = int(rng.integers(1, 62))
        p = rng.integers(0, 50, size=(4, w)).astype(np.float64)
        q = rng.integers(0, 50, size=(4, w)).astype(np.float64)
        p[:, rng.integers(0, w)] = [7, 0, 0, 0]      # a column with zeros in p
        q[rng.integers(0, 4), rng.integers(0, w)] = 0  # q = 0 where p may be > 0 -> inf
        if rng.random() < 0.3:
            p, q = p / max(p.sum(), 1), q / max(q.sum(), 1)
        with np.errstate(all="ignore"):
            want = entropy(p, q)
        got = column_kl(p, q)
        assert got.tobytes() == np.asarray(want, dtype=np.float64).tobytes()


def test_window_rows_keep_the_reference_order():
    from nanomotif_b200.growth import window_rows

    rng = np.random.default_rng(12)
    lens = np.array([500, 90, 41, 42, 3000])
    n = 4000
    cid = rng.integers(-1, 5, n)
    pos = rng.integers(0, 3000, n)
    strand = rng.integers(0, 2, n)
    frac = rng.choice([0.2, 0.69, 0.7, 0.9], n)
    pad, high = 20, 0.7
    want_c, want_p, want_s = [], [], []
    for c in range(len(lens)):  # the per-contig, per-strand passes of find_motifs_bin.py:625-672
        sel = (frac >= high) & (cid == c)
        for s in (0, 1):
            p = pos[sel & (strand == s)]
            p = p[(p > pad) & (p < lens[c] - pad)]
            want_c += [c] * len(p)
            want_p += p.tolist()
            want_s += [s] * len(p)
    got_c, got_p, got_s = window_rows(lens, cid, pos, strand, frac, high, pad)
    assert got_c.tolist() == want_c and got_p.tolist() == want_p and got_s.tolist() == want_s and len(want_p) > 300


def test_load_fasta_host(tmp_path):
    """fasta.py:35-49 + the upper-casing of seq.py:55: names cut at the first blank, multi-line records, gzip."""
    import gzip

    from nanomotif_b200 import dataload

    text = ">c1 first contig\nacgtNN\nACGT\n\n>c2\tx\nTTTT\r\n>c3\n"
    (tmp_path / "a.fa").write_text(text)
    (tmp_path / "a.fa.gz").write_bytes(gzip.compress(text.encode()))
    want = {"c1": "ACGTNNACGT", "c2": "TTTT", "c3": ""}
    assert dataload.load_fasta(str(tmp_path / "a.fa")) == want
    assert dataload.load_fasta(str(tmp_path / "a.fa.gz")) == want
    assert list(dataload.load_fasta(str(tmp_path / "a.fa"), trim_names=True, trim_character="1")) == ["c", "c2\tx", "c3"]


def test_mtstream_reproduces_random_sample_and_its_state():
    """growth.MTStream draws random.sample(range(n), k) natively (nmb_mt_sample: MT19937 + the two branches of
    Lib/random.py on a copy of the Python generator's state): same picks in the same order, and the Python generator
    continues exactly where the loop of random.sample calls (seq.py:202-225, one per contig) would have left it --
    set branch, pool branch and k = 0 alike."""
    import random

    from nanomotif_b200.growth import MTStream

    rng = np.random.default_rng(0)
    cases = [(16000, 650), (100, 50), (62000, 2500), (1250000, 50000), (300, 50), (278, 50), (277, 50), (5, 5), (1000, 0),
             (70000, 50), (2**20, 100), (2**20 + 1, 1000), (4097, 4000)]
    for _ in range(30):
        n = int(rng.integers(60, 400000))
        cases.append((n, int(min(n, max(50, rng.integers(1, n // 20 + 2))))))
    random.seed(2403)  # the reference's import-time seed (seq.py:6)
    random.random()
    state = random.getstate()
    want = [random.sample(range(n), k) for n, k in cases]
    tail = [random.random() for _ in range(5)]
    random.setstate(state)
    stream = MTStream()
    got = [stream.sample(n, k) for n, k in cases]
    stream.sync()
    assert [g.tolist() for g in got] == want
    assert [random.random() for _ in range(5)] == tail
    with pytest.raises(ValueError):
        MTStream().sample(5, 6)
    random.setstate(state)  # one native call for many consecutive samples
    stream = MTStream()
    many = stream.sample_many([n for n, _ in cases], [k for _, k in cases])
    stream.sync()
    assert many.tolist() == [x for w in want for x in w]
    assert [random.random() for _ in range(5)] == tail


def test_string_columns_of_every_frame_kind_give_the_same_strings():
    """dataload._string_column: Arrow string / large_string / dictionary / chunked / sliced arrays, numpy object and
    unicode arrays and pandas Series all come out as offsets + bytes (or codes + dictionary) of the same strings."""
    import pandas as pd
    import pyarrow as pa

    from nanomotif_b200.dataload import _string_column

    values = ["contig_1", "contig_10", "", "k141_7 flag", "c", "contig_1"] * 50

    def strings_of(col, **kw):
        kind, a, b = _string_column(col, **kw)
        if kind == "utf8":
            data = bytes(b)
            return [data[a[i]:a[i + 1]].decode() for i in range(len(a) - 1)]
        assert kind == "dict"
        return [b[i] for i in a]

    arr = pa.array(values, type=pa.string())
    for col in (arr, pa.array(values, type=pa.large_string()), arr.dictionary_encode(), pa.chunked_array([arr[:100], arr[100:]]),
                np.array(values, dtype=object), np.array(values), pd.Series(values)):
        assert strings_of(col) == values
    assert strings_of(arr.slice(7, 120)) == values[7:127]                      # offsets that do not start at 0
    assert strings_of(pa.array(values, type=pa.large_string()).slice(200, 33)) == values[200:233]
    kind, codes, _ = _string_column(np.array([0, 1, 1, 0], dtype=np.uint8))    # strand codes pass through
    assert kind == "codes" and codes.tolist() == [0, 1, 1, 0]
    assert strings_of(np.array([21839, 21839, 7]), ints_are_codes=False) == ["21839", "21839", "7"]  # modkit codes


def test_name_hashes_vectorised_equal_scalar_fnv1a():
    """The contig-name table of the device parsers (nmb_bed_parse / nmb_lookup_strings) hashes every name with FNV-1a;
    the vectorised hash must equal the byte-by-byte definition the kernels implement."""
    from nanomotif_b200 import dataload as D

    names = [f"contig_{i}" for i in range(3000)] + ["", "a", "é", "x" * 70, "bin_1|contig 9", "NODE_1_length_5000_cov_3.2"]
    enc = [n.encode() for n in names]
    got = D._fnv1a64_many(enc)
    assert got.dtype == np.uint64 and got.tolist() == [D._fnv1a64(b) for b in enc]
    assert D._fnv1a64(b"") == 0xCBF29CE484222325 and D._fnv1a64(b"a") == 0xAF63DC4C8601EC8C  # published FNV-1a vectors
    assert D._fnv1a64_many([]).shape == (0,) and D._fnv1a64_many([b""]).tolist() == [0xCBF29CE484222325]
