"""K6: modkit bedMethyl text parsed, filtered and turned into class planes on the device, against the
oracle's plain-Python restatement of dataload.py:72-100 and the host loader."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O


@pytest.fixture(scope="module")
def dl():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200 import dataload

    return dataload


def bed_text(rng, n_contigs=5, length=20000, depth=(4, 40)):
    """Synthetic bedMethyl text in modkit's layout (sorted by contig, position), plus the contig strings."""
    from nanomotif_b200 import synth

    lines, contigs = [], {}
    for i in range(n_contigs):
        L = int(length * (0.3 + rng.random()))
        seq = synth.random_sequence(rng, L, 0.5)
        name = f"contig_{i}" if i != 3 else "NODE_3_length_12_cov_1.5"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=int(rng.integers(*depth)), planted=synth.DEFAULT_PLANTED if i % 2 else (),
                               with_counts=True)
        for j in range(len(p["position"])):
            pos, cov, nm = int(p["position"][j]), int(p["Nvalid_cov"][j]), int(p["n_mod"][j])
            pct = 100.0 * nm / cov
            row = [name, pos, pos + 1, synth.MOD_TYPES[p["mod_type"][j]], cov, "+-"[p["strand"][j]], pos, pos + 1, "255,0,0",
                   cov, f"{pct:.2f}", nm, cov - nm, 0, 0, 0, int(p["n_diff"][j]), 0]
            lines.append("\t".join(str(x) for x in row))
    return lines, contigs


def check_rows(table, want, names_known=None):
    keep = np.ones(len(want["contig"]), dtype=bool) if names_known is None else np.isin(want["contig"], names_known)
    sel = lambda k: np.array(want[k], dtype=object)[keep]
    np.testing.assert_array_equal(table.contig, sel("contig"))
    np.testing.assert_array_equal(table.position, sel("position").astype(np.int64))
    np.testing.assert_array_equal(table.strand, sel("strand"))
    np.testing.assert_array_equal(table.mod_type.astype(str), sel("mod_type").astype(str))
    np.testing.assert_array_equal(table.Nvalid_cov, sel("Nvalid_cov").astype(np.int64))
    got, exp = table.fraction_mod, sel("fraction_mod").astype(np.float64)
    assert got.tobytes() == exp.tobytes()  # bit-identical doubles
    return keep


def test_parse_matches_oracle_and_host_loader(dl, tmp_path):
    rng = np.random.default_rng(17)
    lines, contigs = bed_text(rng)
    names = list(contigs)
    for text in ("\n".join(lines) + "\n", "\n".join(lines), "\r\n".join(lines) + "\r\n", "\n\n".join(lines[:50]) + "\n\n"):
        want = O.load_pileup_text(text)
        rows = dl.parse_bedmethyl(text.encode(), names, with_counts=True)
        t = rows.to_table()
        check_rows(t, want)
        np.testing.assert_array_equal(t.extra["n_mod"], np.array(want["n_mod"], dtype=np.int64))
        np.testing.assert_array_equal(t.extra["n_diff"], np.array(want["n_diff"], dtype=np.int64))
        # every percentage here has two decimals: the fixed-point key reproduces the double exactly
        key = t.extra["percent_x100"].astype(np.float64)
        assert ((key / 100.0) / 100.0).tobytes() == t.fraction_mod.tobytes()
    # the host loader (pyarrow) on the same file
    path = tmp_path / "p.bed"
    path.write_text("\n".join(lines) + "\n")
    host = dl.load_pileup(str(path))
    dev = dl.load_pileup_device(str(path), names).to_table()
    np.testing.assert_array_equal(host.position, dev.position)
    assert host.fraction_mod.tobytes() == dev.fraction_mod.tobytes()
    # rows of contigs missing from the table are dropped (or kept with id -1 on request)
    some = names[:2] + names[3:]
    text = "\n".join(lines) + "\n"
    want = O.load_pileup_text(text)
    keep = check_rows(dl.parse_bedmethyl(text.encode(), some).to_table(), want, some)
    assert 0 < keep.sum() < len(keep)
    assert len(dl.parse_bedmethyl(text.encode(), some, keep_unknown_contigs=True)) == len(keep)


def test_number_formats_and_errors(dl):
    base = ["c", "7", "8", "a", "12", "+", "7", "8", "255,0,0", "12", "50.00", "6", "6", "0", "0", "0", "0", "0"]

    def line(**kw):
        f = list(base)
        for k, v in kw.items():
            f[int(k[1:])] = v
        return "\t".join(f)

    pcts = ["0.00", "100.00", "33.33", "66.67", "7", "7.5", "99.999", "0.125", "12.3456789", "000.10", "1e1", "NA", "null"]
    text = "\n".join(line(f10=p, f1=str(i)) for i, p in enumerate(pcts)) + "\n"
    t = dl.parse_bedmethyl(text.encode(), ["c"]).to_table()
    want = O.load_pileup_text(text)
    for got, exp, p in zip(t.fraction_mod, want["fraction_mod"], pcts):
        if p == "1e1":          # exponent notation is outside the parser's grammar: reported as not-a-number
            assert np.isnan(got)
        elif exp is None:
            assert np.isnan(got)
        else:
            assert got == exp, p
    keys = t.extra["percent_x100"].tolist()
    assert keys[:6] == [0, 10000, 3333, 6667, 700, 750] and keys[6] == 0xFFFF and keys[9] == 10
    with pytest.raises(ValueError):
        dl.parse_bedmethyl(b"c\t1\t2\ta\n", ["c"])
    # strand other than + / -, unknown mod type, 8-byte mod code, negative / huge numbers
    text = "\n".join([line(f5="."), line(f3="zzz"), line(f3="21839"), line(f9="-3"), line(f1="123456789012")]) + "\n"
    rows = dl.parse_bedmethyl(text.encode(), ["c"])
    assert rows.strand.cpu().tolist() == [2, 0, 0, 0, 0] and rows.mod_type.cpu().tolist() == [0, 255, 2, 0, 0]
    assert rows.Nvalid_cov.cpu().tolist()[3] == -3 and rows.position.cpu().tolist()[4] == 123456789012


def test_filters_and_class_planes_on_device(dl):
    import torch

    from nanomotif_b200.device import DeviceAssembly, DevicePileup
    from nanomotif_b200.pileup import strand_codes

    rng = np.random.default_rng(23)
    lines, contigs = bed_text(rng, n_contigs=6, length=30000, depth=(4, 12))
    text = "\n".join(lines) + "\n"
    names = list(contigs)
    rows = dl.parse_bedmethyl(text.encode(), names)
    t0 = rows.to_table()
    r1 = rows.filter_coverage(5)
    k1 = O.filter_pileup(t0.Nvalid_cov)
    t1 = r1.to_table()
    np.testing.assert_array_equal(t1.position, t0.position[k1])
    r2 = r1.filter_min_mod_frequency()
    k2 = O.filter_pileup_minimummod_frequency(t1.contig, t1.mod_type, t1.fraction_mod)
    t2 = r2.to_table()
    np.testing.assert_array_equal(t2.position, t1.position[k2])
    np.testing.assert_array_equal(t2.contig, t1.contig[k2])
    assert 0 < k2.sum() < len(k2)
    r3 = r2.filter_adjacency()
    k3 = O.filter_pileup_adjacency_filter(t2.contig, t2.strand, t2.position, t2.fraction_mod)
    t3 = r3.to_table()
    np.testing.assert_array_equal(t3.position, t2.position[k3])
    assert t3.fraction_mod.tobytes() == t2.fraction_mod[k3].tobytes()
    assert 0 < k3.sum() < len(k3)
    # class planes from device rows == class planes from the host table
    asm = DeviceAssembly.from_sequences(contigs)
    a = r3.class_planes(asm, 0.3, 0.7)
    cid = np.array([names.index(c) for c in t3.contig], dtype=np.int32)
    mt = np.array([("a", "m", "21839").index(m) for m in t3.mod_type], dtype=np.uint8)
    b = DevicePileup.from_columns(asm, cid, t3.position, strand_codes(t3.strand), t3.fraction_mod, 0.3, 0.7, mt, 3)
    assert torch.equal(a.class_records, b.class_records) and int(a.class_records.ne(0).sum()) > 1000


def write_bgzf(data: bytes, level=6, strategy=0, block=0xff00) -> bytes:
    """BGZF writer (SAM spec 4.1) for the tests: independent gzip members with the BC extra subfield + EOF marker."""
    import struct
    import zlib

    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
    for c in chunks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        raw = co.compress(c) + co.flush()
        bsize = 12 + 6 + len(raw) + 8
        out += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += raw + struct.pack("<II", zlib.crc32(c), len(c))
    return bytes(out)


def test_bgzf_inflate_on_device(dl, tmp_path):
    import gzip
    import zlib

    rng = np.random.default_rng(29)
    lines, contigs = bed_text(rng, n_contigs=3, length=9000)
    text = ("\n".join(lines) + "\n").encode()
    cases = {
        "text": text,
        "runs": b"A" * 70000 + b"ACGT" * 30000 + bytes(200000),          # long matches, distance 1 and 4 overlaps
        "random": rng.integers(0, 256, 200001, dtype=np.uint8).tobytes(),  # incompressible: stored blocks
        "tiny": b"x",
        "empty": b"",
    }
    for name, data in cases.items():
        for level, strategy in ((6, 0), (9, 0), (1, 0), (0, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            z = write_bgzf(data, level, strategy)
            assert gzip.decompress(z) == data  # the writer produces valid gzip members
            got = dl.inflate_bgzf_device(z)
            assert got.cpu().numpy().tobytes() == data, (name, level, strategy)
    # corruption is reported: a flipped payload byte breaks the CRC (or the code tables)
    z = bytearray(write_bgzf(text))
    z[len(z) // 2] ^= 0x55
    with pytest.raises(ValueError):
        dl.inflate_bgzf_device(bytes(z))
    assert dl.bgzf_blocks(gzip.compress(text)) is None  # plain gzip is not BGZF
    # the loader: bgzip-compressed, gzip-compressed and plain files give the same rows
    names = list(contigs)
    want = O.load_pileup_text(text.decode())
    for fname, payload in (("p.bed.gz", write_bgzf(text)), ("q.bed.gz", gzip.compress(text)), ("r.bed", text)):
        path = tmp_path / fname
        path.write_bytes(payload)
        check_rows(dl.load_pileup_device(str(path), names).to_table(), want)


def test_fasta_parsed_on_device(dl, tmp_path):
    """parse_fasta_device == load_fasta (names, lengths) and the packed records equal those of the host-loaded strings."""
    import gzip

    import torch

    from nanomotif_b200.device import DeviceAssembly

    rng = np.random.default_rng(41)
    recs = []
    for i, L in enumerate((70001, 1, 5, 131072, 900, 61)):
        seq = "".join(rng.choice(list("ACGTacgtNnRY"), size=L, p=[0.2] * 4 + [0.04] * 4 + [0.01] * 4))
        recs.append((f"contig_{i} some description {i}", seq))
    variants = []
    for width, eol, tail in ((60, "\n", "\n"), (80, "\r\n", "\r\n"), (10 ** 9, "\n", ""), (70, "\n", "\n\n")):
        parts = []
        for name, seq in recs:
            parts.append(">" + name + eol)
            for j in range(0, len(seq), width):
                parts.append(seq[j:j + width] + eol)
            if rng.random() < 0.5:
                parts.append(eol)  # blank line inside / after a record
        text = "".join(parts)
        variants.append(text[:len(text) - len(eol)] + tail if tail != eol else text)
    want = {name.split()[0]: seq.upper() for name, seq in recs}
    ref = DeviceAssembly.from_sequences(want)
    for text in variants:
        asm = dl.parse_fasta_device(text.encode())
        assert asm.names == list(want) and asm.lengths.tolist() == [len(s) for s in want.values()]
        assert torch.equal(asm.seq_records, ref.seq_records) and torch.equal(asm.nonacgt, ref.nonacgt)
    asm = dl.parse_fasta_device(variants[0].encode(), trim_names=True, trim_character="_")
    assert asm.names == ["contig"] * len(recs)
    # files: plain, gzip, bgzip
    for fname, payload in (("a.fa", variants[0].encode()), ("b.fa.gz", gzip.compress(variants[0].encode())),
                           ("c.fa.gz", write_bgzf(variants[0].encode()))):
        path = tmp_path / fname
        path.write_bytes(payload)
        asm = dl.load_fasta_device(str(path))
        assert asm.names == list(want) and torch.equal(asm.seq_records, ref.seq_records)
        assert dl.load_fasta(str(path)) == want


def test_scoring_straight_from_device_rows(dl):
    """bedMethyl text -> K6 rows -> MultiBinScorer without a host copy of the pileup == the string-table path."""
    import nanomotif_b200 as nmb

    rng = np.random.default_rng(51)
    lines, contigs = bed_text(rng, n_contigs=4, length=15000)
    text = "\n".join(lines) + "\n"
    names = list(contigs)
    bins = {"b0": {n: contigs[n] for n in names[:2]}, "b1": {n: contigs[n] for n in names[2:]}}
    want = O.load_pileup_text(text)
    table = dict(contig=np.array(want["contig"], dtype=object), position=np.array(want["position"]),
                 strand=np.array(want["strand"], dtype=object), fraction_mod=np.array(want["fraction_mod"]),
                 mod_type=np.array(want["mod_type"], dtype=object))
    # rows parsed against a DIFFERENT contig order and mod-type list than the scorer's
    rows = dl.parse_bedmethyl(text.encode(), names[::-1] + ["absent"], mod_types=("21839", "a", "m"))
    motifs = [nmb.Motif("GATC", 1), nmb.Motif("CC[AT]GG", 1), nmb.Motif("A", 0), nmb.Motif("CCGG", 0)]
    a = nmb.MultiBinScorer(table, bins, ["a", "m", "21839"], 0.3, 0.7)
    b = nmb.MultiBinScorer(rows, bins, ["a", "m", "21839"], 0.3, 0.7)
    reqs = lambda s: [(s.context(bn, mt), motifs) for bn in bins for mt in ("a", "m", "21839")]
    for x, y in zip(a.score_batch(reqs(a)), b.score_batch(reqs(b))):
        np.testing.assert_array_equal(x, y)
    assert sum(int(x.sum()) for x in a.score_batch(reqs(a))) > 1000
