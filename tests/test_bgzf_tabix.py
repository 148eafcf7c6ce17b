"""BGZF writer + tabix index (the reference's epymetheus.bgzf_pileup / query_pileup_records pair, dataload.py:102-152):
host-side format checks on the CPU, index-driven fetch + device inflate / parse on the GPU."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from nanomotif_b200 import bgzf

HERE = os.path.dirname(os.path.abspath(__file__))


def _bed_text(rng, contigs):
    """modkit-style 18-column lines, sorted by contig and position; returns (text, {contig: its lines})."""
    lines, per = [], {}
    for name, length in contigs:
        pos = np.sort(rng.choice(length, size=min(length, int(rng.integers(200, 4000))), replace=False))
        mine = []
        for p in pos.tolist():
            cov = int(rng.integers(1, 60))
            n_mod = int(rng.integers(0, cov + 1))
            strand = "+-"[int(rng.integers(0, 2))]
            mt = ("a", "m", "21839")[int(rng.integers(0, 3))]
            mine.append(f"{name}\t{p}\t{p + 1}\t{mt}\t{cov}\t{strand}\t{p}\t{p + 1}\t255,0,0\t{cov}\t{100 * n_mod / cov:.2f}\t{n_mod}"
                        f"\t{cov - n_mod}\t0\t0\t0\t{int(rng.integers(0, 4))}\t0")
        per[name] = mine
        lines += mine
    return "\n".join(lines) + "\n", per


@pytest.fixture(scope="module")
def written(tmp_path_factory):
    rng = np.random.default_rng(8)
    contigs = [("contig_7", 300000), ("k141_2", 900), ("contig_10", 70000), ("z", 20000), ("contig_1", 150000)]
    text, per = _bed_text(rng, contigs)
    path = str(tmp_path_factory.mktemp("bgzf") / "pileup.bed")
    with open(path, "w") as f:
        f.write(text)
    gz = bgzf.bgzf_pileup(path)
    return dict(text=text, per=per, gz=gz, contigs=[c for c, _ in contigs])


def test_real_tabix_index_is_read():
    """The index tabix wrote for the reference's bundled dataset (a data fixture): names, format and spans."""
    idx = bgzf.TabixIndex.read(os.path.join(HERE, "golden", "geobacillus-plasmids.pileup.bed.gz.tbi"))
    assert idx.contigs == ["contig_3", "contig_2"]
    assert idx.format == bgzf.TBX_UCSC and idx.cols == (1, 2, 3) and idx.meta == ord("#") and idx.skip == 0
    assert idx.span("contig_3") == (0, (2938871 << 16) | 11451)
    assert idx.span("contig_2") == ((2938871 << 16) | 11451, (5249729 << 16) | 59922)
    assert idx.counts == [3 * 65536 + 2210, 2 * 65536 + 38926] and idx.span("nope") is None
    import gzip as _gzip

    with _gzip.open(os.path.join(HERE, "golden", "geobacillus-plasmids.pileup.bed.gz.tbi"), "rb") as f:
        assert idx.to_bytes() == f.read()  # our serialiser reproduces tabix's own bytes


def test_written_file_is_bgzf_and_gzip(written):
    with open(written["gz"], "rb") as f:
        raw = f.read()
    assert gzip.decompress(raw).decode() == written["text"]  # any gzip reader sees the concatenated members
    assert raw.endswith(bgzf._EOF)
    at, n = 0, 0
    while at < len(raw):  # every member: gzip magic, FEXTRA with the 'BC' subfield, <= 64 KiB in and out
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04" and raw[at + 12:at + 14] == b"BC"
        bsize = struct.unpack_from("<H", raw, at + 16)[0] + 1
        isize = struct.unpack_from("<I", raw, at + bsize - 4)[0]
        assert bsize <= 65536 and isize <= 65536
        at += bsize
        n += 1
    assert at == len(raw) and n == -(-len(written["text"]) // bgzf.BLOCK_DATA) + 1


def test_index_spans_address_exactly_each_contig(written):
    idx = bgzf.TabixIndex.read(written["gz"] + ".tbi")
    assert idx.contigs == written["contigs"] == bgzf.tabix_contigs(written["gz"])
    assert idx.counts == [len(written["per"][c]) for c in written["contigs"]]
    with open(written["gz"], "rb") as f:
        raw = f.read()

    def inflate_member(off):
        bsize = struct.unpack_from("<H", raw, off + 16)[0] + 1
        return zlib.decompress(raw[off + 18:off + bsize - 8], -15), bsize

    def read_range(b, e):  # host-side walk of the virtual-offset range (test helper)
        out, off, u = [], b >> 16, b & 0xFFFF
        while off < (e >> 16):
            data, bsize = inflate_member(off)
            out.append(data[u:])
            off, u = off + bsize, 0
        if e & 0xFFFF:
            out.append(inflate_member(off)[0][u:e & 0xFFFF])
        return b"".join(out).decode()

    for c in written["contigs"]:
        assert read_range(*idx.span(c)) == "\n".join(written["per"][c]) + "\n", c
        # binning index: every chunk of every bin holds only records of that bin; all chunks together = the contig
        i = idx.names.index(c)
        seen = 0
        for b, chunks in idx.bins[i].items():
            for cb, ce in chunks:
                for line in read_range(cb, ce).splitlines():
                    f = line.split("\t")
                    assert f[0] == c and int(bgzf.reg2bin([int(f[1])], [int(f[2])])[0]) == b
                    seen += 1
        assert seen == len(written["per"][c])
        # linear index: window w -> an offset at or before the first record with position >= 16384 w
        for w, v in enumerate(idx.linear[i]):
            first = next((ln for ln in written["per"][c] if int(ln.split("\t")[1]) >= 16384 * w), None)
            if first is not None:
                assert first + "\n" in read_range(v, idx.span(c)[1])


def test_unsorted_pileup_is_refused(tmp_path):
    p = tmp_path / "bad.bed"
    row = lambda c, pos: f"{c}\t{pos}\t{pos + 1}\ta\t9\t+\t{pos}\t{pos + 1}\t255,0,0\t9\t50.00\t4\t5\t0\t0\t0\t0\t0"
    p.write_text("\n".join([row("a", 5), row("b", 1), row("a", 9)]) + "\n")
    with pytest.raises(ValueError, match="grouped by contig"):
        bgzf.bgzf_pileup(str(p))
    p.write_text("\n".join([row("a", 5), row("a", 3)]) + "\n")
    with pytest.raises(ValueError, match="ascending"):
        bgzf.bgzf_pileup(str(p))


@pytest.mark.gpu
def test_fetch_contigs_reads_only_their_blocks_and_parses_like_the_whole_file(written):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200 import dataload

    size = os.path.getsize(written["gz"])
    for want in (["k141_2"], ["contig_10", "contig_7"], ["contig_1", "absent", "z"], written["contigs"]):
        text, n_read = bgzf.fetch_contigs_device(written["gz"], want)
        expect = "".join("\n".join(written["per"][c]) + "\n" for c in written["contigs"] if c in want)
        assert bytes(text.cpu().numpy()).decode() == expect
        if len(want) == 1:
            assert n_read < size / 2  # a small contig costs its own members, not the file
        rows = dataload.load_contigs_pileup_bgzip(written["gz"], want).to_table()
        ref = dataload.parse_bedmethyl(expect.encode(), want).to_table()
        for col in ("contig", "position", "strand", "mod_type", "fraction_mod", "Nvalid_cov"):
            assert getattr(rows, col).tolist() == getattr(ref, col).tolist(), col
        assert len(rows) == sum(len(written["per"][c]) for c in want if c in written["per"])
    empty = dataload.load_contigs_pileup_bgzip(written["gz"], ["absent"])
    assert len(empty) == 0
