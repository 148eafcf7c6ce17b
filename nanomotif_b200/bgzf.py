"""BGZF + tabix for modkit pileups: the writer the reference calls ``epymetheus.bgzf_pileup`` and the index-driven
fetch behind ``epymetheus.query_pileup_records`` (nanomotif/dataload.py:102-152, tests/test_cli_commands.py:183,234,
docs/source/required_files.md:21-67; the .tbi is what find_motifs_bin.py's builder demands next to a .gz pileup).

    bgzf_pileup(path)                      text bedMethyl -> path.gz (BGZF) + path.gz.tbi (tabix index), host side
    TabixIndex.read(path_gz + ".tbi")      contig names (= pysam.TabixFile(path).contigs) and virtual-offset spans
    fetch_contigs_device(path_gz, names)   ONLY the BGZF blocks that hold those contigs are read from disk, inflated on
                                           the GPU (K7) and returned as one device text buffer of whole lines

Formats (SAM spec section 4.1 for BGZF, the tabix paper / htslib tbx.c for the index): a BGZF file is a chain of gzip
members of <= 64 KiB of data whose extra field 'BC' stores the member size; a virtual offset is
(file offset of the member) << 16 | (offset inside its inflated data).  The index holds, per contig, the UCSC binning
index (bin -> chunks of virtual offsets), a linear index over 16 kbp windows, and the pseudo-bin 37450 with the
contig's whole span and its record count -- checked against the index tabix itself wrote for the reference's bundled
dataset (tests/golden/geobacillus-plasmids.pileup.bed.gz.tbi).

Compression is host zlib, one block per task in a thread pool: writing is a one-off conversion outside the scoring
path.  Everything on the read side that touches data (inflate, parse) runs on the device.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

BLOCK_DATA = 0xFF00  # bytes of text per block, what bgzip uses
_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_PSEUDO_BIN = 37450
TBX_UCSC = 0x10000


def _member(data: bytes, level: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    bsize = 12 + 6 + len(body) + 8  # header + extra + body + crc/isize
    if bsize > 0x10000:  # incompressible: store
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        body = co.compress(data) + co.flush()
        bsize = 12 + 6 + len(body) + 8
    head = struct.pack("<4BIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + struct.pack("<BBHH", 66, 67, 2, bsize - 1)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def bgzf_compress(data, level: int = 6, threads: int | None = None) -> tuple[bytes, np.ndarray]:
    """BGZF image of `data` (with the 28-byte EOF member) and the file offset of every data member."""
    view = memoryview(data)
    chunks = [bytes(view[i:i + BLOCK_DATA]) for i in range(0, len(view), BLOCK_DATA)]
    with ThreadPoolExecutor(threads or min(16, os.cpu_count() or 1)) as pool:  # zlib releases the GIL
        members = list(pool.map(lambda c: _member(c, level), chunks))
    sizes = np.fromiter((len(m) for m in members), dtype=np.int64, count=len(members))
    offsets = np.zeros(len(members), dtype=np.int64)
    if len(members):
        offsets[1:] = np.cumsum(sizes)[:-1]
    return b"".join(members) + _EOF, offsets


def reg2bin(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    """UCSC binning scheme (tabix paper, htslib hts_reg2bin with min_shift 14, 5 levels); end exclusive."""
    beg = np.asarray(beg, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64) - 1
    out = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, first in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = first + (beg[hit] >> shift)
        done |= hit
    return out


class TabixIndex:
    """Per contig: bins {bin: [(voff_begin, voff_end)]}, linear index, whole span, record count."""

    def __init__(self, names, bins, linear, spans, counts, fmt=TBX_UCSC, cols=(1, 2, 3), meta=ord("#"), skip=0):
        self.names, self.bins, self.linear, self.spans, self.counts = list(names), bins, linear, spans, counts
        self.format, self.cols, self.meta, self.skip = fmt, cols, meta, skip
        self._index = {n: i for i, n in enumerate(self.names)}

    @property
    def contigs(self) -> list[str]:
        return list(self.names)

    def span(self, contig: str) -> tuple[int, int] | None:
        """(virtual offset of the contig's first record, virtual offset just past its last record)."""
        i = self._index.get(contig)
        return None if i is None else self.spans[i]

    # ---- file format ----
    @classmethod
    def read(cls, path: str) -> "TabixIndex":
        import gzip

        with gzip.open(path, "rb") as f:
            return cls.read_bytes(f.read(), path)

    @classmethod
    def read_bytes(cls, d: bytes, path: str = "<bytes>") -> "TabixIndex":
        if d[:4] != b"TBI\x01":
            raise ValueError(f"{path}: not a tabix index")
        n_ref, fmt, c_seq, c_beg, c_end, meta, skip, l_nm = struct.unpack_from("<8i", d, 4)
        names = [b.decode() for b in d[36:36 + l_nm].split(b"\0")[:n_ref]]
        p = 36 + l_nm
        bins, linear, spans, counts = [], [], [], []
        for _ in range(n_ref):
            (n_bin,) = struct.unpack_from("<i", d, p)
            p += 4
            ref_bins, span, count = {}, None, 0
            for _ in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", d, p)
                p += 8
                chunks = [struct.unpack_from("<QQ", d, p + 16 * i) for i in range(n_chunk)]
                p += 16 * n_chunk
                if b == _PSEUDO_BIN:
                    span = chunks[0]
                    count = chunks[1][0] if n_chunk > 1 else 0
                else:
                    ref_bins[b] = chunks
            (n_intv,) = struct.unpack_from("<i", d, p)
            p += 4
            ioff = list(struct.unpack_from(f"<{n_intv}Q", d, p))
            p += 8 * n_intv
            if span is None and ref_bins:  # an index without the pseudo-bin: the union of all chunks
                span = (min(c[0] for cs in ref_bins.values() for c in cs), max(c[1] for cs in ref_bins.values() for c in cs))
            bins.append(ref_bins)
            linear.append(ioff)
            spans.append(span)
            counts.append(count)
        return cls(names, bins, linear, spans, counts, fmt, (c_seq, c_beg, c_end), meta, skip)

    def to_bytes(self) -> bytes:
        names = b"".join(n.encode() + b"\0" for n in self.names)
        out = [b"TBI\x01", struct.pack("<8i", len(self.names), self.format, *self.cols, self.meta, self.skip, len(names)), names]
        for ref_bins, ioff, span, count in zip(self.bins, self.linear, self.spans, self.counts):
            out.append(struct.pack("<i", len(ref_bins) + 1))
            for b, chunks in ref_bins.items():
                out.append(struct.pack("<Ii", b, len(chunks)))
                out += [struct.pack("<QQ", *c) for c in chunks]
            out.append(struct.pack("<Ii", _PSEUDO_BIN, 2) + struct.pack("<QQ", *span) + struct.pack("<QQ", count, 0))
            out.append(struct.pack("<i", len(ioff)) + struct.pack(f"<{len(ioff)}Q", *ioff))
        out.append(struct.pack("<Q", 0))  # n_no_coor
        return b"".join(out)

    def write(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(bgzf_compress(self.to_bytes())[0])


def build_index(contig_codes: np.ndarray, names, beg: np.ndarray, end: np.ndarray, line_voff: np.ndarray,
                end_voff: int) -> TabixIndex:
    """Tabix index of records given in FILE ORDER: contig code, [beg, end) and the virtual offset of every line;
    end_voff = virtual offset just past the last line.  Contigs must be contiguous, positions ascending."""
    n = len(contig_codes)
    change = np.flatnonzero(np.diff(contig_codes)) + 1
    starts = np.concatenate([[0], change]).astype(np.int64)
    stops = np.concatenate([change, [n]]).astype(np.int64)
    order = [int(contig_codes[s]) for s in starts]
    if len(set(order)) != len(order):
        raise ValueError("pileup is not grouped by contig: tabix needs a file sorted by contig and position")
    next_voff = np.concatenate([line_voff[1:], [end_voff]]).astype(np.uint64)
    bins_all = reg2bin(beg, end)
    out_names, bins, linear, spans, counts = [], [], [], [], []
    for code, s, e in zip(order, starts.tolist(), stops.tolist()):
        b, v0, v1, pos = bins_all[s:e], line_voff[s:e], next_voff[s:e], beg[s:e]
        if np.any(np.diff(pos) < 0):
            raise ValueError(f"contig {names[code]}: positions are not ascending")
        run = np.concatenate([[0], np.flatnonzero(np.diff(b)) + 1, [e - s]])  # runs of lines that share a bin
        ref_bins: dict = {}
        for r0, r1 in zip(run[:-1].tolist(), run[1:].tolist()):
            ref_bins.setdefault(int(b[r0]), []).append((int(v0[r0]), int(v1[r1 - 1])))
        # linear index: smallest virtual offset of a record overlapping each 16 kbp window
        w0, w1 = pos >> 14, (end[s:e] - 1) >> 14
        n_win = int(w1.max()) + 1
        ioff = np.full(n_win, np.iinfo(np.uint64).max, dtype=np.uint64)
        np.minimum.at(ioff, w0, v0.astype(np.uint64))
        if np.any(w1 != w0):
            np.minimum.at(ioff, w1, v0.astype(np.uint64))
        for w in range(n_win - 2, -1, -1):  # windows without a record inherit the next one's offset (htslib does the same)
            if ioff[w] == np.iinfo(np.uint64).max:
                ioff[w] = ioff[w + 1]
        out_names.append(str(names[code]))
        bins.append(ref_bins)
        linear.append([int(x) for x in ioff])
        spans.append((int(v0[0]), int(v1[-1])))
        counts.append(e - s)
    return TabixIndex(out_names, bins, linear, spans, counts)


def bgzf_pileup(path: str, out: str | None = None, level: int = 6, threads: int | None = None) -> str:
    """Compress a modkit bedMethyl text file to BGZF and write its tabix index (``-p bed``: sequence column 1, begin
    column 2, end column 3, zero-based half-open) -- the file pair the reference reads with query_pileup_records.
    Returns the path of the .gz file (`path` + ".gz" unless `out` is given)."""
    import pyarrow as pa
    import pyarrow.csv as pacsv

    with open(path, "rb") as f:
        data = f.read()
    out = out or path + ".gz"
    if not data:
        raise ValueError(f"{path} is empty")
    if not data.endswith(b"\n"):
        data += b"\n"
    arr = np.frombuffer(data, dtype=np.uint8)
    nl = np.flatnonzero(arr == 10)
    line_start = np.concatenate([[0], nl[:-1] + 1]).astype(np.int64)
    table = pacsv.read_csv(
        pa.BufferReader(data),
        read_options=pacsv.ReadOptions(column_names=[f"c{i}" for i in range(1, 19)]),
        parse_options=pacsv.ParseOptions(delimiter="\t"),
        convert_options=pacsv.ConvertOptions(column_types={"c1": pa.string(), "c2": pa.int64(), "c3": pa.int64()},
                                             include_columns=["c1", "c2", "c3"]))
    if table.num_rows != len(line_start):
        raise ValueError(f"{path}: {len(line_start)} lines but {table.num_rows} records (blank or comment lines?)")
    contig = table.column("c1").combine_chunks().dictionary_encode()
    codes = contig.indices.to_numpy(zero_copy_only=False).astype(np.int64)
    names = contig.dictionary.to_pylist()
    beg = table.column("c2").to_numpy().astype(np.int64)
    end = table.column("c3").to_numpy().astype(np.int64)
    image, member_off = bgzf_compress(data, level, threads)
    blk = line_start // BLOCK_DATA
    line_voff = (member_off[blk].astype(np.uint64) << np.uint64(16)) | (line_start - blk * BLOCK_DATA).astype(np.uint64)
    end_voff = (len(image) - len(_EOF)) << 16  # the start of the EOF member
    index = build_index(codes, names, beg, end, line_voff, end_voff)
    with open(out, "wb") as f:
        f.write(image)
    index.write(out + ".tbi")
    return out


def tabix_contigs(path_gz: str) -> list[str]:
    """Contig names of an indexed pileup (what ``pysam.TabixFile(path).contigs`` returns, find_motifs_bin.py build())."""
    return TabixIndex.read(path_gz + ".tbi").contigs


def contig_block_ranges(index: TabixIndex, contigs) -> list[tuple[int, int, int, int]]:
    """Merged (file offset of the first member, offset inside it, file offset of the last member, offset inside it)
    ranges that cover the requested contigs, in file order."""
    spans = sorted(s for s in (index.span(c) for c in contigs) if s is not None and s[1] > s[0])
    merged: list[list[int]] = []
    for b, e in spans:
        if merged and b <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([b, e])
    return [(b >> 16, b & 0xFFFF, e >> 16, e & 0xFFFF) for b, e in merged]


def fetch_contigs_device(path_gz: str, contigs, device=None, index: TabixIndex | None = None):
    """The text lines of the requested contigs as ONE device uint8 tensor: only the members that hold them are read
    from disk (virtual offsets of the tabix index), inflated on the GPU (K7, one warp per member) and trimmed to
    whole lines.  Returns (text tensor, compressed bytes read)."""
    import torch

    from .dataload import bgzf_blocks, inflate_bgzf_device
    from .device import _require_cuda

    d = _require_cuda(device)
    index = index or TabixIndex.read(path_gz + ".tbi")
    parts, n_read = [], 0
    with open(path_gz, "rb") as f:
        for c0, u0, c1, u1 in contig_block_ranges(index, contigs):
            f.seek(c1)
            head = f.read(18)
            last_size = 0
            if u1 > 0:  # the range ends inside member c1: that member is needed too
                if len(head) < 18 or head[12:14] != b"BC":
                    raise ValueError(f"{path_gz}: no BGZF member at offset {c1}")
                last_size = struct.unpack_from("<H", head, 16)[0] + 1
            f.seek(c0)
            comp = f.read(c1 + last_size - c0)
            n_read += len(comp)
            blocks = bgzf_blocks(comp)
            if blocks is None:
                raise ValueError(f"{path_gz}: bytes {c0}..{c1 + last_size} are not whole BGZF members")
            text = inflate_bgzf_device(comp, d)
            end = (int(blocks["out_off"][-1]) + u1) if u1 > 0 else int(text.numel())
            parts.append(text[u0:end])
    if not parts:
        return torch.zeros(0, dtype=torch.uint8, device=d), 0
    return (parts[0] if len(parts) == 1 else torch.cat(parts)), n_read
