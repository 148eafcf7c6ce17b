"""Streamed host -> device scoring: the pileup arrives in blocks and scans start while later blocks are
still crossing PCIe.

The reference loads the whole pileup, then scores (find_motifs_bin.py:382-427 then :1265-1331).  On a B200
the scan of an E. coli-sized bin takes about as long as the host->device copy of its pileup rows, so the
end-to-end path overlaps the two: block k's rows are copied on a side stream while the class planes of
block k-1 are built and every job whose mod type is complete is scanned on the compute stream.  Results
are identical to the one-shot path (DevicePileup.from_compact + scan_count): the class planes are a
bitwise OR over rows, so the block order does not matter.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Sequence

import numpy as np
import torch

from ._lib import check, lib, ptr
from .device import (DeviceAssembly, DevicePileup, MotifPrograms, PreparedJobs, _require_cuda, _stream, scan_count,
                     threshold_keys)


class HostBlock(NamedTuple):
    """Compact pileup rows (device.compact_rows) of some mod types; `modtypes` lists the mod-type indices
    that may occur in the block's flags.  Arrays are numpy or (preferably pinned) host tensors.

    `tiles` = (tile_begin, tile_count) declares that the block holds ALL rows of its mod types that fall into
    those tiles of the packed assembly (and none outside): the jobs of these mod types are then scanned tile
    range by tile range, as the blocks arrive, instead of waiting for the mod type's last block.  For pileups
    that arrive in genome order (a bgzip stream).  When the host can choose, one block per mod type is faster:
    measured on bench.py's cfg 2 step, 2.05 ms against 2.34-2.41 ms for three or four tile ranges (each launch
    then holds every mod type's jobs on a third of the tiles: more launches, shorter ones, longer tails)."""

    position: object        # int32 [n]
    flags: object           # uint8 [n]  strand | mod type << 1
    percent_x100: object    # uint16 [n]
    contig_row_off: object  # int64 [n_contigs + 1]
    modtypes: tuple
    tiles: tuple | None = None


def _host_tensor(a, dtype) -> torch.Tensor:
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t


_SIDE_STREAMS: dict[int, torch.cuda.Stream] = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[idx]


def blocks_by_modtype(position, flags, percent_x100, contig_row_off, n_modtypes: int) -> list[HostBlock]:
    """Split compact rows (grouped by contig) into one block per mod type, keeping the contig grouping."""
    position, flags = np.asarray(position), np.asarray(flags)
    percent_x100, off = np.asarray(percent_x100), np.asarray(contig_row_off, dtype=np.int64)
    mt = flags >> 1
    cid = np.repeat(np.arange(len(off) - 1), np.diff(off))
    out = []
    for t in range(n_modtypes):
        sel = np.flatnonzero(mt == t)
        o = np.zeros(len(off), dtype=np.int64)
        np.cumsum(np.bincount(cid[sel], minlength=len(off) - 1), out=o[1:])
        out.append(HostBlock(position[sel], flags[sel], percent_x100[sel], o, (t,)))
    return out


def blocks_by_position(position, flags, percent_x100, contig_row_off, lengths, n_modtypes: int,
                       fractions=(0.125, 0.292, 0.292, 0.291)) -> list[HostBlock]:
    """Split compact rows (grouped by contig, positions ascending inside a contig) into blocks of consecutive
    TILES of the packed assembly, each holding every mod type; `fractions` = share of the rows per block (a
    small first block puts the first scan on the GPU early).  Cuts fall on tile borders."""
    from . import _lib
    from .device import plan_layout

    position, flags = np.asarray(position), np.asarray(flags)
    percent_x100, off = np.asarray(percent_x100), np.asarray(contig_row_off, dtype=np.int64)
    starts, n_tiles = plan_layout(np.asarray(lengths, dtype=np.int64))
    cid = np.repeat(np.arange(len(off) - 1), np.diff(off))
    tile = (starts[cid] + position.astype(np.int64)) // _lib.TILE_BP
    if len(tile) > 1 and np.any(tile[1:] < tile[:-1]):
        raise ValueError("rows must be grouped by contig with ascending positions")
    n = len(position)
    cuts, out = [0], []
    for f in np.cumsum(fractions)[:-1]:
        r = int(min(n, max(cuts[-1], round(f * n))))
        while 0 < r < n and tile[r] == tile[r - 1]:  # move the cut to the next tile border
            r += 1
        cuts.append(r)
    cuts.append(n)
    for b, e in zip(cuts[:-1], cuts[1:]):
        if e <= b:
            continue
        o = np.zeros(len(off), dtype=np.int64)
        np.cumsum(np.bincount(cid[b:e], minlength=len(off) - 1), out=o[1:])
        t0, t1 = int(tile[b]), int(tile[e - 1])
        out.append(HostBlock(position[b:e], flags[b:e], percent_x100[b:e], o, tuple(range(n_modtypes)), (t0, t1 - t0 + 1)))
    return out


def score_host_blocks(names: Sequence[str], lengths, ascii_u8, ascii_off, blocks: Sequence[HostBlock], packed_motifs,
                      jobs: np.ndarray, n_out_rows: int, *, low: float = 0.3, high: float = 0.7, n_modtypes: int = 1,
                      device=None, reduce_over_ranks: bool = False, out_host: torch.Tensor | None = None,
                      motifs_per_item: int | None = None, timeline: dict | None = None) -> torch.Tensor:
    """Pack the contigs, stream the pileup blocks and run every job of `jobs` (numpy _lib.JOB_DTYPE table; a
    job is launched as soon as every block that lists its mod type is on the device).  Returns the int64
    counts [n_out_rows, 4] on the host (`out_host`, pinned, when given).  `reduce_over_ranks` all-reduces the
    counts over the default process group before the copy back (contig-sharded multi-GPU runs)."""
    import time

    d = _require_cuda(device)
    key_low, key_high = threshold_keys(low, high)
    t_start = time.perf_counter()
    mark = (lambda name: timeline.__setitem__(name, (time.perf_counter() - t_start) * 1e3)) if timeline is not None else (lambda name: None)
    with torch.cuda.device(d):
        compute = torch.cuda.current_stream()
        side = _side_stream(d)
        side.wait_stream(compute)  # earlier work on the caller's stream stays ordered before ours
        jobs = np.asarray(jobs).copy()
        staged = []
        n_contigs = len(names)
        with torch.cuda.stream(side):
            # the big copies first, in the order they are needed: nothing else stands between the call and PCIe
            ascii_t = ascii_u8 if isinstance(ascii_u8, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ascii_u8))
            ascii_d = ascii_t.to(d, non_blocking=True)
            ev_ascii = torch.cuda.Event()
            ev_ascii.record(side)
            for b in blocks:
                cols = (_host_tensor(b.position, torch.int32), _host_tensor(b.flags, torch.uint8),
                        _host_tensor(b.percent_x100, torch.uint16), _host_tensor(b.contig_row_off, torch.int64))
                if int(cols[3].numel()) != n_contigs + 1 or len({int(c.numel()) for c in cols[:3]}) != 1:
                    raise ValueError("compact pileup columns differ in length")
                dev_cols = tuple(c.to(d, non_blocking=True) for c in cols)
                ev = torch.cuda.Event()
                ev.record(side)
                staged.append((dev_cols, ev, tuple(b.modtypes)))
        mark("host: big copies issued")

        # compute stream, while the copies run: layout, pack, motif programs, job tables (small H2D copies of their own)
        compute.wait_event(ev_ascii)
        asm = DeviceAssembly(names, lengths, ascii_d, ascii_off, d, sync=False)
        jobs["tile_count"] = np.where(jobs["tile_count"] > 0, jobs["tile_count"], asm.n_tiles)
        progs = MotifPrograms(packed_motifs, d)
        groups = {}
        for j in range(len(jobs)):
            groups.setdefault(int(jobs["modtype"][j]), []).append(j)
        ranged = {mt for b in blocks if b.tiles is not None for mt in b.modtypes}
        if any(b.tiles is None and set(b.modtypes) & ranged for b in blocks):
            raise ValueError("a mod type must come either in tile-ranged blocks or in plain blocks, not both")
        # every launch's job table in ONE small copy: per plain mod type, and per tile-ranged block its mod types'
        # jobs clipped to its tiles
        plain = [mt for mt in groups if mt not in ranged]
        tables = [jobs[groups[mt]] for mt in plain]
        ranged_at = []
        for b in blocks:
            if b.tiles is None:
                ranged_at.append(None)
                continue
            t0, tn = b.tiles
            sel = jobs[np.isin(jobs["modtype"], b.modtypes)].copy()
            lo = np.maximum(sel["tile_begin"], t0)
            hi = np.minimum(sel["tile_begin"] + sel["tile_count"], t0 + tn)
            sel["tile_begin"], sel["tile_count"] = lo, np.maximum(hi - lo, 0)
            sel = sel[sel["tile_count"] > 0]
            ranged_at.append(len(tables) if len(sel) else None)
            if len(sel):
                tables.append(sel)
        all_prepared = PreparedJobs.many(tables, d, motifs_per_item)
        prepared = {mt: all_prepared[i] for i, mt in enumerate(plain)}
        prepared_ranged = [None if i is None else all_prepared[i] for i in ranged_at]
        mark("host: assembly, motifs, jobs issued")
        # class planes block by block, scans as soon as their mod type is complete
        pile = DevicePileup(asm, n_modtypes, low, high)
        view = asm.view()
        out = torch.zeros((n_out_rows, 4), dtype=torch.int64, device=d)
        check(lib.nmb_clear_class_planes(C.byref(view), n_modtypes, ptr(pile.class_records), _stream()),
              "nmb_clear_class_planes")
        pending = {mt: sum(mt in s[2] for s in staged) for mt in prepared}
        launched = set()

        def launch_ready():
            for mt, left in pending.items():
                if left == 0 and mt not in launched:
                    launched.add(mt)
                    scan_count(asm, pile, progs, prepared[mt], n_out_rows, out=out)

        launch_ready()  # jobs whose mod type has no rows at all
        for (dev_cols, ev, mts), sub in zip(staged, prepared_ranged):
            compute.wait_event(ev)
            pos, fl, key, off = dev_cols
            check(lib.nmb_add_class_planes_compact(ptr(pos), ptr(fl), ptr(key), ptr(off), int(pos.numel()), key_low,
                                                   key_high, C.byref(view), n_modtypes, ptr(pile.class_records),
                                                   _stream()), "nmb_add_class_planes_compact")
            if sub is not None:  # tile-ranged block: its tiles are complete for its mod types
                scan_count(asm, pile, progs, sub, n_out_rows, out=out)
            for mt in mts:
                if mt in pending:
                    pending[mt] -= 1
            launch_ready()
        if reduce_over_ranks:
            import torch.distributed as dist

            dist.all_reduce(out)
        mark("host: all launches issued")
        if out_host is None:
            out_host = torch.empty((n_out_rows, 4), dtype=torch.int64, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        compute.synchronize()  # also the point after which the side stream's buffers may be reused
        side.synchronize()
        mark("done")
    return out_host
