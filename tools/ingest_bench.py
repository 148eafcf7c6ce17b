#!/usr/bin/env python
"""Throughput of the device bedMethyl ingest (K6) next to the host CSV reader (development tool).

    python tools/ingest_bench.py [--rows 2000000]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanomotif_b200 import dataload, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    args = ap.parse_args()
    rng = np.random.default_rng(1)
    seq = synth.random_sequence(rng, int(args.rows / 1.5), 0.5)
    p = synth.synth_pileup(seq, rng, depth=30, with_counts=True)
    n = len(p["position"])
    pct = np.char.mod("%.2f", 100.0 * p["n_mod"] / p["Nvalid_cov"])
    pos, cov, nm = p["position"].astype(str), p["Nvalid_cov"].astype(str), p["n_mod"].astype(str)
    mt = np.array(synth.MOD_TYPES)[p["mod_type"]]
    st = np.array(["+", "-"])[p["strand"]]
    end = (p["position"] + 1).astype(str)
    lines = ["\t".join(("contig_0", pos[i], end[i], mt[i], cov[i], st[i], pos[i], end[i], "255,0,0", cov[i], pct[i], nm[i],
                        "0", "0", "0", "0", "1", "0")) for i in range(n)]
    text = ("\n".join(lines) + "\n").encode()
    del lines
    dev = torch.device("cuda", 0)
    pinned = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).pin_memory()
    on_dev = pinned.to(dev)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    t_res = timed(lambda: dataload.parse_bedmethyl(on_dev, ["contig_0"], with_counts=True))
    t_h2d = timed(lambda: dataload.parse_bedmethyl(pinned, ["contig_0"], with_counts=True))
    t_all = timed(lambda: dataload.parse_bedmethyl(pinned, ["contig_0"]).filter_coverage().filter_min_mod_frequency().filter_adjacency())
    with tempfile.NamedTemporaryFile(suffix=".bed") as f:
        f.write(text)
        f.flush()
        t0 = time.perf_counter()
        dataload.load_pileup(f.name, with_counts=True)
        t_host = time.perf_counter() - t0
    # K7: the same text bgzip-compressed (64 KB blocks, level 6), inflated on the device
    import struct
    import zlib

    z = bytearray()
    for i in list(range(0, len(text), 0xff00)) + [len(text)]:
        c = text[i:i + 0xff00]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = co.compress(c) + co.flush()
        z += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(raw) + 8 - 1)
        z += raw + struct.pack("<II", zlib.crc32(c), len(c))
    z = bytes(z)
    t_inf = timed(lambda: dataload.inflate_bgzf_device(z), reps=3)
    t0 = time.perf_counter()
    for _ in range(3):
        blocks = dataload.bgzf_blocks(z)
    t_walk = (time.perf_counter() - t0) / 3
    import ctypes as C  # kernel alone, block table and compressed bytes resident
    from nanomotif_b200._lib import check, lib, ptr
    from nanomotif_b200.device import _stream, _to_device
    comp = _to_device(np.frombuffer(z, dtype=np.uint8), dev)
    tabs = [_to_device(blocks[k] if k != "crc" else blocks[k].view(np.int32), dev) for k in ("in_off", "in_len", "out_off", "out_len", "crc")]
    outb = torch.empty(blocks["total"], dtype=torch.uint8, device=dev)
    status = torch.zeros(len(blocks["in_len"]), dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(4):
        if rep == 1:
            ev[0].record()
        check(lib.nmb_bgzf_inflate(ptr(comp), ptr(tabs[0]), ptr(tabs[1]), ptr(tabs[2]), ptr(tabs[3]), ptr(tabs[4]),
                                   len(blocks["in_len"]), ptr(outb), ptr(status), _stream()), "inflate")
    ev[1].record()
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    t_kernel = ev[0].elapsed_time(ev[1]) / 3e3
    print(f"bgzf: {len(blocks['in_len'])} blocks; header walk on the host {t_walk * 1e3:.2f} ms; inflate kernel alone {t_kernel * 1e3:.2f} ms "
          f"({len(text) / 1e9 / t_kernel:.1f} GB/s of text)")
    t0 = time.perf_counter()
    zlib_out = b"".join(zlib.decompress(z[o:o + n], -15) for o, n in zip(*(dataload.bgzf_blocks(z)[k] for k in ("in_off", "in_len"))))
    t_zlib = time.perf_counter() - t0
    assert zlib_out == text
    gb = len(text) / 1e9
    print(f"bgzf {len(z) / 1e6:.1f} MB -> {len(text) / 1e6:.1f} MB: device inflate incl. H2D + header walk {t_inf * 1e3:8.2f} ms "
          f"({len(text) / 1e9 / t_inf:.1f} GB/s of text), zlib on one host core {t_zlib * 1e3:8.2f} ms")
    print(f"rows {n}  text {gb * 1e3:.1f} MB ({len(text) / n:.1f} B/row)")
    print(f"device parse, text resident      {t_res * 1e3:8.2f} ms  {gb / t_res:7.1f} GB/s  {n / t_res / 1e6:8.1f} Mrows/s")
    print(f"H2D (pinned) + device parse      {t_h2d * 1e3:8.2f} ms  {gb / t_h2d:7.1f} GB/s  {n / t_h2d / 1e6:8.1f} Mrows/s")
    print(f"H2D + parse + the three filters  {t_all * 1e3:8.2f} ms  {gb / t_all:7.1f} GB/s")
    print(f"host loader (pyarrow CSV, {os.cpu_count()} threads) {t_host * 1e3:8.2f} ms  {gb / t_host:7.1f} GB/s")


if __name__ == "__main__":
    main()
