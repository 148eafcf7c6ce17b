"""The C-ABI shared library loads and exports every symbol include/nmb200.h declares (no GPU needed)."""
import ctypes
import os
import re

from nanomotif_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nmb200.h")).read()
    return re.findall(r"NMB_API\s+(?:const\s+char\s*\*|int64_t|int)\s*(nmb_[a-z0-9_]+)\s*\(", text)


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in nmb200.h but not exported by libnmb200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in nanomotif_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_constants_match_header():
    text = open(os.path.join(ROOT, "include", "nmb200.h")).read()

    def const(name):
        return int(re.search(rf"#define {name} (\d+)", text).group(1))

    assert const("NMB_ABI_VERSION") == _lib.ABI_VERSION == _lib.lib.nmb_abi_version()
    for name in ("CHUNK_WORDS", "CHUNK_BP", "TILE_WORDS", "TILE_BP", "TILE_CHUNKS", "HALO_WORDS", "MIN_GAP_BP",
                 "MAX_MOTIF_LEN", "MAX_WINDOW", "MAX_MOTIFS_PER_ITEM"):
        assert const("NMB_" + name) == getattr(_lib, name), name
    assert _lib.lib.nmb_program_bytes() == 256
    assert _lib.MOTIF_DTYPE.itemsize == 64 and _lib.JOB_DTYPE.itemsize == 48
    assert ctypes.sizeof(_lib.NmbAssembly) == 40


def test_argument_errors_do_not_need_a_gpu():
    # invalid arguments are rejected before any CUDA call and leave a message
    rc = _lib.lib.nmb_compile_motifs(None, -1, None, None)
    assert rc == -1 and b"n_motifs" in _lib.lib.nmb_last_error()
    rc = _lib.lib.nmb_compact_positions(None, None, 0, 10, None, None, 0, None, None)
    assert rc == -1
