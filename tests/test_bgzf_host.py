"""The DEFLATE decoder of nanomotif_b200/csrc/bgzf.cu (K7), compiled AS HOST CODE by g++ (the device qualifiers are
stripped) and fuzzed against zlib -- the reference implementation of RFC 1951 -- on streams of every block type.
CPU only: checks the algorithm the GPU threads run, not the launch."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MAIN = r'''
#include "core.h"
#include <vector>
#include <zlib.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
int main() {
    srand(1);
    static uint16_t lit_tab[1 << 10], dist_tab[1 << 8];
    for (int t = 0; t < 400; ++t) {
        int n = (t % 7 == 0) ? 0 : rand() % 66000;
        std::vector<uint8_t> data(n + 1);
        int mode = t % 4;
        for (int i = 0; i < n; ++i)
            data[i] = mode == 0 ? rand() & 0xFF : mode == 1 ? "ACGT\t\n0123456789."[rand() % 17] : mode == 2 ? 'A'
                      : (uint8_t)("contig_1\t12\t13\ta\t"[i % 18]);
        int level = (t / 4) % 10, strat = (t / 40) % 5;  // default, filtered, huffman only, rle, fixed
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 9, strat);
        std::vector<uint8_t> comp(deflateBound(&zs, n) + 64);
        zs.next_in = data.data(); zs.avail_in = n; zs.next_out = comp.data(); zs.avail_out = comp.size();
        deflate(&zs, Z_FINISH);
        int clen = zs.total_out;
        deflateEnd(&zs);
        std::vector<uint8_t> c2(comp.begin(), comp.begin() + clen), out(n + 1);
        int produced = -1;
        int st = nmb::inflate_stream(c2.data(), clen, out.data(), n, &produced, 0, 1, lit_tab, dist_tab);
        if (st != 0 || produced != n || memcmp(out.data(), data.data(), n)) {
            printf("FAIL t=%d n=%d level=%d strategy=%d status=%d produced=%d\n", t, n, level, strat, st, produced);
            return 1;
        }
        {   // the chunked CRC-32 of the kernel: chunks combined like zlib's crc32_combine
            int chunk = (n + 31) / 32;
            uint32_t total = 0, shift = nmb::crc_x8n((uint32_t)chunk);
            for (int l = 0; l < 32; ++l) {
                int b0 = std::min(l * chunk, n), b1 = std::min(b0 + chunk, n);
                if (b1 == b0) continue;
                uint32_t c = nmb::crc32_bytes(out.data() + b0, b1 - b0);
                total = l == 0 ? c : nmb::crc_multmodp(b1 - b0 == chunk ? shift : nmb::crc_x8n(b1 - b0), total) ^ c;
            }
            if (total != (uint32_t)crc32(0, data.data(), n)) { printf("CRC FAIL t=%d n=%d\n", t, n); return 1; }
        }
        if (clen > 8) {  // truncated and corrupted streams end with an error or different bytes, never out of bounds
            nmb::inflate_stream(c2.data(), clen / 2, out.data(), n, &produced, 0, 1, lit_tab, dist_tab);
            c2[clen / 3] ^= 0x5A;
            nmb::inflate_stream(c2.data(), clen, out.data(), n, &produced, 0, 1, lit_tab, dist_tab);
        }
    }
    printf("all ok\n");
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_device_inflate_algorithm_against_zlib(tmp_path):
    src = open(os.path.join(ROOT, "nanomotif_b200", "csrc", "bgzf.cu")).read()
    core = src[src.index("namespace nmb {"):src.index("__global__ void __launch_bounds__(128) bgzf_inflate_kernel")]
    core = core.replace("__device__ __forceinline__", "static inline").replace("__device__ const", "static const")
    core = core.replace("__device__ ", "static ")
    core = core.replace("static const volatile", "const volatile").replace("static constexpr", "constexpr")
    (tmp_path / "core.h").write_text("#include <stdint.h>\n#include <stdio.h>\n" + core + "}\n")
    (tmp_path / "main.cpp").write_text(MAIN)
    exe = tmp_path / "fuzz"
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address", "-o", str(exe), str(tmp_path / "main.cpp"), "-lz"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "zlib.h" in build.stderr:
        pytest.skip("zlib development header not available")
    if build.returncode != 0:  # no sanitizer runtime: plain build
        build = subprocess.run(["g++", "-O1", "-o", str(exe), str(tmp_path / "main.cpp"), "-lz"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "all ok" in run.stdout, run.stdout + run.stderr
