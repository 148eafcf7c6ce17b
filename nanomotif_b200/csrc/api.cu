// Misc entry points of libnmb200: version, error string, device query.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace nmb {
char *last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace nmb

extern "C" {

int nmb_abi_version(void) { return NMB_ABI_VERSION; }

const char *nmb_last_error(void) { return nmb::last_error_buffer(); }

int nmb_device_sm_count(void) {
    int dev = 0, n = 0;
    NMB_CUDA(cudaGetDevice(&dev));
    NMB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

int nmb_program_bytes(void) { return nmb::kProgramBytesPerMotif; }

// ---- host: CPython's random.sample(range(n), k) on a transplanted MT19937 state -------------------------------
// The reference draws its background windows with random.sample(valid_starts, n) (nanomotif/seq.py:202-225), whose
// picks depend only on len(valid_starts), n and the generator's 32-bit word stream.  Python spends ~0.5 us per pick;
// cfg 3 needs 45 M of them.  This is the same algorithm (Lib/random.py: sample, _randbelow_with_getrandbits) on the
// standard MT19937 recurrence, so the picks and the state afterwards are bit-identical.
static inline uint32_t mt_next(nmb_mt19937 *s) {
    constexpr int N = 624, M = 397;
    if (s->pos >= N) {
        uint32_t *mt = s->key;
        int kk = 0;
        for (; kk < N - M; ++kk) {
            const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < N - 1; ++kk) {
            const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        const uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        s->pos = 0;
    }
    uint32_t y = s->key[s->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static inline int64_t mt_randbelow(nmb_mt19937 *s, int64_t n, int shift) {  // getrandbits(n.bit_length()) until < n
    for (;;) {
        const int64_t r = (int64_t)(mt_next(s) >> shift);
        if (r < n) return r;
    }
}

int nmb_mt_sample(nmb_mt19937 *state, int64_t n, int64_t k, int64_t *out) {
    NMB_REQUIRE(state && n >= 0 && k >= 0 && k <= n && n < (1ll << 32), "nmb_mt_sample: need 0 <= k <= n < 2^32");
    NMB_REQUIRE(state->pos >= 0 && state->pos <= 624, "nmb_mt_sample: bad generator position");
    if (k == 0) return NMB_OK;
    NMB_REQUIRE(out, "nmb_mt_sample: null output");
    // setsize = 21; if k > 5: setsize += 4 ** ceil(log(k * 3, 4))   (Lib/random.py)
    int64_t setsize = 21;
    if (k > 5) {
        int64_t p = 1;
        while (p < 3 * k) p *= 4;
        setsize += p;
    }
    if (n <= setsize) {  // pool branch: partial shuffle of list(range(n))
        int64_t *pool = (int64_t *)malloc((size_t)n * sizeof(int64_t));
        NMB_REQUIRE(pool, "nmb_mt_sample: out of memory");
        for (int64_t i = 0; i < n; ++i) pool[i] = i;
        for (int64_t i = 0; i < k; ++i) {
            const int64_t m = n - i;
            int bits = 0;
            while ((m >> bits) != 0) ++bits;
            const int64_t j = mt_randbelow(state, m, 32 - bits);
            out[i] = pool[j];
            pool[j] = pool[m - 1];
        }
        free(pool);
        return NMB_OK;
    }
    int bits = 0;
    while ((n >> bits) != 0) ++bits;
    const int shift = 32 - bits;
    uint64_t *seen = (uint64_t *)calloc((size_t)(n / 64 + 1), sizeof(uint64_t));
    NMB_REQUIRE(seen, "nmb_mt_sample: out of memory");
    for (int64_t i = 0; i < k; ++i) {
        int64_t j = mt_randbelow(state, n, shift);
        while (seen[j >> 6] >> (j & 63) & 1) j = mt_randbelow(state, n, shift);
        seen[j >> 6] |= 1ull << (j & 63);
        out[i] = j;
    }
    free(seen);
    return NMB_OK;
}

// `count` consecutive samples from the same stream: sample i = random.sample(range(n[i]), k[i]) written at
// out + sum(k[0..i)) -- one call per bin instead of one per contig.
int nmb_mt_sample_many(nmb_mt19937 *state, const int64_t *n, const int64_t *k, int64_t count, int64_t *out) {
    NMB_REQUIRE(state && count >= 0 && (count == 0 || (n && k)), "nmb_mt_sample_many: bad arguments");
    int64_t at = 0;
    for (int64_t i = 0; i < count; ++i) {
        const int rc = nmb_mt_sample(state, n[i], k[i], out ? out + at : nullptr);
        if (rc != NMB_OK) return rc;
        at += k[i];
    }
    return NMB_OK;
}

// ---- host: motif strings -> nmb_motif records ----------------------------------------------------------------------
// What Motif.new_stripped_motif + Motif.split do before a scan (nanomotif/motif.py:213-245), for a whole batch: a
// lock-step search round packs the <= 4 children of every (bin, mod type) search, and at ~2 us per motif the Python
// loop was half of the host time of a round.  Tokens: A C G T, '.', and [..] classes of A C G T; allowed-set bits in
// the order of nanomotif/constants.py:21-28.  Anything else (other letters, unbalanced or empty brackets, no
// constrained position, too long, mod position outside the stripped motif) sets status 1 and the caller's slow path
// raises the precise error.
static inline int base_bit(unsigned char c) {
    switch (c) {
        case 'A': return 1;
        case 'T': return 2;
        case 'G': return 4;
        case 'C': return 8;
        default: return 0;
    }
}

int nmb_pack_motifs(const char *text, const int64_t *offset, const int32_t *mod_pos, int32_t n, int32_t strip,
                    int32_t mod_pos_override, nmb_motif *out, uint8_t *status) {
    NMB_REQUIRE(n >= 0 && (n == 0 || (text && offset && out && status)), "nmb_pack_motifs: bad arguments");
    NMB_REQUIRE(n == 0 || mod_pos || mod_pos_override >= 0, "nmb_pack_motifs: no mod positions");
    for (int32_t i = 0; i < n; ++i) {
        nmb_motif rec;
        memset(&rec, 0, sizeof(rec));
        memset(&out[i], 0, sizeof(rec));
        status[i] = 1;
        int64_t a = offset[i], b = offset[i + 1];
        if (b < a) continue;
        int64_t mp = mod_pos ? mod_pos[i] : 0;
        if (strip) {  // flanking '.' characters go, mod_position is re-based; an all-wildcard motif stays as it is
            int64_t lead = a;
            while (lead < b && text[lead] == '.') ++lead;
            if (lead < b) {
                mp -= lead - a;
                a = lead;
                while (b > a && text[b - 1] == '.') --b;
            }
        }
        if (mod_pos_override >= 0) mp = mod_pos_override;
        int len = 0, constrained = 0;
        bool ok = true;
        for (int64_t k = a; k < b; ++k) {
            const unsigned char c = (unsigned char)text[k];
            int mask;
            if (c == '.') {
                mask = 0xF;
            } else if (c == '[') {
                mask = 0;
                for (++k; k < b && text[k] != ']'; ++k) {
                    const int bit = base_bit((unsigned char)text[k]);
                    if (!bit) ok = false;
                    mask |= bit;
                }
                if (k >= b || mask == 0) ok = false;  // unmatched bracket / empty class
            } else {
                mask = base_bit(c);
                if (!mask) ok = false;
            }
            if (!ok || len >= NMB_MAX_MOTIF_LEN) {
                ok = false;
                break;
            }
            constrained += mask != 0xF;
            rec.allowed[len++] = (uint8_t)mask;
        }
        if (!ok || len == 0 || constrained == 0 || mp < 0 || mp >= len) continue;
        rec.len = (uint8_t)len;
        rec.mod_pos = (uint8_t)mp;
        out[i] = rec;
        status[i] = 0;
    }
    return NMB_OK;
}

}  // extern "C"
