/*
 * ORACLE (test infrastructure only) -- plain-C restatement of the reference's
 * rational-Bloom-filter hot path, fast enough to check the CUDA path at the
 * 1080p / 4K / 8K sizes of BASELINE.json.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under
 * new_bloom_filter_repo_b200/ links, loads or calls it.
 *
 * Pinned by tests/test_oracle_c.py against oracle/rbf_oracle.py and against the
 * fixtures in tests/golden/ that were generated from the real reference
 * (ross39/new_bloom_filter_repo @ 7e37ed8) by tests/golden/make_golden.py.
 *
 * Reference citations: ivc = improved_video_compressor.py,
 * rbf = rational_bloom_filter.py.  XXH64 is the published algorithm of
 * libxxhash 0.8.2 (python-xxhash, requirements.txt:9), not in /root/reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC ... -lm).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t lane) {
    acc += lane * P2; acc = rotl64(acc, 31); return acc * P1;
}
static inline uint64_t xmerge(uint64_t h, uint64_t v) { h ^= xround(0, v); return h * P1 + P4; }

/* xxhash.xxh64_intdigest(data, seed)  -- call sites ivc:77,78,94; rbf:27,115,116,134 */
uint64_t orc_xxh64(const uint8_t* p, uint64_t len, uint64_t seed) {
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8));
            v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + P5;
    }
    h += len;
    while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* str(item) for a non-negative integer (ivc:77: xxh64_intdigest(str(item), seed)) */
static inline int u64_to_dec(uint64_t v, uint8_t* out) {
    uint8_t tmp[20]; int n = 0;
    do { tmp[n++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    return n;
}

uint64_t orc_hash_index_item(uint64_t item, uint64_t seed) {
    uint8_t s[20]; int n = u64_to_dec(item, s);
    return orc_xxh64(s, (uint64_t)n, seed);
}

/* (h1 + i*h2) % size over unbounded ints (ivc:81) == ((h1%m) + i*(h2%m)) % m */
uint64_t orc_probe_index(uint64_t h1, uint64_t h2, uint64_t i, uint64_t m) {
    unsigned __int128 t = (unsigned __int128)(h1 % m) + (unsigned __int128)i * (h2 % m);
    return (uint64_t)(t % m);
}

/* Correctly-rounded double of h / (2^64 - 1)   (Python int/int true division, ivc:95).
 * h/(2^64-1) = h*2^-64 * (1 + 2^-64 + ...): the exact value sits a hair above
 * h*2^-64, so rounding h to 53 significant bits with the tie broken upwards is exact. */
double orc_unit_div(uint64_t h) {
    if (h == 0) return 0.0;
    int bl = 64 - __builtin_clzll(h);
    if (bl <= 53) return ldexp((double)h, -64);
    int sh = bl - 53;
    uint64_t top = h >> sh, rem = h & ((1ULL << sh) - 1), half = 1ULL << (sh - 1);
    if (rem >= half) top += 1;           /* sticky bit makes an exact half round up */
    return ldexp((double)top, sh - 64);
}

/* T with  h < T  <=>  h/(2^64-1) < p_act   (ivc:95-97).  p_act <= 0 -> 0. */
uint64_t orc_activation_threshold(double p_act) {
    if (!(p_act > 0.0)) return 0;
    if (orc_unit_div(UINT64_MAX) < p_act) return UINT64_MAX; /* p_act > 1: every h (callers keep p_act < 1) */
    uint64_t lo = 0, hi = UINT64_MAX;     /* smallest h with unit_div(h) >= p_act */
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (orc_unit_div(mid) < p_act) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* BloomFilterCompressor._calculate_optimal_params (ivc:161-196); returns l, writes k.
 * (0,0) -> l = 0, *k = 0.  Must be compiled with -ffp-contract=off. */
uint64_t orc_optimal_params(uint64_t n, double p, double* k_out) {
    *k_out = 0.0;
    if (p <= 0.0001) return 0;
    if (p >= 0.32453) return 0;
    double q = 1 - p;
    double L = log(2.0);
    double k = log2(q * pow(L, 2.0) / p);                /* ivc:185 */
    if (isnan(k) || k <= 0) return 0;
    double gamma = 1 / L;
    double lf = p * (double)n * k * gamma;               /* ivc:193, left to right */
    uint64_t l = (uint64_t)lf;
    *k_out = k > 0.1 ? k : 0.1;
    return l > 1 ? l : 1;
}

typedef struct {
    uint64_t size; uint64_t floor_k; uint64_t act_T; int has_act;
    uint64_t s1, s2, sa; uint8_t* bits;                  /* one byte per bit, as ivc:59 */
} orc_filter;

static void filter_init(orc_filter* f, uint64_t size, double k, uint64_t s1, uint64_t s2,
                        uint64_t sa, uint8_t* bits) {
    f->size = size; f->floor_k = (uint64_t)floor(k);
    double p_act = k - floor(k);                         /* ivc:58 */
    f->act_T = orc_activation_threshold(p_act); f->has_act = p_act > 0.0;
    f->s1 = s1; f->s2 = s2; f->sa = sa; f->bits = bits;
}

static inline void filter_add(orc_filter* f, const uint8_t* s, int n) {      /* ivc:99-114 */
    uint64_t h1 = orc_xxh64(s, n, f->s1) % f->size, h2 = orc_xxh64(s, n, f->s2) % f->size;
    uint64_t idx = h1;
    for (uint64_t i = 0; i < f->floor_k; i++) { f->bits[idx] = 1; idx += h2; if (idx >= f->size) idx -= f->size; }
    if (f->has_act && orc_xxh64(s, n, f->sa) < f->act_T) f->bits[idx] = 1;
}

static inline int filter_check(const orc_filter* f, const uint8_t* s, int n) { /* ivc:116-138 */
    uint64_t idx = orc_xxh64(s, n, f->s1) % f->size, h2 = 0;
    if (f->floor_k >= 1) {
        if (!f->bits[idx]) return 0;
    }
    if (f->floor_k >= 2 || f->has_act) h2 = orc_xxh64(s, n, f->s2) % f->size;
    for (uint64_t i = 1; i < f->floor_k; i++) {
        idx += h2; if (idx >= f->size) idx -= f->size;
        if (!f->bits[idx]) return 0;
    }
    if (f->has_act && orc_xxh64(s, n, f->sa) < f->act_T) {
        if (f->floor_k >= 1) { idx += h2; if (idx >= f->size) idx -= f->size; }
        if (!f->bits[idx]) return 0;
    }
    return 1;
}

/* BloomFilterCompressor.compress insert + witness loops (ivc:232-253) for given (k, l).
 * mask: n bytes of 0/1.  bits: l bytes (zeroed here).  witness: up to n bytes.
 * Returns the witness length. */
uint64_t orc_compress_kl(const uint8_t* mask, uint64_t n, double k, uint64_t l,
                         uint64_t s1, uint64_t s2, uint64_t sa,
                         uint8_t* bits, uint8_t* witness) {
    orc_filter f; memset(bits, 0, l); filter_init(&f, l, k, s1, s2, sa, bits);
    uint8_t s[20];
    for (uint64_t i = 0; i < n; i++) if (mask[i] == 1) { int len = u64_to_dec(i, s); filter_add(&f, s, len); }
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; i++) { int len = u64_to_dec(i, s); if (filter_check(&f, s, len)) witness[w++] = mask[i]; }
    return w;
}

/* BloomFilterCompressor.decompress loop (ivc:286-304).  Returns witness bits consumed. */
uint64_t orc_decompress_kl(const uint8_t* bits, uint64_t l, const uint8_t* witness, uint64_t wlen,
                           uint64_t n, double k, uint64_t s1, uint64_t s2, uint64_t sa, uint8_t* out) {
    orc_filter f; filter_init(&f, l, k, s1, s2, sa, (uint8_t*)bits);
    uint8_t s[20]; uint64_t j = 0;
    memset(out, 0, n);
    for (uint64_t i = 0; i < n; i++) {
        int len = u64_to_dec(i, s);
        if (filter_check(&f, s, len)) { out[i] = j < wlen ? witness[j] : 0; j++; }
    }
    return j;
}

/* Filter over arbitrary byte strings (rbf:139-182).  items: concatenated bytes, offs[count+1]. */
void orc_filter_add_strings(uint8_t* bits, uint64_t size, double k, uint64_t s1, uint64_t s2, uint64_t sa,
                            const uint8_t* items, const uint64_t* offs, uint64_t count) {
    orc_filter f; filter_init(&f, size, k, s1, s2, sa, bits);
    for (uint64_t j = 0; j < count; j++) filter_add(&f, items + offs[j], (int)(offs[j + 1] - offs[j]));
}
void orc_filter_check_strings(const uint8_t* bits, uint64_t size, double k, uint64_t s1, uint64_t s2, uint64_t sa,
                              const uint8_t* items, const uint64_t* offs, uint64_t count, uint8_t* out) {
    orc_filter f; filter_init(&f, size, k, s1, s2, sa, (uint8_t*)bits);
    for (uint64_t j = 0; j < count; j++) out[j] = (uint8_t)filter_check(&f, items + offs[j], (int)(offs[j + 1] - offs[j]));
}

/* _calculate_frame_diff mask part (ivc:788-808), interleaved H*W*C frames, Y = channel 0.
 * sample_bytes 1 (uint8) or 2 (uint16, little endian).  Returns the ones count. */
uint64_t orc_frame_diff_mask(const void* prev, const void* curr, uint64_t npix, int channels,
                             int sample_bytes, double threshold, uint8_t* mask) {
    uint64_t ones = 0;
    for (uint64_t i = 0; i < npix; i++) {
        int16_t a, b;
        if (sample_bytes == 1) {
            a = (int16_t)((const uint8_t*)prev)[i * channels];
            b = (int16_t)((const uint8_t*)curr)[i * channels];
        } else {
            a = (int16_t)((const uint16_t*)prev)[i * channels];   /* astype(int16) wraps */
            b = (int16_t)((const uint16_t*)curr)[i * channels];
        }
        int16_t d = (int16_t)(a - b);                            /* int16 arithmetic wraps */
        int16_t ad = (int16_t)(d < 0 ? -d : d);                  /* abs(-32768) stays -32768 */
        uint8_t m = (double)ad > threshold;                      /* ivc:808 */
        mask[i] = m; ones += m;
    }
    return ones;
}
