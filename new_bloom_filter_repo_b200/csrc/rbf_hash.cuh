// rbf_hash.cuh -- exact integer restatement of the reference's per-index hashing,
// shaped for sm_100a: every function is __host__ __device__ so the same code is
// unit-tested on the CPU (tests/test_cabi_host.py) and runs inside the kernels.
//
// Reference semantics reproduced here (all bit-exact):
//   * item -> str(item) -> XXH64(seed)          improved_video_compressor.py:77,78,94
//   * index_i = (h1 + i*h2) % size (big ints)   improved_video_compressor.py:81
//       == ((h1 % m) + i*(h2 % m)) % m
//   * activation: h/(2**64-1) < p_activation    improved_video_compressor.py:94-97
//       == h < T  with T found on the host (rbf_activation_threshold)
//
// B200 shaping: a pixel index is hashed as its DECIMAL STRING.  XXH64's short-input
// path folds the length in first and then consumes the bytes left to right, so all
// indices of one "century" (same i/100, same digit count) share the hash state up to
// their last two characters, and the ten indices of a "decade" share it up to the
// last one.  century_state() is computed once per 100 indices, decade_state() once
// per 10, finish() per index: 9 IMAD + ~9 ALU per hash instead of ~70.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RBF_HD __host__ __device__ __forceinline__
#else
#define RBF_HD inline
#endif

namespace rbf {

static constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL;
static constexpr uint64_t XP2 = 0xC2B2AE3D27D4EB4FULL;
static constexpr uint64_t XP3 = 0x165667B19E3779F9ULL;
static constexpr uint64_t XP4 = 0x85EBCA77C2B2AE63ULL;
static constexpr uint64_t XP5 = 0x27D4EB2F165667C5ULL;

RBF_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// h * c (mod 2^64) for a compile-time constant c.  On the device: IMAD.WIDE.U32 + 2 IMAD.  Left to itself ptxas emits four
// instructions (two IMAD, IMAD.WIDE, IADD: one instruction more for a shorter dependency chain); the kernels have enough
// independent chains per warp to prefer the issue slot -- three multiplies per hash, ~4 % of all instructions of the query.
RBF_HD uint64_t mul64c(uint64_t h, uint64_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t p0, p1;
    asm("{\n .reg .u64 t;\n mul.wide.u32 t, %2, %4;\n mov.b64 {%0, %1}, t;\n mad.lo.u32 %1, %3, %4, %1;\n mad.lo.u32 %1, %2, %5, %1;\n}"
        : "=&r"(p0), "=&r"(p1)
        : "r"((uint32_t)h), "r"((uint32_t)(h >> 32)), "r"((uint32_t)c), "r"((uint32_t)(c >> 32)));
    return (uint64_t)p0 | ((uint64_t)p1 << 32);
#else
    return h * c;
#endif
}

RBF_HD uint64_t avalanche(uint64_t h) {
    h ^= h >> 33; h = mul64c(h, XP2); h ^= h >> 29; h = mul64c(h, XP3); h ^= h >> 32;
    return h;
}
// one trailing byte of the short-input path
RBF_HD uint64_t byte_step(uint64_t h, uint32_t b) { h ^= (uint64_t)b * XP5; return mul64c(rotl64(h, 11), XP1); }
// the 4-byte lane:  h ^= lane*P1 ; h = rotl(h,23)*P2 + P3     (v = h ^ lane*P1 already formed)
RBF_HD uint64_t lane4_fin(uint64_t v) { return mul64c(rotl64(v, 23), XP2) + XP3; }
// the 8-byte lane:  k1 = lane*P2 (already formed) ; h ^= rotl(k1,31)*P1 ; h = rotl(h,27)*P1 + P4
RBF_HD uint64_t lane8_fin(uint64_t h, uint64_t k1) {
    k1 = mul64c(rotl64(k1, 31), XP1); h ^= k1; return mul64c(rotl64(h, 27), XP1) + XP4;
}
RBF_HD uint64_t xround(uint64_t acc, uint64_t lane) { acc += lane * XP2; return mul64c(rotl64(acc, 31), XP1); }

RBF_HD uint64_t rd_le(const uint8_t* p, int nbytes) {
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// Full XXH64 over a byte string (string API rbf:115-116,134; any length).
RBF_HD uint64_t xxh64_bytes(const uint8_t* p, uint32_t len, uint64_t seed) {
    uint32_t i = 0;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        for (; i + 32 <= len; i += 32) {
            v1 = xround(v1, rd_le(p + i, 8));       v2 = xround(v2, rd_le(p + i + 8, 8));
            v3 = xround(v3, rd_le(p + i + 16, 8));  v4 = xround(v4, rd_le(p + i + 24, 8));
        }
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = (h ^ xround(0, v1)) * XP1 + XP4;  h = (h ^ xround(0, v2)) * XP1 + XP4;
        h = (h ^ xround(0, v3)) * XP1 + XP4;  h = (h ^ xround(0, v4)) * XP1 + XP4;
    } else {
        h = seed + XP5;
    }
    h += (uint64_t)len;
    for (; i + 8 <= len; i += 8) h = lane8_fin(h, rd_le(p + i, 8) * XP2);
    if (i + 4 <= len) { h = lane4_fin(h ^ (rd_le(p + i, 4) * XP1)); i += 4; }
    for (; i < len; i++) h = byte_step(h, p[i]);
    return avalanche(h);
}

RBF_HD int ndigits_u32(uint32_t v) {
    int n = 1;
    if (v >= 100000000u) { n += 8; v /= 100000000u; }
    if (v >= 10000u) { n += 4; v /= 10000u; }
    if (v >= 100u) { n += 2; v /= 100u; }
    if (v >= 10u) n += 1;
    return n;
}

// ASCII decimal digits of v packed little-endian: first (most significant) digit in byte 0.
RBF_HD uint64_t ascii_le_u32(uint32_t v, int nd) {      // nd <= 8
    uint64_t a = 0;
    for (int j = 0; j < nd; j++) { a = (a << 8) | (uint64_t)(48u + v % 10u); v /= 10u; }
    return a;
}

// XXH64 of str(v): the slow, obviously-correct route (KATs, API single-item calls).
RBF_HD uint64_t xxh64_decimal(uint32_t v, uint64_t seed) {
    uint8_t s[10];
    int nd = ndigits_u32(v);
    for (int j = nd - 1; j >= 0; j--) { s[j] = (uint8_t)(48u + v % 10u); v /= 10u; }
    return xxh64_bytes(s, (uint32_t)nd, seed);
}

// ---------------------------------------------------------------------------------
// Century / decade / finish decomposition.
// A century is hq = i/100; its strings have L = digits(hq)+2 characters and end in the
// two digits x (tens) and y (units).  Where those two characters fall in XXH64's
// [8-byte lane][4-byte lane][bytes...] consumption order depends only on L:
//   L = 2,3,6,7,10 : x and y are both single-byte steps            -> K_BB
//   L = 5          : x is the top byte of the 4-byte lane, y a byte -> K_4B
//   L = 9          : x is the top byte of the 8-byte lane, y a byte -> K_8B
//   L = 4          : x,y are bytes 2,3 of the 4-byte lane           -> K_44
//   L = 8          : x,y are bytes 6,7 of the 8-byte lane           -> K_88
// Lane products distribute over the digit positions: (base + x<<s)*P == base*P + (x*P)<<s.
// Century 0 holds the 1-digit indices (decade 0) and the 2-digit ones (decades 1..9).
// ---------------------------------------------------------------------------------
enum Kind : int { K_BB = 0, K_4B = 1, K_8B = 2, K_44 = 3, K_88 = 4 };

struct Century {
    uint64_t asc;   // prefix digits (L-2 of them), ASCII, first digit in byte 0
    uint32_t hq;    // i / 100
    int L;          // string length of every index in the century (2 for century 0)
    int kind;
};

RBF_HD Century make_century(uint32_t hq) {
    Century c;
    c.hq = hq;
    if (hq == 0) { c.L = 2; c.asc = 0; c.kind = K_BB; return c; }
    int nd = ndigits_u32(hq);
    c.L = nd + 2;
    c.asc = ascii_le_u32(hq, nd);
    c.kind = (c.L == 4) ? K_44 : (c.L == 5) ? K_4B : (c.L == 8) ? K_88 : (c.L == 9) ? K_8B : K_BB;
    return c;
}

RBF_HD uint64_t seed_h0(uint64_t seed, int L) { return seed + XP5 + (uint64_t)L; }

// Per (century, seed): everything that does not depend on x, y.
RBF_HD uint64_t century_state(const Century& c, uint64_t seed) {
    switch (c.kind) {
    case K_4B: return (uint64_t)((uint32_t)(c.asc & 0xffffffu) | 0x30000000u) * XP1;
    case K_44: return (uint64_t)((uint32_t)(c.asc & 0xffffu) | 0x30300000u) * XP1;
    case K_8B: return ((c.asc & 0x00ffffffffffffffULL) | 0x3000000000000000ULL) * XP2;
    case K_88: return ((c.asc & 0x0000ffffffffffffULL) | 0x3030000000000000ULL) * XP2;
    default: {  // K_BB: consume the L-2 prefix characters exactly as XXH64 would
        uint64_t h = seed_h0(seed, c.L), a = c.asc;
        int r = c.L - 2;
        if (r >= 8) { h = lane8_fin(h, a * XP2); r -= 8; a = 0; }
        if (r >= 4) { h = lane4_fin(h ^ ((uint64_t)(uint32_t)a * XP1)); a >>= 32; r -= 4; }
        for (; r > 0; r--) { h = byte_step(h, (uint32_t)(a & 0xff)); a >>= 8; }
        return h;
    }
    }
}

// Per (decade, seed): fold in the tens digit x.
RBF_HD uint64_t decade_state(const Century& c, uint64_t C, uint64_t seed, uint32_t x) {
    switch (c.kind) {
    case K_4B: return lane4_fin(seed_h0(seed, 5) ^ (C + (((uint64_t)x * XP1) << 24)));
    case K_8B: return lane8_fin(seed_h0(seed, 9), C + (((uint64_t)x * XP2) << 56));
    case K_44: return C + (((uint64_t)x * XP1) << 16);
    case K_88: return C + (((uint64_t)x * XP2) << 48);
    default:
        if (c.hq == 0 && x == 0) return seed_h0(seed, 1);          // indices 0..9: one character
        return byte_step(C, 48u + x);
    }
}

// Per index: fold in the units digit y and avalanche.
RBF_HD uint64_t finish(int kind, uint64_t D, uint64_t seed, uint32_t y) {
    switch (kind) {
    case K_44: return avalanche(lane4_fin(seed_h0(seed, 4) ^ (D + (((uint64_t)y * XP1) << 24))));
    case K_88: return avalanche(lane8_fin(seed_h0(seed, 8), D + (((uint64_t)y * XP2) << 56)));
    default:   return avalanche(byte_step(D, 48u + y));
    }
}


// Compile-time-kind versions used by the staged query kernel (no switch in the inner loops).
template <int KIND>
RBF_HD uint64_t decade_state_t(uint64_t C, uint64_t seed, uint32_t x) {      // hq >= 1 only
    if (KIND == K_4B) return lane4_fin(seed_h0(seed, 5) ^ (C + (((uint64_t)x * XP1) << 24)));
    if (KIND == K_8B) return lane8_fin(seed_h0(seed, 9), C + (((uint64_t)x * XP2) << 56));
    if (KIND == K_44) return C + (((uint64_t)x * XP1) << 16);
    if (KIND == K_88) return C + (((uint64_t)x * XP2) << 48);
    return byte_step(C, 48u + x);
}
template <int KIND>
RBF_HD uint64_t finish_t(uint64_t D, uint64_t seed, uint32_t y) {
    if (KIND == K_44) return avalanche(lane4_fin(seed_h0(seed, 4) ^ (D + (((uint64_t)y * XP1) << 24))));
    if (KIND == K_88) return avalanche(lane8_fin(seed_h0(seed, 8), D + (((uint64_t)y * XP2) << 56)));
    return avalanche(byte_step(D, 48u + y));
}


// rotl(x ^ b, 11) == rotl(x, 11) ^ rotl(b, 11): for the kinds whose last character is a single-byte step the
// rotation of the decade state is hoisted out of the per-position work and the ten per-digit constants
// rotl((48 + y) * P5, 11) come from a table.
RBF_HD constexpr uint64_t rot_digit_const(uint32_t y) {
    return (((uint64_t)(48u + y) * XP5) << 11) | (((uint64_t)(48u + y) * XP5) >> 53);
}
template <int KIND> RBF_HD constexpr bool kind_ends_in_byte() { return KIND == K_BB || KIND == K_4B || KIND == K_8B; }
// Dx = rotl(decade_state, 11) for byte kinds, the plain decade state otherwise
template <int KIND>
RBF_HD uint64_t decade_prep(uint64_t D) { return kind_ends_in_byte<KIND>() ? rotl64(D, 11) : D; }
template <int KIND>
RBF_HD uint64_t finish_prep(uint64_t Dx, uint64_t seed, uint32_t y, uint64_t rot_const) {
    if (kind_ends_in_byte<KIND>()) return avalanche(mul64c(Dx ^ rot_const, XP1));
    return finish_t<KIND>(Dx, seed, y);
}

// ---------------------------------------------------------------------------------
// h mod m for a per-frame constant m (the Bloom size l).  Barrett with M = floor(2^64/m),
// truncated partial products: q_est in [Q-3, Q], so r_est = h - q_est*m < 4m fits 32 bits
// when m <= 2^30 and two conditional subtractions finish it.  8 instructions on sm_100a
// (2 IMAD.HI, 2 IMAD, 2x(IADD, UMIN)).  m > 2^30 or m == 1 take the exact 64-bit `%`.
// ---------------------------------------------------------------------------------
struct FastMod {
    uint32_t m, Mh, Ml, fast;
};

inline FastMod make_fastmod(uint32_t m) {              // host
    FastMod f; f.m = m; f.Mh = 0; f.Ml = 0; f.fast = 0;
    if (m >= 2 && m <= (1u << 30)) {
        uint64_t M = ~0ULL / m;                       // floor((2^64-1)/m)
        if ((m & (m - 1)) == 0) M += 1;               // m | 2^64  ->  floor(2^64/m) is one more
        f.Mh = (uint32_t)(M >> 32); f.Ml = (uint32_t)M; f.fast = 1;
    }
    return f;
}

RBF_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
RBF_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

RBF_HD uint32_t mod_u64(uint64_t h, const FastMod& f) {
    if (f.fast) {
        uint32_t hh = (uint32_t)(h >> 32), hl = (uint32_t)h;
        uint32_t q = hh * f.Mh + umulhi32(hh, f.Ml) + umulhi32(hl, f.Mh);
        uint32_t r = hl - q * f.m;
        r = umin32(r, r - 2u * f.m);
        r = umin32(r, r - f.m);
        return r;
    }
    return (uint32_t)(h % (uint64_t)f.m);
}
// (a + b) mod m for a, b < m  (m < 2^32; the 33-bit sum is handled exactly)
RBF_HD uint32_t addmod(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t s = a + b;
    return (s < a || s >= m) ? s - m : s;
}

// Probe index i of the double hashing, exact for unbounded (h1 + i*h2) % m.
RBF_HD uint32_t probe_index(uint64_t h1, uint64_t h2, uint32_t i, const FastMod& f) {
    uint32_t idx = mod_u64(h1, f), step = mod_u64(h2, f);
    for (uint32_t j = 0; j < i; j++) idx = addmod(idx, step, f.m);
    return idx;
}

// bit j of a packbits (MSB-first) byte stream lives at bit ((j & 31) ^ 7) of LE word j >> 5;
// converting a word between LSB-first (internal) and packbits order reverses the bits of each byte.
RBF_HD uint32_t bitrev_bytes(uint32_t w) {
    w = ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu);
    w = ((w & 0x33333333u) << 2) | ((w >> 2) & 0x33333333u);
    w = ((w & 0x55555555u) << 1) | ((w >> 1) & 0x55555555u);
    return w;
}

}  // namespace rbf
