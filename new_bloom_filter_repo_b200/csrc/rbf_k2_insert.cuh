// rbf_k2_insert.cuh -- Per-filter constants and K2: insert the set positions into the frame's Bloom filter (ivc:235-237, ivc:99-114).  Included by rbf_kernels.cu inside namespace rbf.
#pragma once

// ------------------------------------------------------------------------------------------
// Per-filter constants kept in registers / shared memory
// ------------------------------------------------------------------------------------------
struct FilterK {
    FastMod fm;
    uint64_t T, s1, s2, sA;
    uint32_t fk, has_act, nm;
};
__device__ __forceinline__ FilterK filter_consts(const FrameJob& J) {
    FilterK k;
    k.fm = J.fm; k.T = J.act_T; k.s1 = J.seed1; k.s2 = J.seed2; k.sA = J.seedA; k.fk = J.floor_k; k.has_act = J.has_act; k.nm = J.neg_m;
    return k;
}

// fast reductions for 2 <= m <= 2^30 (the specialised insert / query paths are only taken then)
__device__ __forceinline__ uint32_t mod_fast(uint64_t h, const FastMod& f, uint32_t neg_m) {
    const uint32_t hh = (uint32_t)(h >> 32), hl = (uint32_t)h;
    // q = hh*Mh + hi(hh*Ml) + hi(hl*Mh);  r = hl - q*m = hl + q*neg_m (neg_m = 2^32 - m from the host), with q's two parts
    // folded into two IMADs (no add, and no zeroed register pair for an accumulating IMAD.HI)
    uint32_t r;
    asm("{\n .reg .u32 a, b;\n mul.hi.u32 a, %1, %4;\n mad.lo.u32 a, %1, %3, a;\n mul.hi.u32 b, %2, %3;\n mad.lo.u32 %0, b, %5, %2;\n mad.lo.u32 %0, a, %5, %0;\n}"
        : "=&r"(r) : "r"(hh), "r"(hl), "r"(f.Mh), "r"(f.Ml), "r"(neg_m));
    r = min(r, r - 2u * f.m);
    return min(r, r - f.m);
}
__device__ __forceinline__ uint32_t addmod_fast(uint32_t a, uint32_t b, uint32_t m) {
    const uint32_t s = a + b;
    return min(s, s - m);
}

// add_index on a global, LSB-first bit array (ivc:99-114) given the three hashes
__device__ __forceinline__ void insert_hashes(uint32_t* __restrict__ bits, const FilterK& K, uint64_t h1, uint64_t h2,
                                              uint64_t hA) {
    uint32_t idx = mod_u64(h1, K.fm);
    const uint32_t step = mod_u64(h2, K.fm);
    for (uint32_t i = 0; i < K.fk; i++) {
        red_or_global(bits + (idx >> 5), 1u << (idx & 31u));
        idx = addmod(idx, step, K.fm.m);
    }
    if (K.has_act && hA < K.T) red_or_global(bits + (idx >> 5), 1u << (idx & 31u));
}

// ------------------------------------------------------------------------------------------
// K2: insert.  One thread owns a century (100 positions); the few set positions of the mask
// are hashed with the shared century/decade prefix states and OR-ed into the bit array in L2.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_insert(const FrameJob* __restrict__ jobs) {
    const FrameJob& J = jobs[blockIdx.y];
    if (J.l == 0) return;
    const FilterK K = filter_consts(J);
    const uint32_t ncent = (J.n + 99u) / 100u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncent; c += gridDim.x * blockDim.x) {
        const uint32_t nvalid = min(100u, J.n - 100u * c);
        Bits128 mb = load_bits100(J.mask, c, nvalid);
        if ((mb.lo | mb.hi) == 0ull) continue;
        const Century cen = make_century(c);
        const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
        for (int half = 0; half < 2; half++) {
            uint64_t v = half ? mb.hi : mb.lo;
            while (v) {
                const uint32_t pos = (uint32_t)(__ffsll((long long)v) - 1) + 64u * half;
                v &= v - 1ull;
                const uint32_t x = pos / 10u, y = pos - 10u * x;
                const uint64_t h1 = finish(cen.kind, decade_state(cen, C1, K.s1, x), K.s1, y);
                const uint64_t h2 = finish(cen.kind, decade_state(cen, C2, K.s2, x), K.s2, y);
                const uint64_t hA = K.has_act ? finish(cen.kind, decade_state(cen, CA, K.sA, x), K.sA, y) : 0ull;
                insert_hashes(J.bits, K, h1, h2, hA);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2 (dense, the default): the same insert with all lanes busy.  The mask is sparse (p ~ 5 %), so a lane
// looping over its own set positions leaves most of the warp idle.  Here a warp takes a slab of
// 32 centuries: the set positions of the slab are compacted (count, warp scan, scatter) into a
// per-warp item list (rounds of at most I2_CAP items), and the list is consumed 32 items at a time:
// the owner's century states come by shuffle, two-character finish, the one-IMAD remainder of the
// query kernels, floor(k)-specialised straight-line probes, RED.OR (fire-and-forget) into the
// L2-resident bit array.
// (Round 2 also tried building the array in shared memory -- a private copy per CTA, merged with one RED
// per non-zero word: 8.0 us per 4K pair against 6.8 for this kernel; shared-memory atomics on random words
// run at < 1 lane-op per clock per SM.  profiles/r02_kbench_variants_smem_sweep.jsonl, insert_variant 2.)
// ------------------------------------------------------------------------------------------
constexpr int I2_WARPS = 4;
constexpr int I2_CAP = 512;                                           // list entries per warp and round (uint16 each)

template <int KIND, int FKT>
__device__ __forceinline__ void insert_batch(const FilterK& K, uint32_t* __restrict__ bits, uint32_t tag, bool have, uint64_t C1,
                                             uint64_t C2, uint64_t CA) {
    const uint32_t owner = (tag >> 7) & 31u, pos = tag & 127u, x = (pos * 205u) >> 11, y = pos - 10u * x;   // pos < 128: /10 exact
    const uint64_t C1o = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)C1, owner) | ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(C1 >> 32), owner) << 32);
    const uint64_t C2o = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)C2, owner) | ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(C2 >> 32), owner) << 32);
    uint64_t CAo = 0;
    if (K.has_act) CAo = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)CA, owner) | ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(CA >> 32), owner) << 32);
    if (!have) return;
    uint32_t idx = mod_fast(finish_t<KIND>(decade_state_t<KIND>(C1o, K.s1, x), K.s1, y), K.fm, K.nm);
    const uint32_t step = mod_fast(finish_t<KIND>(decade_state_t<KIND>(C2o, K.s2, x), K.s2, y), K.fm, K.nm);
    if (FKT > 0) {
#pragma unroll
        for (int p = 0; p < FKT; p++) { red_or_global(bits + (idx >> 5), 1u << (idx & 31u)); idx = addmod_fast(idx, step, K.fm.m); }
    } else {
        for (uint32_t p = 0; p < K.fk; p++) { red_or_global(bits + (idx >> 5), 1u << (idx & 31u)); idx = addmod_fast(idx, step, K.fm.m); }
    }
    if (K.has_act && finish_t<KIND>(decade_state_t<KIND>(CAo, K.sA, x), K.sA, y) < K.T) red_or_global(bits + (idx >> 5), 1u << (idx & 31u));
}

template <int KIND>
__device__ __noinline__ void insert_slab_dense(const FilterK K, uint32_t* __restrict__ bits, Bits128 mb, const Century cen, uint32_t list_addr) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
    uint64_t lo = mb.lo, hi = mb.hi;
    uint32_t total;
    do {                                                             // rounds of at most I2_CAP items (one round unless p > 16 % locally)
        const uint32_t cnt = __popcll(lo) + __popcll(hi);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t off = inc - cnt;
        while (lo && off < (uint32_t)I2_CAP) {
            const uint32_t b = (uint32_t)(__ffsll((long long)lo) - 1); lo &= lo - 1ull;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * off), "h"((uint16_t)((lane << 7) | b)) : "memory");
            off++;
        }
        while (hi && off < (uint32_t)I2_CAP) {
            const uint32_t b = (uint32_t)(__ffsll((long long)hi) + 63); hi &= hi - 1ull;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * off), "h"((uint16_t)((lane << 7) | b)) : "memory");
            off++;
        }
        __syncwarp();
        const uint32_t ntot = min(total, (uint32_t)I2_CAP);
#pragma unroll 1
        for (uint32_t base = 0; base < ntot; base += 32u) {
            const uint32_t i = base + lane;
            uint16_t t16;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t16) : "r"(list_addr + 2u * min(i, (uint32_t)(I2_CAP - 1))) : "memory");
            const bool have = i < ntot;
            if (K.fk == 3u) insert_batch<KIND, 3>(K, bits, t16, have, C1, C2, CA);
            else if (K.fk == 2u) insert_batch<KIND, 2>(K, bits, t16, have, C1, C2, CA);
            else insert_batch<KIND, 0>(K, bits, t16, have, C1, C2, CA);
        }
        __syncwarp();
    } while (total > (uint32_t)I2_CAP);
}

__global__ void __launch_bounds__(I2_WARPS * 32) k_insert2(const FrameJob* __restrict__ jobs) {
    const FrameJob& J = jobs[blockIdx.y];
    if (J.l == 0) return;
    __shared__ __align__(16) uint16_t s_list[I2_WARPS][I2_CAP];
    const FilterK K = filter_consts(J);
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;   // warp-uniform for the compiler
    const uint32_t list_addr = smem_u32(&s_list[warp][0]);
    const uint32_t ncent = (J.n + 99u) / 100u;
    const uint32_t nslab = (ncent + 31u) / 32u;
    const bool fast = K.fm.fast != 0u;                               // 2 <= m <= 2^30: the specialised path
    for (uint32_t sl = blockIdx.x * I2_WARPS + warp; sl < nslab; sl += gridDim.x * I2_WARPS) {
        const uint32_t slab = sl * 32u, c = slab + lane;
        const bool active = c < ncent;
        Bits128 mb; mb.lo = 0; mb.hi = 0;
        if (active) mb = load_bits100(J.mask, c, min(100u, J.n - 100u * c));
        if (!__any_sync(0xffffffffu, (mb.lo | mb.hi) != 0ull)) continue;
        const uint32_t last = min(slab + 31u, ncent - 1u);
        const bool uniform = fast && slab >= 1u && ndigits_u32(slab) == ndigits_u32(last);
        if (uniform) {
            const Century cen = make_century(active ? c : slab);
            switch (cen.kind) {
            case K_4B: insert_slab_dense<K_4B>(K, J.bits, mb, cen, list_addr); break;
            case K_8B: insert_slab_dense<K_8B>(K, J.bits, mb, cen, list_addr); break;
            case K_44: insert_slab_dense<K_44>(K, J.bits, mb, cen, list_addr); break;
            case K_88: insert_slab_dense<K_88>(K, J.bits, mb, cen, list_addr); break;
            default:   insert_slab_dense<K_BB>(K, J.bits, mb, cen, list_addr); break;
            }
        } else if ((mb.lo | mb.hi) != 0ull) {             // century 0, a digit-count boundary or a huge filter: per-lane form
            const Century cen = make_century(c);
            const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
            for (int half = 0; half < 2; half++) {
                uint64_t v = half ? mb.hi : mb.lo;
                while (v) {
                    const uint32_t pos = (uint32_t)(__ffsll((long long)v) - 1) + 64u * half;
                    v &= v - 1ull;
                    const uint32_t x = pos / 10u, y = pos - 10u * x;
                    insert_hashes(J.bits, K, finish(cen.kind, decade_state(cen, C1, K.s1, x), K.s1, y),
                                  finish(cen.kind, decade_state(cen, C2, K.s2, x), K.s2, y),
                                  K.has_act ? finish(cen.kind, decade_state(cen, CA, K.sA, x), K.sA, y) : 0ull);
                }
            }
        }
    }
}
