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
// K2 (dense): the same insert with all lanes busy.  The mask is sparse (p ~ 5 %), so a lane
// looping over its own set positions leaves most of the warp idle.  Here a warp takes a slab of
// 32 centuries: every lane publishes its three century states to shared memory, the set
// positions of the slab are compacted (count, warp scan, scatter) into a per-warp item list, and
// the list is consumed 32 items at a time: two-character finish from the owner's century state,
// Barrett reduction, RED.OR into the bit array.
// ------------------------------------------------------------------------------------------
constexpr int I2_WARPS = 4;
constexpr int I2_LIST = 3200;                                        // worst case: every position of the slab set

template <int KIND>
__device__ __forceinline__ void insert_slab_dense(const FilterK& K, uint32_t* __restrict__ bits, const Bits128 mb,
                                                  const Century& cen, uint64_t* cs, uint16_t* list, uint32_t lane) {
    cs[lane * 3 + 0] = century_state(cen, K.s1);
    cs[lane * 3 + 1] = century_state(cen, K.s2);
    cs[lane * 3 + 2] = century_state(cen, K.sA);
    const uint32_t cnt = __popcll(mb.lo) + __popcll(mb.hi);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t off = inc - cnt;
    uint64_t v = mb.lo;
    while (v) { list[off++] = (uint16_t)((lane << 7) | (uint32_t)(__ffsll((long long)v) - 1)); v &= v - 1ull; }
    v = mb.hi;
    while (v) { list[off++] = (uint16_t)((lane << 7) | (uint32_t)(__ffsll((long long)v) + 63)); v &= v - 1ull; }
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t i = base + lane;
        if (i < total) {
            const uint32_t tag = list[i];
            const uint32_t owner = tag >> 7, pos = tag & 127u, x = pos / 10u, y = pos - 10u * x;
            const uint64_t h1 = finish_t<KIND>(decade_state_t<KIND>(cs[owner * 3 + 0], K.s1, x), K.s1, y);
            const uint64_t h2 = finish_t<KIND>(decade_state_t<KIND>(cs[owner * 3 + 1], K.s2, x), K.s2, y);
            uint32_t idx = mod_u64(h1, K.fm);
            const uint32_t step = mod_u64(h2, K.fm);
            for (uint32_t p = 0; p < K.fk; p++) {
                red_or_global(bits + (idx >> 5), 1u << (idx & 31u));
                idx = addmod(idx, step, K.fm.m);
            }
            if (K.has_act) {
                const uint64_t hA = finish_t<KIND>(decade_state_t<KIND>(cs[owner * 3 + 2], K.sA, x), K.sA, y);
                if (hA < K.T) red_or_global(bits + (idx >> 5), 1u << (idx & 31u));
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(I2_WARPS * 32) k_insert2(const FrameJob* __restrict__ jobs) {
    const FrameJob& J = jobs[blockIdx.y];
    if (J.l == 0) return;
    __shared__ uint64_t s_cs[I2_WARPS][32 * 3];
    __shared__ uint16_t s_list[I2_WARPS][I2_LIST];
    const FilterK K = filter_consts(J);
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;   // warp-uniform for the compiler
    const uint32_t ncent = (J.n + 99u) / 100u;
    const uint32_t nslab = (ncent + 31u) / 32u;
    for (uint32_t sl = blockIdx.x * I2_WARPS + warp; sl < nslab; sl += gridDim.x * I2_WARPS) {
        const uint32_t slab = sl * 32u, c = slab + lane;
        const bool active = c < ncent;
        Bits128 mb; mb.lo = 0; mb.hi = 0;
        if (active) mb = load_bits100(J.mask, c, min(100u, J.n - 100u * c));
        if (!__any_sync(0xffffffffu, (mb.lo | mb.hi) != 0ull)) continue;
        const uint32_t last = min(slab + 31u, ncent - 1u);
        const bool uniform = slab >= 1u && ndigits_u32(slab) == ndigits_u32(last);
        if (uniform) {
            const Century cen = make_century(active ? c : slab);
            switch (cen.kind) {
            case K_4B: insert_slab_dense<K_4B>(K, J.bits, mb, cen, s_cs[warp], s_list[warp], lane); break;
            case K_8B: insert_slab_dense<K_8B>(K, J.bits, mb, cen, s_cs[warp], s_list[warp], lane); break;
            case K_44: insert_slab_dense<K_44>(K, J.bits, mb, cen, s_cs[warp], s_list[warp], lane); break;
            case K_88: insert_slab_dense<K_88>(K, J.bits, mb, cen, s_cs[warp], s_list[warp], lane); break;
            default:   insert_slab_dense<K_BB>(K, J.bits, mb, cen, s_cs[warp], s_list[warp], lane); break;
            }
        } else if ((mb.lo | mb.hi) != 0ull) {             // century 0 or a digit-count boundary: per-lane form
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
