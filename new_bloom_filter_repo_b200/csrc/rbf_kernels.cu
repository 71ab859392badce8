// rbf_kernels.cu -- hand-written sm_100a kernels of the rational-Bloom hot path.
//
//   K1  k_threshold      |Y_prev - Y_curr| > thr  -> packed mask + ones count      (ivc:788-808, ivc:211)
//   K2  k_insert         insert every set position into the frame's Bloom filter    (ivc:235-237, ivc:99-114)
//   K3  k_query          Bloom test of ALL n positions -> pass mask                 (ivc:245-253, ivc:116-138)
//   K3b k_witness        ordered witness = mask bits at passing positions           (ivc:253)
//   K4b k_expand         decode: out[i] = witness[rank(i)] for passing i            (ivc:299-304)
//
// The Bloom bit array of the frame being queried is staged into shared memory with
// cp.async.bulk (TMA bulk copy, mbarrier complete_tx); whatever exceeds the 227 KB of a
// CTA is probed through L2.  Integer hashing / bit tests only: no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rbf_kernels.cuh"

namespace rbf {

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA), cache-hinted loads
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {   // read-once frame data: do not pollute L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// fire-and-forget bit set in a global bit array (RED.OR, no return value -> no round trip to wait for)
__device__ __forceinline__ void red_or_global(uint32_t* addr, uint32_t v) {
    asm volatile("red.global.or.b32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// 100-bit helpers (a "century" of positions)
// ------------------------------------------------------------------------------------------
struct Bits128 {
    uint64_t lo, hi;
};

// bits [100c, 100c+nvalid) of a naturally packed bit array (buffer padded by >= 16 B)
__device__ __forceinline__ Bits128 load_bits100(const uint32_t* __restrict__ base, uint32_t c, uint32_t nvalid) {
    uint64_t off = 100ull * c;
    uint32_t w = (uint32_t)(off >> 5), sh = (uint32_t)off & 31u;
    uint32_t a0 = __ldg(base + w), a1 = __ldg(base + w + 1), a2 = __ldg(base + w + 2), a3 = __ldg(base + w + 3);
    uint32_t b0 = __funnelshift_r(a0, a1, sh), b1 = __funnelshift_r(a1, a2, sh), b2 = __funnelshift_r(a2, a3, sh),
             b3 = a3 >> sh;
    Bits128 r;
    r.lo = (uint64_t)b0 | ((uint64_t)b1 << 32);
    r.hi = (uint64_t)b2 | ((uint64_t)b3 << 32);
    if (nvalid >= 64) {
        r.hi &= (1ull << (nvalid - 64)) - 1ull;      // nvalid <= 100
    } else {
        r.hi = 0;
        r.lo &= (1ull << nvalid) - 1ull;              // nvalid < 64
    }
    return r;
}

// OR a (<=128-bit) value into a zero-initialised bit stream at an arbitrary bit offset
__device__ __forceinline__ void or_bits128(uint32_t* __restrict__ W, uint64_t bitoff, uint64_t lo, uint64_t hi) {
    uint32_t w = (uint32_t)(bitoff >> 5), sh = (uint32_t)bitoff & 31u;
    uint32_t s0 = (uint32_t)lo, s1 = (uint32_t)(lo >> 32), s2 = (uint32_t)hi, s3 = (uint32_t)(hi >> 32);
    uint32_t o0 = s0 << sh, o1 = __funnelshift_l(s0, s1, sh), o2 = __funnelshift_l(s1, s2, sh),
             o3 = __funnelshift_l(s2, s3, sh), o4 = __funnelshift_l(s3, 0u, sh);
    if (o0) red_or_global(W + w, o0);
    if (o1) red_or_global(W + w + 1, o1);
    if (o2) red_or_global(W + w + 2, o2);
    if (o3) red_or_global(W + w + 3, o3);
    if (o4) red_or_global(W + w + 4, o4);
}

// ------------------------------------------------------------------------------------------
// K1: threshold + popcount.  One thread = 32 consecutive pixels = one mask word; the pixel
// bytes are read with 128-bit streaming loads (2*PB of them per frame, all issued up front).
// PB = bytes per pixel (channels * sample bytes), S = sample bytes; Y is the first sample.
// ------------------------------------------------------------------------------------------
template <int PB, int S>
__device__ __forceinline__ int absdiff_sample(uint32_t a, uint32_t b) {
    if (S == 1) {
        int d = (int)a - (int)b;
        return d < 0 ? -d : d;
    } else {                                           // numpy int16 wrap-around (ivc:801)
        int16_t x = (int16_t)(uint16_t)a, y = (int16_t)(uint16_t)b;
        int16_t d = (int16_t)(x - y);
        int16_t ad = (int16_t)(d < 0 ? -d : d);      // abs(-32768) stays -32768
        return (int)ad;
    }
}

// 256-bit streaming load: one 32 B sector per thread per instruction (LDG.E.256 on sm_100a)
__device__ __forceinline__ void ldg256_stream(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

// Four consecutive 8-bit YUV444 pixels = three 32-bit words per frame.  Byte-SIMD:
//   nm = 4-bit mask of  |Ya - Yb| > thr  (VABSDIFF4 + per-byte compare),  nd = 4-bit "any byte differs".
// gt_or / gt_and fold the out-of-range thresholds (thr < 0: always, thr > 254: never) into the compare.
__device__ __forceinline__ void yuv8_group4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2,
                                            uint32_t thr4, uint32_t gt_or, uint32_t gt_and, uint32_t& nm, uint32_t& nd) {
    const uint32_t ya = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);   // Y bytes of pixels 0..3
    const uint32_t yb = __byte_perm(__byte_perm(b0, b1, 0x0630), b2, 0x5210);
    const uint32_t gt = (__vcmpgtu4(__vabsdiffu4(ya, yb), thr4) & gt_and) | gt_or;
    nm = ((gt & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t x0 = a0 ^ b0, x1 = a1 ^ b1, x2 = a2 ^ b2;
    const uint32_t f0 = x0 & 0x00ffffffu, f1 = __funnelshift_r(x0, x1, 24) & 0x00ffffffu,
                   f2 = __funnelshift_r(x1, x2, 16) & 0x00ffffffu, f3 = x2 >> 8;
    nd = min(f0, 1u) | (min(f1, 1u) << 1) | (min(f2, 1u) << 2) | (min(f3, 1u) << 3);
}

template <int PB, int S>
__global__ void __launch_bounds__(256) k_threshold(const PairJob* __restrict__ pairs, uint32_t npix, int thr, int any_mode,
                                                   uint32_t* __restrict__ ones, uint32_t* __restrict__ resid) {
    const PairJob pj = pairs[blockIdx.y];
    const uint32_t nwords = (npix + 31u) >> 5;
    uint32_t cnt_ones = 0, cnt_res = 0;
    const uint32_t thr4 = (uint32_t)(thr < 0 ? 0 : (thr > 254 ? 254 : thr)) * 0x01010101u;
    const uint32_t gt_or = thr < 0 ? 0xffffffffu : 0u, gt_and = thr > 254 ? 0u : 0xffffffffu;
    const uint32_t any_mask = any_mode ? 0xfu : 0u;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        const uint32_t px0 = w << 5;
        uint32_t m = 0, r = 0;
        if (px0 + 32u <= npix) {
            if (PB == 3 && S == 1) {
                uint32_t A[24], B[24];
                const uint8_t* pa = pj.prev + (size_t)px0 * 3;
                const uint8_t* pb = pj.curr + (size_t)px0 * 3;
#pragma unroll
                for (int j = 0; j < 3; j++) ldg256_stream(pa + 32 * j, A + 8 * j);
#pragma unroll
                for (int j = 0; j < 3; j++) ldg256_stream(pb + 32 * j, B + 8 * j);
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    uint32_t nm, nd;
                    yuv8_group4(A[3 * g], A[3 * g + 1], A[3 * g + 2], B[3 * g], B[3 * g + 1], B[3 * g + 2], thr4, gt_or, gt_and, nm, nd);
                    nm |= nd & any_mask;
                    m |= nm << (4 * g);
                    r |= (nd & ~nm) << (4 * g);
                }
            } else {
                constexpr int NV = PB;                     // 256-bit loads per 32 pixels
                constexpr int HV = (NV > 3) ? NV / 2 : NV; // at most 3 in flight per frame
                constexpr int HALVES = NV / HV;
                constexpr int PXH = 32 / HALVES;
#pragma unroll
                for (int hf = 0; hf < HALVES; hf++) {
                    uint32_t A[8 * HV], B[8 * HV];
                    const uint8_t* pa = pj.prev + (size_t)(px0 + hf * PXH) * PB;
                    const uint8_t* pb = pj.curr + (size_t)(px0 + hf * PXH) * PB;
#pragma unroll
                    for (int j = 0; j < HV; j++) ldg256_stream(pa + 32 * j, A + 8 * j);
#pragma unroll
                    for (int j = 0; j < HV; j++) ldg256_stream(pb + 32 * j, B + 8 * j);
#pragma unroll
                    for (int k = 0; k < PXH; k++) {
                        const int o = k * PB;              // byte offset of the pixel (compile-time)
                        const uint32_t smask = (S == 1) ? 0xffu : 0xffffu;
                        const uint32_t ya = (A[o >> 2] >> (8 * (o & 3))) & smask;
                        const uint32_t yb = (B[o >> 2] >> (8 * (o & 3))) & smask;
                        uint32_t anyd = 0;                 // any byte of the pixel differs
#pragma unroll
                        for (int q = 0; q < PB; q++) {
                            const int oq = o + q;
                            anyd |= ((A[oq >> 2] ^ B[oq >> 2]) >> (8 * (oq & 3))) & 0xffu;
                        }
                        const uint32_t bit = ((absdiff_sample<PB, S>(ya, yb) > thr) || (any_mode && anyd != 0u)) ? 1u : 0u;
                        m |= bit << (hf * PXH + k);
                        r |= ((anyd != 0u && bit == 0u) ? 1u : 0u) << (hf * PXH + k);
                    }
                }
            }
        } else {                                        // ragged last word: scalar loads
            for (uint32_t k = 0; k < 32u && px0 + k < npix; k++) {
                const uint8_t* a = pj.prev + (size_t)(px0 + k) * PB;
                const uint8_t* b = pj.curr + (size_t)(px0 + k) * PB;
                uint32_t ya = a[0], yb = b[0];
                if (S == 2) { ya |= (uint32_t)a[1] << 8; yb |= (uint32_t)b[1] << 8; }
                uint32_t anyd = 0;
                for (int q = 0; q < PB; q++) anyd |= (uint32_t)(a[q] ^ b[q]);
                const uint32_t bit = ((absdiff_sample<PB, S>(ya, yb) > thr) || (any_mode && anyd != 0u)) ? 1u : 0u;
                m |= bit << k;
                r |= ((anyd != 0u && bit == 0u) ? 1u : 0u) << k;
            }
        }
        pj.mask[w] = m;
        cnt_ones += __popc(m);
        cnt_res += __popc(r);
    }
    // block reduction -> one atomic per block
    __shared__ uint32_t s_o[8], s_r[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt_ones += __shfl_xor_sync(0xffffffffu, cnt_ones, d);
        cnt_res += __shfl_xor_sync(0xffffffffu, cnt_res, d);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_o[warp] = cnt_ones; s_r[warp] = cnt_res; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t to = 0, tr = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) { to += s_o[i]; tr += s_r[i]; }
        if (to) atomicAdd(ones + blockIdx.y, to);
        if (tr) atomicAdd(resid + blockIdx.y, tr);
    }
}

// ------------------------------------------------------------------------------------------
// K1 (TMA variant): persistent CTAs, 4-stage mbarrier ring, one elected thread issues
// cp.async.bulk copies of a 2 x 12 KB tile (prev, curr); for 8-bit YUV444 each lane reads
// four pixels (three words, bank-conflict free) and runs the same byte-SIMD as above; the
// per-lane nibbles are OR-reduced over 8-lane groups (REDUX) into mask words.
// ------------------------------------------------------------------------------------------
constexpr int TMA_STAGES = 4;
constexpr int TMA_TILE_BYTES = 12288;                  // per frame per stage (4096 px at 3 B/px)
constexpr int TMA_THREADS = 256;

template <int PB, int S>
__global__ void __launch_bounds__(TMA_THREADS) k_threshold_tma(const PairJob* __restrict__ pairs, int F, uint32_t npix,
                                                               int thr, int any_mode, uint32_t* __restrict__ ones,
                                                               uint32_t* __restrict__ resid) {
    constexpr uint32_t TP = TMA_TILE_BYTES / PB;       // pixels per tile (multiple of 32*8)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[TMA_STAGES];
    uint8_t* bufA = smem;                               // [STAGES][TILE]
    uint8_t* bufB = smem + TMA_STAGES * TMA_TILE_BYTES;
    const uint32_t tiles_per_frame = npix / TP;         // full tiles only (bulk copies need 16 B multiples)
    const uint64_t total = (uint64_t)tiles_per_frame * (uint64_t)F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NWARP = TMA_THREADS / 32;
    constexpr uint32_t WPW = TP / 32 / NWARP;           // mask words per warp per tile
    const uint32_t thr4 = (uint32_t)(thr < 0 ? 0 : (thr > 254 ? 254 : thr)) * 0x01010101u;
    const uint32_t gt_or = thr < 0 ? 0xffffffffu : 0u, gt_and = thr > 254 ? 0u : 0xffffffffu;
    const uint32_t any_mask = any_mode ? 0xfu : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](uint64_t t, int slot) {
        const uint32_t f = (uint32_t)(t / tiles_per_frame), ti = (uint32_t)(t % tiles_per_frame);
        const PairJob pj = pairs[f];
        mbar_expect_tx(&full[slot], 2 * TMA_TILE_BYTES);
        bulk_g2s(bufA + slot * TMA_TILE_BYTES, pj.prev + (size_t)ti * TMA_TILE_BYTES, TMA_TILE_BYTES, &full[slot]);
        bulk_g2s(bufB + slot * TMA_TILE_BYTES, pj.curr + (size_t)ti * TMA_TILE_BYTES, TMA_TILE_BYTES, &full[slot]);
    };

    // prologue
    uint64_t t0 = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES - 1; s++) {
            uint64_t t = t0 + (uint64_t)s * gridDim.x;
            if (t < total) issue(t, s);
        }
    }
    uint32_t it = 0, acc_o = 0, acc_r = 0, acc_f = 0xffffffffu;     // per-thread counts of the current frame
    auto flush_counts = [&]() {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            acc_o += __shfl_xor_sync(0xffffffffu, acc_o, d);
            acc_r += __shfl_xor_sync(0xffffffffu, acc_r, d);
        }
        if (lane == 0 && acc_f != 0xffffffffu) {
            if (acc_o) atomicAdd(ones + acc_f, acc_o);
            if (acc_r) atomicAdd(resid + acc_f, acc_r);
        }
        acc_o = 0; acc_r = 0;
    };
    for (uint64_t t = t0; t < total; t += gridDim.x, it++) {
        const int slot = it % TMA_STAGES;
        const uint32_t parity = (it / TMA_STAGES) & 1u;
        if (threadIdx.x == 0) {                         // refill the slot freed by the previous iteration
            uint64_t tn = t + (uint64_t)(TMA_STAGES - 1) * gridDim.x;
            if (tn < total) { fence_proxy_async(); issue(tn, (it + TMA_STAGES - 1) % TMA_STAGES); }
        }
        const uint32_t f = (uint32_t)(t / tiles_per_frame), ti = (uint32_t)(t % tiles_per_frame);
        if (f != acc_f) { flush_counts(); acc_f = f; }  // warp-uniform
        mbar_wait(&full[slot], parity);
        const uint8_t* a = bufA + slot * TMA_TILE_BYTES;
        const uint8_t* b = bufB + slot * TMA_TILE_BYTES;
        uint32_t* mask_out = pairs[f].mask + (size_t)ti * (TP / 32) + warp * WPW;
        if (PB == 3 && S == 1) {
            const uint32_t* a32 = reinterpret_cast<const uint32_t*>(a) + (size_t)warp * WPW * 24;   // 32 px = 24 words
            const uint32_t* b32 = reinterpret_cast<const uint32_t*>(b) + (size_t)warp * WPW * 24;
#pragma unroll 2
            for (uint32_t k = 0; k < WPW / 4; k++) {      // 128 pixels (4 mask words) per iteration
                const uint32_t o = (k * 32u + lane) * 3u;
                uint32_t nm, nd;
                yuv8_group4(a32[o], a32[o + 1], a32[o + 2], b32[o], b32[o + 1], b32[o + 2], thr4, gt_or, gt_and, nm, nd);
                nm |= nd & any_mask;
                const uint32_t nr = nd & ~nm;
                const uint32_t grp = 0xffu << (lane & 24);
                const uint32_t wm = __reduce_or_sync(grp, nm << (4 * (lane & 7)));
                const uint32_t wr = __reduce_or_sync(grp, nr << (4 * (lane & 7)));
                if ((lane & 7) == 0) { mask_out[k * 4 + (lane >> 3)] = wm; acc_o += __popc(wm); acc_r += __popc(wr); }
            }
        } else {
            uint32_t myword = 0;
#pragma unroll 4
            for (uint32_t k = 0; k < WPW; k++) {
                const uint32_t px = (warp * WPW + k) * 32u + lane;
                const uint8_t* pa = a + px * PB;
                const uint8_t* pb = b + px * PB;
                uint32_t ya, yb, anyd = 0;
                if (S == 1) { ya = pa[0]; yb = pb[0]; }
                else { ya = *reinterpret_cast<const uint16_t*>(pa); yb = *reinterpret_cast<const uint16_t*>(pb); }
#pragma unroll
                for (int q = 0; q < PB; q += S) {
                    if (S == 1) anyd |= (uint32_t)(pa[q] ^ pb[q]);
                    else anyd |= (uint32_t)(*reinterpret_cast<const uint16_t*>(pa + q) ^ *reinterpret_cast<const uint16_t*>(pb + q));
                }
                const bool bit = (absdiff_sample<PB, S>(ya, yb) > thr) || (any_mode && anyd != 0u);
                const uint32_t bm = __ballot_sync(0xffffffffu, bit);
                const uint32_t br = __ballot_sync(0xffffffffu, (!bit) && anyd != 0u);
                if (lane == (int)k) myword = bm;
                if (lane == 0) { acc_o += __popc(bm); acc_r += __popc(br); }
            }
            if (lane < (int)WPW) mask_out[lane] = myword;
        }
        __syncthreads();                                // slot may be refilled next iteration
    }
    flush_counts();
}

// remainder of each frame after the last full TMA tile: same maths with guarded scalar loads
template <int PB, int S>
__global__ void __launch_bounds__(256) k_threshold_tail(const PairJob* __restrict__ pairs, uint32_t npix, uint32_t px_begin,
                                                        int thr, int any_mode, uint32_t* __restrict__ ones, uint32_t* __restrict__ resid) {
    const PairJob pj = pairs[blockIdx.y];
    const uint32_t w0 = px_begin >> 5, nwords = (npix + 31u) >> 5;
    for (uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        uint32_t m = 0, r = 0;
        for (uint32_t k = 0; k < 32u && (w << 5) + k < npix; k++) {
            const uint8_t* a = pj.prev + (size_t)((w << 5) + k) * PB;
            const uint8_t* b = pj.curr + (size_t)((w << 5) + k) * PB;
            uint32_t ya = a[0], yb = b[0];
            if (S == 2) { ya |= (uint32_t)a[1] << 8; yb |= (uint32_t)b[1] << 8; }
            uint32_t anyd = 0;
            for (int q = 0; q < PB; q++) anyd |= (uint32_t)(a[q] ^ b[q]);
            const uint32_t bit = ((absdiff_sample<PB, S>(ya, yb) > thr) || (any_mode && anyd != 0u)) ? 1u : 0u;
            m |= bit << k;
            r |= ((anyd != 0u && bit == 0u) ? 1u : 0u) << k;
        }
        pj.mask[w] = m;
        if (m) atomicAdd(ones + blockIdx.y, __popc(m));
        if (r) atomicAdd(resid + blockIdx.y, __popc(r));
    }
}

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

// ------------------------------------------------------------------------------------------
// K3: query.  Persistent CTAs split the batch's centuries evenly; for every frame segment the
// CTA stages the frame's bit array into shared memory with TMA bulk copies (the tail that does
// not fit is probed through L2) and tests all positions.  Output: pass mask, 128 bits/century.
// ------------------------------------------------------------------------------------------
constexpr int QT = 512;

struct BitView {
    const uint32_t* sm;      // shared-memory copy of words [0, sm_words)
    const uint32_t* gl;      // whole array in global memory
    uint32_t sm_words;
};
__device__ __forceinline__ uint32_t test_bit(const BitView& bv, uint32_t idx) {
    const uint32_t w = idx >> 5;
    const uint32_t word = (w < bv.sm_words) ? bv.sm[w] : __ldg(bv.gl + w);
    return (word >> (idx & 31u)) & 1u;
}

// check_index (ivc:116-138) for the position with decade states D1, D2, DA and units digit y
__device__ __forceinline__ uint32_t check_one(const BitView& bv, const FilterK& K, int kind, uint64_t D1, uint64_t D2,
                                              uint64_t DA, uint32_t y) {
    uint32_t idx = mod_u64(finish(kind, D1, K.s1, y), K.fm);
    uint32_t ok = 1u;
    if (K.fk >= 1u) ok = test_bit(bv, idx);
    if (ok && (K.fk >= 2u || K.has_act)) {
        const uint32_t step = mod_u64(finish(kind, D2, K.s2, y), K.fm);
        for (uint32_t i = 1; i < K.fk && ok; i++) {
            idx = addmod(idx, step, K.fm.m);
            ok = test_bit(bv, idx);
        }
        if (ok && K.has_act && finish(kind, DA, K.sA, y) < K.T) {
            if (K.fk >= 1u) idx = addmod(idx, step, K.fm.m);
            ok = test_bit(bv, idx);
        }
    }
    return ok;
}

__device__ __forceinline__ Bits128 query_century(const BitView& bv, const FilterK& K, uint32_t c, uint32_t nvalid) {
    const Century cen = make_century(c);
    const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
    Bits128 res; res.lo = 0; res.hi = 0;
#pragma unroll 1
    for (uint32_t x = 0; x < 10u; x++) {
        const uint64_t D1 = decade_state(cen, C1, K.s1, x), D2 = decade_state(cen, C2, K.s2, x),
                       DA = decade_state(cen, CA, K.sA, x);
        uint32_t dres = 0;
#pragma unroll
        for (uint32_t y = 0; y < 10u; y++) dres |= check_one(bv, K, cen.kind, D1, D2, DA, y) << y;
        const uint32_t p0 = 10u * x;
        if (p0 < 64u) {
            res.lo |= (uint64_t)dres << p0;
            if (p0 > 54u) res.hi |= (uint64_t)dres >> (64u - p0);
        } else {
            res.hi |= (uint64_t)dres << (p0 - 64u);
        }
    }
    if (nvalid < 100u) {
        if (nvalid >= 64u) res.hi &= (1ull << (nvalid - 64u)) - 1ull;
        else { res.hi = 0; res.lo &= (1ull << nvalid) - 1ull; }
    }
    return res;
}

__global__ void __launch_bounds__(QT) k_query(const FrameJob* __restrict__ jobs, const uint32_t* __restrict__ cent_prefix,
                                              int F, uint32_t smem_words_cap) {
    extern __shared__ __align__(128) uint32_t sbits[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t total = cent_prefix[F];
    const uint32_t lo = (uint32_t)(((uint64_t)total * blockIdx.x) / gridDim.x);
    const uint32_t hi = (uint32_t)(((uint64_t)total * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (lo >= hi) return;
    // first frame with cent_prefix[f+1] > lo
    int f = 0;
    { int a = 0, b = F; while (a < b) { int m = (a + b) >> 1; if (cent_prefix[m + 1] > lo) b = m; else a = m + 1; } f = a; }
    uint32_t parity = 0, g = lo;
    while (g < hi) {
        while (cent_prefix[f + 1] <= g) f++;
        const FrameJob& J = jobs[f];
        const uint32_t seg_end = min(hi, cent_prefix[f + 1]);
        const uint32_t c_begin = g - cent_prefix[f], c_end = seg_end - cent_prefix[f];
        const uint32_t nwords = (J.l + 31u) >> 5;
        const uint32_t sw = min((nwords + 3u) & ~3u, smem_words_cap);     // 16 B granules; buffer is padded
        __syncthreads();                                                    // everyone done with the previous array
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, sw * 4u);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(J.bits);
            uint8_t* dst = reinterpret_cast<uint8_t*>(sbits);
            for (uint32_t off = 0; off < sw * 4u; off += 32768u)
                bulk_g2s(dst + off, src + off, min(32768u, sw * 4u - off), &bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const FilterK K = filter_consts(J);
        BitView bv; bv.sm = sbits; bv.gl = J.bits; bv.sm_words = sw;
        uint4* pass4 = reinterpret_cast<uint4*>(J.pass);
        for (uint32_t c = c_begin + threadIdx.x; c < c_end; c += QT) {
            const Bits128 r = query_century(bv, K, c, min(100u, J.n - 100u * c));
            pass4[c] = make_uint4((uint32_t)r.lo, (uint32_t)(r.lo >> 32), (uint32_t)r.hi, (uint32_t)(r.hi >> 32));
        }
        g = seg_end;
    }
}

// ------------------------------------------------------------------------------------------
// K3 (staged): the same query as a warp-synchronous pipeline of DENSE stages.
//
// check_index is a pure conjunction (ivc:127-136), so its probes may be evaluated in any order
// and abandoned at the first zero.  Lane t of a warp owns century slab+t; the warp walks the
// 100 positions of its 32 centuries in lockstep:
//   stage A  (all positions)        h1 -> probe 0.  ~1/2 survive (the Bloom fill is ~1/2).
//   stage B  (survivors of A)       h2 -> probes 1..floor_k-1.
//   stage C  (survivors of B)       activation hash -> the floor_k+1'th probe if activated.
// Survivors are compacted through per-warp shared-memory rings (ballot + popc), so stages B and C
// always run with 32 busy lanes instead of diverging per lane.  A stage-B record carries the
// owner's decade state of seed 2; stage C fetches the owner's century state of the activation
// seed with a shuffle.  Positions whose mask bit is set are known to pass (a Bloom filter has no
// false negatives) and skip the hashing.  Results are identical to the per-lane form above.
// ------------------------------------------------------------------------------------------
constexpr int Q2_WARPS = 24;
__constant__ uint64_t c_rot_digit[16] = {rot_digit_const(0), rot_digit_const(1), rot_digit_const(2), rot_digit_const(3),
                                         rot_digit_const(4), rot_digit_const(5), rot_digit_const(6), rot_digit_const(7),
                                         rot_digit_const(8), rot_digit_const(9), 0, 0, 0, 0, 0, 0};
constexpr int Q2_THREADS = Q2_WARPS * 32;
constexpr int Q2_RING = 64;                               // entries per ring (two drains' worth)
constexpr int Q2_WARP_WORDS = (Q2_RING * 16 + Q2_RING * 8 + 32 * 16 + 128) / 4;   // B ring, C ring, pass accumulators, digit table

// explicit shared-space accesses (32-bit shared addresses): no generic-pointer resolution in the hot loops
__device__ __forceinline__ void sts128_if(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool p) {
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %5, 0;\n @q st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n}" ::"r"(addr), "r"(a), "r"(b),
                 "r"(c), "r"(d), "r"((uint32_t)p)
                 : "memory");
}
__device__ __forceinline__ void sts64_if(uint32_t addr, uint32_t a, uint32_t b, bool p) {
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q st.shared.v2.u32 [%0], {%1,%2};\n}" ::"r"(addr), "r"(a), "r"(b),
                 "r"((uint32_t)p)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void red_or_shared_if(uint32_t addr, uint32_t v, bool p) {
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q red.shared.or.b32 [%0], %1;\n}" ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}

// bit `idx` of the Bloom array.  PM (probe mode):
//   0  the whole array is in this CTA's shared memory
//   1  words [0, sm_words) in shared memory, the rest through L2 (read-only path)
//   2  the array is split over the shared memories of a 2-CTA cluster (DSMEM): words [0, sm_words) live in
//      rank 0 (cluster address sm_addr), the rest in rank 1 (sm_addr1 is pre-biased by -4*sm_words)
template <int PM>
__device__ __forceinline__ uint32_t probe_bit(uint32_t sm_addr, uint32_t sm_addr1, const uint32_t* __restrict__ gl,
                                              uint32_t sm_words, uint32_t idx) {
    const uint32_t w = idx >> 5;
    uint32_t word;
    if (PM == 1) {
        // the unconditional mov makes the two predicated loads a full definition for ptxas (otherwise `word` stays live
        // across loop iterations and is spilled right behind the load, stalling on it)
        asm("{\n .reg .pred q;\n setp.lt.u32 q, %1, %2;\n mov.u32 %0, 0;\n @q ld.shared.u32 %0, [%3];\n @!q ld.global.nc.u32 %0, [%4];\n}"
            : "=r"(word)
            : "r"(w), "r"(sm_words), "r"(sm_addr + 4u * w), "l"(gl + w));
    } else if (PM == 2) {
        asm("{\n .reg .pred q;\n .reg .u32 b;\n setp.lt.u32 q, %1, %2;\n selp.u32 b, %3, %4, q;\n mad.lo.u32 b, %1, 4, b;\n"
            " ld.shared::cluster.u32 %0, [b];\n}"
            : "=r"(word)
            : "r"(w), "r"(sm_words), "r"(sm_addr), "r"(sm_addr1));
    } else {
        asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(sm_addr + 4u * w));
    }
    return (word >> (idx & 31u)) & 1u;
}

// fast reductions for 2 <= m <= 2^30 (the staged kernel is only launched then)
__device__ __forceinline__ uint32_t mod_fast(uint64_t h, const FastMod& f, uint32_t neg_m) {
    const uint32_t hh = (uint32_t)(h >> 32), hl = (uint32_t)h;
    const uint32_t q = hh * f.Mh + __umulhi(hh, f.Ml) + __umulhi(hl, f.Mh);
    uint32_t r = q * neg_m + hl;                                       // hl - q*m in one IMAD (neg_m = 2^32 - m from the host)
    r = min(r, r - 2u * f.m);
    return min(r, r - f.m);
}
__device__ __forceinline__ uint32_t addmod_fast(uint32_t a, uint32_t b, uint32_t m) {
    const uint32_t s = a + b;
    return min(s, s - m);
}

// pass bit of (owner lane, x, y) into the owner's 128-bit accumulator
__device__ __forceinline__ void deliver_pass(uint32_t pacc_addr, uint32_t tag, bool p) {
    const uint32_t owner = tag >> 8, pos = 10u * ((tag >> 4) & 15u) + (tag & 15u);
    red_or_shared_if(pacc_addr + 16u * owner + 4u * (pos >> 5), 1u << (pos & 31u), p);
}

struct RingState {                 // warp-uniform
    uint32_t qb_head, qb_cnt, qc_head, qc_cnt;
};

// stage B (32 survivors of A) and stage C (32 survivors of B); `force` drains partial batches
template <int KIND, int FKT, int PM>
__device__ __forceinline__ void drain_stages(const FilterK& K, uint32_t sm_addr, uint32_t sm_addr1, const uint32_t* __restrict__ gl,
                                             uint32_t sm_words, uint32_t qb_addr, uint32_t qc_addr, uint32_t pacc_addr,
                                             uint32_t lane, uint32_t lt, uint64_t CA, RingState& R, bool force) {
    if (R.qb_cnt >= 32u || (force && R.qb_cnt > 0u)) {               // ---- stage B
        __syncwarp();
        const uint32_t nb = min(32u, R.qb_cnt);
        const bool have = lane < nb;
        const uint4 r = lds128(qb_addr + 16u * ((R.qb_head + lane) & (Q2_RING - 1)));
        R.qb_head = (R.qb_head + nb) & (Q2_RING - 1);
        R.qb_cnt -= nb;
        const uint64_t rbB = kind_ends_in_byte<KIND>() ? ({ const uint2 t = lds64(pacc_addr + 512u + 8u * (r.y & 15u)); (uint64_t)t.x | ((uint64_t)t.y << 32); }) : 0ull;
        const uint32_t stepm = have ? mod_fast(finish_prep<KIND>((uint64_t)r.z | ((uint64_t)r.w << 32), K.s2, r.y & 15u, rbB), K.fm, K.nm) : 0u;
        uint32_t idx = have ? r.x : 0u;
        uint32_t ok = have ? 1u : 0u;
        if (FKT > 0) {                                               // floor_k known at compile time: straight-line probes
#pragma unroll
            for (int i = 1; i < FKT; i++) {
                idx = addmod_fast(idx, stepm, K.fm.m);
                ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
            }
        } else {
            for (uint32_t i = 1; i < K.fk; i++) {
                idx = addmod_fast(idx, stepm, K.fm.m);
                ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
                if (!__any_sync(0xffffffffu, ok != 0u)) break;
            }
        }
        if (K.has_act) {
            idx = addmod_fast(idx, stepm, K.fm.m);                   // index of probe floor_k
            const uint32_t b2 = __ballot_sync(0xffffffffu, ok != 0u);
            sts64_if(qc_addr + 8u * ((R.qc_head + R.qc_cnt + __popc(b2 & lt)) & (Q2_RING - 1)), idx, r.y, ok != 0u);
            R.qc_cnt += __popc(b2);
        } else {
            deliver_pass(pacc_addr, r.y, ok != 0u);
        }
    }
    if (R.qc_cnt >= 32u || (force && R.qb_cnt == 0u && R.qc_cnt > 0u)) {   // ---- stage C
        __syncwarp();
        const uint32_t nc = min(32u, R.qc_cnt);
        const bool have = lane < nc;
        const uint2 r = lds64(qc_addr + 8u * ((R.qc_head + lane) & (Q2_RING - 1)));
        R.qc_head = (R.qc_head + nc) & (Q2_RING - 1);
        R.qc_cnt -= nc;
        const uint32_t tag = have ? r.y : 0u;
        const uint32_t owner = tag >> 8;
        const uint64_t CAo = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)CA, owner) |
                             ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(CA >> 32), owner) << 32);
        const uint64_t hA = finish_t<KIND>(decade_state_t<KIND>(CAo, K.sA, (tag >> 4) & 15u), K.sA, tag & 15u);
        const uint32_t pb = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, have ? r.x : 0u);
        deliver_pass(pacc_addr, tag, have && (!(hA < K.T) || pb != 0u));
    }
}

template <int KIND, int FKT, int PM>
__device__ __noinline__ void query_slab_staged(const FilterK K, uint32_t sm_addr, uint32_t sm_addr1, const uint32_t* __restrict__ gl,
                                               uint32_t sm_words, const uint32_t* __restrict__ mask, uint32_t n,
                                               uint32_t slab_c0, uint32_t c_end, uint4* __restrict__ pass4,
                                               uint32_t qb_addr, uint32_t qc_addr, uint32_t pacc_addr) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = slab_c0 + lane;
    const bool active = c < c_end;
    const Century cen = make_century(active ? c : slab_c0);
    const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
    const uint32_t nvalid = active ? min(100u, n - 100u * c) : 0u;
    // positions that need no hashing: known members (mask bit set) and positions beyond n
    uint64_t skip_lo = 0, skip_hi = 0;
    if (active && mask != nullptr) { const Bits128 mb = load_bits100(mask, c, nvalid); skip_lo = mb.lo; skip_hi = mb.hi; }
    if (lane < 10u) sts64_if(pacc_addr + 512u + 8u * lane, (uint32_t)c_rot_digit[lane], (uint32_t)(c_rot_digit[lane] >> 32), true);
    if (nvalid < 64u) { skip_hi = ~0ull; skip_lo |= ~((1ull << nvalid) - 1ull); }
    else skip_hi |= ~((1ull << (nvalid - 64u)) - 1ull);
    const uint32_t lt = (1u << lane) - 1u;
    RingState R; R.qb_head = 0; R.qb_cnt = 0; R.qc_head = 0; R.qc_cnt = 0;
#pragma unroll 1
    for (uint32_t x = 0; x < 10u; x++) {
        const uint64_t D1 = decade_prep<KIND>(decade_state_t<KIND>(C1, K.s1, x));   // rotation hoisted for byte kinds
        const uint64_t D2 = decade_prep<KIND>(decade_state_t<KIND>(C2, K.s2, x));
        const uint32_t p0 = 10u * x;                                 // bits [p0, p0+10) of the 128-bit skip set
        const uint64_t sh = (p0 < 64u) ? ((skip_lo >> p0) | (p0 ? (skip_hi << (64u - p0)) : 0ull)) : (skip_hi >> (p0 - 64u));
        const uint32_t skip10 = (uint32_t)sh & 0x3ffu;
        const uint32_t tagx = (lane << 8) | (x << 4);
#pragma unroll 1
        for (uint32_t y = 0; y < 10u; y += 2u) {                     // ---- stage A: two positions per lane (ILP)
            const uint32_t idxA = mod_fast(finish_prep<KIND>(D1, K.s1, y, c_rot_digit[y]), K.fm, K.nm);
            const uint32_t idxB = mod_fast(finish_prep<KIND>(D1, K.s1, y + 1u, c_rot_digit[y + 1u]), K.fm, K.nm);
            const uint32_t bA = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idxA);
            const uint32_t bB = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idxB);
            const bool svA = (bA & ~(skip10 >> y) & 1u) != 0u;
            const bool svB = (bB & ~(skip10 >> (y + 1u)) & 1u) != 0u;
#pragma unroll 1
            for (uint32_t j = 0; j < 2u; j++) {                      // push survivors, then run B / C when a batch is ready
                const bool sv = j ? svB : svA;
                const uint32_t b = __ballot_sync(0xffffffffu, sv);
                sts128_if(qb_addr + 16u * ((R.qb_head + R.qb_cnt + __popc(b & lt)) & (Q2_RING - 1)), j ? idxB : idxA,
                          tagx | (y + j), (uint32_t)D2, (uint32_t)(D2 >> 32), sv);
                R.qb_cnt += __popc(b);
                drain_stages<KIND, FKT, PM>(K, sm_addr, sm_addr1, gl, sm_words, qb_addr, qc_addr, pacc_addr, lane, lt, CA, R, false);
            }
        }
    }
#pragma unroll 1
    while (R.qb_cnt | R.qc_cnt)                                      // end of the slab: drain what is left
        drain_stages<KIND, FKT, PM>(K, sm_addr, sm_addr1, gl, sm_words, qb_addr, qc_addr, pacc_addr, lane, lt, CA, R, true);
    __syncwarp();
    uint4 acc = lds128(pacc_addr + 16u * lane);
    sts128_if(pacc_addr + 16u * lane, 0u, 0u, 0u, 0u, true);
    if (active) {
        if (mask != nullptr) {                                       // known members pass (no false negatives); reloaded to save registers
            const Bits128 mb = load_bits100(mask, c, nvalid);
            acc.x |= (uint32_t)mb.lo; acc.y |= (uint32_t)(mb.lo >> 32); acc.z |= (uint32_t)mb.hi; acc.w |= (uint32_t)(mb.hi >> 32);
        }
        pass4[c] = acc;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K3 (dense A+B): stage B's hash is computed speculatively for EVERY position next to stage A's
// (two independent XXH64 chains per position -> ILP, and no A->B ring: at a ~50 % survival rate the
// ring bookkeeping costs more issue slots than the wasted half of the h2 hashes).  Only the ~12 % of
// positions that pass all deterministic probes are compacted into the stage-C ring.
// ------------------------------------------------------------------------------------------
template <int KIND, int FKT, int PM>
__device__ __noinline__ void query_slab_dense(const FilterK K, uint32_t sm_addr, uint32_t sm_addr1,
                                              const uint32_t* __restrict__ gl, uint32_t sm_words,
                                              const uint32_t* __restrict__ mask, uint32_t n, uint32_t slab_c0,
                                              uint32_t c_end, uint4* __restrict__ pass4, uint32_t qc_addr,
                                              uint32_t pacc_addr) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = slab_c0 + lane;
    const bool active = c < c_end;
    const Century cen = make_century(active ? c : slab_c0);
    const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
    const uint32_t nvalid = active ? min(100u, n - 100u * c) : 0u;
    Bits128 mb; mb.lo = 0; mb.hi = 0;
    if (active && mask != nullptr) mb = load_bits100(mask, c, nvalid);
    uint64_t skip_lo = mb.lo, skip_hi = mb.hi;                       // known members and positions beyond n
    if (nvalid < 64u) { skip_hi = ~0ull; skip_lo |= ~((1ull << nvalid) - 1ull); }
    else skip_hi |= ~((1ull << (nvalid - 64u)) - 1ull);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qc_head = 0, qc_cnt = 0;
#pragma unroll 1
    for (uint32_t x = 0; x <= 10u; x++) {                            // x == 10: drain what is left in ring C
        const bool feeding = x < 10u;
        uint64_t D1 = 0, D2 = 0;
        uint32_t skip10 = 0x3ffu;
        if (feeding) {
            D1 = decade_state_t<KIND>(C1, K.s1, x);
            D2 = decade_state_t<KIND>(C2, K.s2, x);
            const uint32_t p0 = 10u * x;
            const uint64_t sh = (p0 < 64u) ? ((skip_lo >> p0) | (p0 ? (skip_hi << (64u - p0)) : 0ull)) : (skip_hi >> (p0 - 64u));
            skip10 = (uint32_t)sh & 0x3ffu;
        }
        const uint32_t tagx = (lane << 8) | (x << 4);
#pragma unroll 1
        for (uint32_t y = 0; y < 10u; y++) {
            if (feeding) {                                           // ---- stages A + B, every position
                const uint32_t idx0 = mod_fast(finish_t<KIND>(D1, K.s1, y), K.fm, K.nm);
                const uint32_t stepm = mod_fast(finish_t<KIND>(D2, K.s2, y), K.fm, K.nm);
                uint32_t ok = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx0) & ~(skip10 >> y) & 1u;
                uint32_t idx = idx0;
                if (FKT > 0) {
#pragma unroll
                    for (int i = 1; i < FKT; i++) {
                        idx = addmod_fast(idx, stepm, K.fm.m);
                        ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
                    }
                } else {
                    for (uint32_t i = 1; i < K.fk; i++) {
                        idx = addmod_fast(idx, stepm, K.fm.m);
                        ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
                    }
                }
                if (K.has_act) {
                    idx = addmod_fast(idx, stepm, K.fm.m);           // index of probe floor_k
                    const uint32_t b2 = __ballot_sync(0xffffffffu, ok != 0u);
                    sts64_if(qc_addr + 8u * ((qc_head + qc_cnt + __popc(b2 & lt)) & (Q2_RING - 1)), idx, tagx | y, ok != 0u);
                    qc_cnt += __popc(b2);
                } else {
                    deliver_pass(pacc_addr, tagx | y, ok != 0u);
                }
            }
            if (qc_cnt >= 32u || (!feeding && qc_cnt > 0u)) {        // ---- stage C: 32 survivors
                __syncwarp();
                const uint32_t nc = min(32u, qc_cnt);
                const bool have = lane < nc;
                const uint2 r = lds64(qc_addr + 8u * ((qc_head + lane) & (Q2_RING - 1)));
                qc_head = (qc_head + nc) & (Q2_RING - 1);
                qc_cnt -= nc;
                const uint32_t tag = have ? r.y : 0u;
                const uint32_t owner = tag >> 8;
                const uint64_t CAo = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)CA, owner) |
                                     ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(CA >> 32), owner) << 32);
                const uint64_t hA = finish_t<KIND>(decade_state_t<KIND>(CAo, K.sA, (tag >> 4) & 15u), K.sA, tag & 15u);
                const uint32_t pb = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, have ? r.x : 0u);
                deliver_pass(pacc_addr, tag, have && (!(hA < K.T) || pb != 0u));
            }
            if (!feeding && qc_cnt == 0u) break;
        }
    }
    __syncwarp();
    uint4 acc = lds128(pacc_addr + 16u * lane);
    sts128_if(pacc_addr + 16u * lane, 0u, 0u, 0u, 0u, true);
    if (active) {
        acc.x |= (uint32_t)mb.lo; acc.y |= (uint32_t)(mb.lo >> 32); acc.z |= (uint32_t)mb.hi; acc.w |= (uint32_t)(mb.hi >> 32);
        pass4[c] = acc;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K3 (decade tiles): compaction without a per-position ring push.  A warp evaluates stage A for a
// whole decade -- ten positions per lane, y a compile-time constant, ten independent XXH64 chains
// per lane (ILP) -- then ONE warp scan of the survivor counts places every survivor in a flat
// per-decade buffer of 4-byte records {idx0:23, lane:5, y:4}.  Stage B consumes the buffer in dense
// batches of 32 and fetches the owner's decade state of seed 2 with a shuffle (all records of the
// buffer belong to the current decade).  Survivors of B go through the small stage-C ring as before.
// Requires m <= 2^23 (4K and 8K frames); larger filters use the ring kernel.
// ------------------------------------------------------------------------------------------
#ifndef RBF_Q3_WARPS
#define RBF_Q3_WARPS 28
#endif
constexpr int Q3_WARPS = RBF_Q3_WARPS, Q3_THREADS = 32 * Q3_WARPS;
#ifndef RBF_Q3_TY
#define RBF_Q3_TY 10
#endif
constexpr int Q3_TY = RBF_Q3_TY;                                        // positions of a decade per lane and tile: 10 or 5
constexpr int Q3_BUF = 32 * Q3_TY;                                       // survivors of one tile of a slab, worst case
constexpr int Q3_WARP_WORDS = Q3_BUF + (Q2_RING * 8 + 32 * 16 + 128) / 4;   // decade buffer, C ring, pass accumulators, digit table

__device__ __forceinline__ void sts32_if(uint32_t addr, uint32_t v, bool p) {
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q st.shared.u32 [%0], %1;\n}" ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

template <int KIND, int PM>
__device__ __forceinline__ void drain_c_ring(const FilterK& K, uint32_t sm_addr, uint32_t sm_addr1, const uint32_t* __restrict__ gl,
                                             uint32_t sm_words, uint32_t qc_addr, uint32_t pacc_addr, uint32_t lane, uint64_t CA,
                                             uint32_t& qc_head, uint32_t& qc_cnt) {
    __syncwarp();
    const uint32_t nc = min(32u, qc_cnt);
    const bool have = lane < nc;
    const uint2 r = lds64(qc_addr + 8u * ((qc_head + lane) & (Q2_RING - 1)));
    qc_head = (qc_head + nc) & (Q2_RING - 1);
    qc_cnt -= nc;
    const uint32_t tag = have ? r.y : 0u;
    const uint32_t owner = (tag >> 8) & 31u;
    const uint64_t CAo = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)CA, owner) |
                         ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(CA >> 32), owner) << 32);
    const uint64_t hA = finish_t<KIND>(decade_state_t<KIND>(CAo, K.sA, (tag >> 4) & 15u), K.sA, tag & 15u);
    const uint32_t pb = probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, have ? r.x : 0u);
    deliver_pass(pacc_addr, tag, have && (!(hA < K.T) || pb != 0u));
}

#ifdef RBF_Q3_INLINE
#define RBF_Q3_FN __forceinline__
#else
#define RBF_Q3_FN __noinline__
#endif
template <int KIND, int FKT, int PM>
__device__ RBF_Q3_FN void query_slab_tiled(const FilterK K, uint32_t sm_addr, uint32_t sm_addr1,
                                              const uint32_t* __restrict__ gl, uint32_t sm_words,
                                              const uint32_t* __restrict__ mask, uint32_t n, uint32_t slab_c0,
                                              uint32_t c_end, uint4* __restrict__ pass4, uint32_t buf_addr) {
    const uint32_t qc_addr = buf_addr + 4u * Q3_BUF, pacc_addr = qc_addr + 8u * Q2_RING, rb_addr = pacc_addr + 512u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = slab_c0 + lane;
    const bool active = c < c_end;
    const Century cen = make_century(active ? c : slab_c0);
    const uint64_t C1 = century_state(cen, K.s1), C2 = century_state(cen, K.s2), CA = century_state(cen, K.sA);
    const uint32_t nvalid = active ? min(100u, n - 100u * c) : 0u;
    uint64_t skip_lo = 0, skip_hi = 0;                               // known members and positions beyond n need no hashing
    if (active && mask != nullptr) { const Bits128 mb = load_bits100(mask, c, nvalid); skip_lo = mb.lo; skip_hi = mb.hi; }
    if (nvalid < 64u) { skip_hi = ~0ull; skip_lo |= ~((1ull << nvalid) - 1ull); }
    else skip_hi |= ~((1ull << (nvalid - 64u)) - 1ull);
    if (lane < 10u) sts64_if(rb_addr + 8u * lane, (uint32_t)c_rot_digit[lane], (uint32_t)(c_rot_digit[lane] >> 32), true);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qc_head = 0, qc_cnt = 0;
#pragma unroll 1
    for (uint32_t x = 0; x < 10u; x++) {
        const uint64_t D1 = decade_prep<KIND>(decade_state_t<KIND>(C1, K.s1, x));
        const uint64_t D2 = decade_prep<KIND>(decade_state_t<KIND>(C2, K.s2, x));
        const uint32_t p0 = 10u * x;
        const uint64_t sh = (p0 < 64u) ? ((skip_lo >> p0) | (p0 ? (skip_hi << (64u - p0)) : 0ull)) : (skip_hi >> (p0 - 64u));
#pragma unroll
        for (int h = 0; h < 10 / Q3_TY; h++) {
        // ---- stage A: Q3_TY positions per lane, y compile-time
        uint32_t idx0[Q3_TY];
        uint32_t sv = 0;
#pragma unroll
        for (int yy = 0; yy < Q3_TY; yy++) {
            const uint32_t y = (uint32_t)(h * Q3_TY + yy);
            idx0[yy] = mod_fast(finish_prep<KIND>(D1, K.s1, y, rot_digit_const(y)), K.fm, K.nm);
            sv |= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx0[yy]) << yy;
        }
        sv &= ~(uint32_t)(sh >> (h * Q3_TY)) & ((1u << Q3_TY) - 1u);
        // ---- one scan per tile places the survivors
        const uint32_t cnt = __popc(sv);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t off = buf_addr + 4u * (inc - cnt);
        const uint32_t ltag = lane << 23;
#pragma unroll
        for (int yy = 0; yy < Q3_TY; yy++) {
            const bool p = ((sv >> yy) & 1u) != 0u;
            sts32_if(off, idx0[yy] | ltag | ((uint32_t)(h * Q3_TY + yy) << 28), p);
            off += p ? 4u : 0u;
        }
        __syncwarp();
        // ---- stage B: dense batches of 32 survivors of this decade
#pragma unroll 1
        for (uint32_t b = 0; b < total; b += 32u) {
            const uint32_t g = b + lane;
            const bool have = g < total;
            const uint32_t rec = lds32(buf_addr + 4u * min(g, (uint32_t)(Q3_BUF - 1)));
            const uint32_t owner = (rec >> 23) & 31u, y = have ? (rec >> 28) : 0u;
            const uint64_t D2o = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)D2, owner) |
                                 ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(D2 >> 32), owner) << 32);
            uint64_t rb = 0;
            if (kind_ends_in_byte<KIND>()) { const uint2 t = lds64(rb_addr + 8u * y); rb = (uint64_t)t.x | ((uint64_t)t.y << 32); }
            const uint32_t stepm = have ? mod_fast(finish_prep<KIND>(D2o, K.s2, y, rb), K.fm, K.nm) : 0u;
            uint32_t idx = have ? (rec & 0x7fffffu) : 0u;
            uint32_t ok = have ? 1u : 0u;
            if (FKT > 0) {
#pragma unroll
                for (int i = 1; i < FKT; i++) {
                    idx = addmod_fast(idx, stepm, K.fm.m);
                    ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
                }
            } else {
                for (uint32_t i = 1; i < K.fk; i++) {
                    idx = addmod_fast(idx, stepm, K.fm.m);
                    ok &= probe_bit<PM>(sm_addr, sm_addr1, gl, sm_words, idx);
                }
            }
            const uint32_t tag = (owner << 8) | (x << 4) | y;
            if (K.has_act) {
                idx = addmod_fast(idx, stepm, K.fm.m);               // index of probe floor_k
                const uint32_t b2 = __ballot_sync(0xffffffffu, ok != 0u);
                sts64_if(qc_addr + 8u * ((qc_head + qc_cnt + __popc(b2 & lt)) & (Q2_RING - 1)), idx, tag, ok != 0u);
                qc_cnt += __popc(b2);
                if (qc_cnt >= 32u) drain_c_ring<KIND, PM>(K, sm_addr, sm_addr1, gl, sm_words, qc_addr, pacc_addr, lane, CA, qc_head, qc_cnt);
            } else {
                deliver_pass(pacc_addr, tag, ok != 0u);
            }
        }
        __syncwarp();                                                // the tile buffer is rewritten next
        }
    }
#pragma unroll 1
    while (qc_cnt) drain_c_ring<KIND, PM>(K, sm_addr, sm_addr1, gl, sm_words, qc_addr, pacc_addr, lane, CA, qc_head, qc_cnt);
    __syncwarp();
    uint4 acc = lds128(pacc_addr + 16u * lane);
    sts128_if(pacc_addr + 16u * lane, 0u, 0u, 0u, 0u, true);
    if (active) {
        if (mask != nullptr) {                                       // known members pass (no false negatives)
            const Bits128 mb = load_bits100(mask, c, nvalid);
            acc.x |= (uint32_t)mb.lo; acc.y |= (uint32_t)(mb.lo >> 32); acc.z |= (uint32_t)mb.hi; acc.w |= (uint32_t)(mb.hi >> 32);
        }
        pass4[c] = acc;
    }
    __syncwarp();
}

template <int PM>
__global__ void __launch_bounds__(Q3_THREADS, 1) k_query3(const FrameJob* __restrict__ jobs, const uint32_t* __restrict__ cent_prefix,
                                                          int F, uint32_t smem_words_cap) {
    extern __shared__ __align__(128) uint32_t dyn[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;   // warp-uniform for the compiler
    const uint32_t buf = smem_u32(dyn + warp * Q3_WARP_WORDS);
    const uint32_t pacc = buf + 4u * Q3_BUF + 8u * Q2_RING;
    uint32_t* sbits = dyn + Q3_WARPS * Q3_WARP_WORDS;
    const uint32_t sb_addr = smem_u32(sbits);
    sts128_if(pacc + 16u * lane, 0u, 0u, 0u, 0u, true);
    const uint32_t total = cent_prefix[F];
    const uint32_t lo = (uint32_t)(((uint64_t)total * blockIdx.x) / gridDim.x);
    const uint32_t hi = (uint32_t)(((uint64_t)total * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (lo >= hi) return;
    int f = 0;
    { int a = 0, b = F; while (a < b) { int m = (a + b) >> 1; if (cent_prefix[m + 1] > lo) b = m; else a = m + 1; } f = a; }
    uint32_t parity = 0, g = lo;
    while (g < hi) {
        while (cent_prefix[f + 1] <= g) f++;
        const FrameJob& J = jobs[f];
        const uint32_t seg_end = min(hi, cent_prefix[f + 1]);
        const uint32_t c_begin = g - cent_prefix[f], c_end = seg_end - cent_prefix[f];
        const uint32_t nwords = (J.l + 31u) >> 5;
        const uint32_t sw = min((nwords + 3u) & ~3u, smem_words_cap);
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, sw * 4u);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(J.bits);
            uint8_t* dst = reinterpret_cast<uint8_t*>(sbits);
            for (uint32_t off = 0; off < sw * 4u; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, sw * 4u - off), &bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const FilterK K = filter_consts(J);
        uint4* pass4 = reinterpret_cast<uint4*>(J.pass);
        for (uint32_t slab = c_begin + 32u * warp; slab < c_end; slab += 32u * Q3_WARPS) {
            const uint32_t last = min(slab + 31u, c_end - 1u);
            const bool uniform = slab >= 1u && ndigits_u32(slab) == ndigits_u32(last) && K.fk >= 1u && K.fm.fast && K.fm.m <= (1u << 23);
            if (uniform) {
#define RBF_TILED(KD)                                                                                                                   \
    if (K.fk == 3u) query_slab_tiled<KD, 3, PM>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, buf);                       \
    else if (K.fk == 2u) query_slab_tiled<KD, 2, PM>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, buf);                  \
    else query_slab_tiled<KD, 0, PM>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, buf);
                switch (make_century(slab).kind) {
                case K_4B: { RBF_TILED(K_4B) } break;
                case K_8B: { RBF_TILED(K_8B) } break;
                case K_44: { RBF_TILED(K_44) } break;
                case K_88: { RBF_TILED(K_88) } break;
                default:   { RBF_TILED(K_BB) } break;
                }
#undef RBF_TILED
            } else {                                    // century 0, a digit-count boundary, floor_k == 0 or a huge filter
                const uint32_t c = slab + lane;
                if (c < c_end) {
                    BitView bv; bv.sm = sbits; bv.gl = J.bits; bv.sm_words = sw;
                    const Bits128 r = query_century(bv, K, c, min(100u, J.n - 100u * c));
                    pass4[c] = make_uint4((uint32_t)r.lo, (uint32_t)(r.lo >> 32), (uint32_t)r.hi, (uint32_t)(r.hi >> 32));
                }
            }
        }
        g = seg_end;
    }
}

template <int KIND, int PM, int ALG = 0>
__device__ __forceinline__ void query_slab_dispatch(const FilterK& K, uint32_t sm_addr, uint32_t sm_addr1, const uint32_t* __restrict__ gl,
                                                    uint32_t sm_words, const uint32_t* __restrict__ mask, uint32_t n,
                                                    uint32_t slab_c0, uint32_t c_end, uint4* __restrict__ pass4, uint32_t qb,
                                                    uint32_t qc, uint32_t pacc) {
    if (ALG == 1) {                                       // dense A+B, ring only before stage C
        if (K.fk == 3u) query_slab_dense<KIND, 3, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qc, pacc);
        else if (K.fk == 2u) query_slab_dense<KIND, 2, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qc, pacc);
        else query_slab_dense<KIND, 0, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qc, pacc);
    } else {
        if (K.fk == 3u) query_slab_staged<KIND, 3, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qb, qc, pacc);
        else if (K.fk == 2u) query_slab_staged<KIND, 2, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qb, qc, pacc);
        else query_slab_staged<KIND, 0, PM>(K, sm_addr, sm_addr1, gl, sm_words, mask, n, slab_c0, c_end, pass4, qb, qc, pacc);
    }
}

template <bool HYBRID, int ALG>
__global__ void __launch_bounds__(Q2_THREADS, 1) k_query2(const FrameJob* __restrict__ jobs,
                                                          const uint32_t* __restrict__ cent_prefix, int F,
                                                          uint32_t smem_words_cap) {
    extern __shared__ __align__(128) uint32_t dyn[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;   // warp-uniform for the compiler
    uint32_t* wq = dyn + warp * Q2_WARP_WORDS;
    const uint32_t qb = smem_u32(wq), qc = smem_u32(wq + Q2_RING * 4), pacc = smem_u32(wq + Q2_RING * 4 + Q2_RING * 2);
    uint32_t* sbits = dyn + Q2_WARPS * Q2_WARP_WORDS;
    const uint32_t sb_addr = smem_u32(sbits);
    sts128_if(pacc + 16u * lane, 0u, 0u, 0u, 0u, true);
    const uint32_t total = cent_prefix[F];
    const uint32_t lo = (uint32_t)(((uint64_t)total * blockIdx.x) / gridDim.x);
    const uint32_t hi = (uint32_t)(((uint64_t)total * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (lo >= hi) return;
    int f = 0;
    { int a = 0, b = F; while (a < b) { int m = (a + b) >> 1; if (cent_prefix[m + 1] > lo) b = m; else a = m + 1; } f = a; }
    uint32_t parity = 0, g = lo;
    while (g < hi) {
        while (cent_prefix[f + 1] <= g) f++;
        const FrameJob& J = jobs[f];
        const uint32_t seg_end = min(hi, cent_prefix[f + 1]);
        const uint32_t c_begin = g - cent_prefix[f], c_end = seg_end - cent_prefix[f];
        const uint32_t nwords = (J.l + 31u) >> 5;
        const uint32_t sw = min((nwords + 3u) & ~3u, smem_words_cap);
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, sw * 4u);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(J.bits);
            uint8_t* dst = reinterpret_cast<uint8_t*>(sbits);
            for (uint32_t off = 0; off < sw * 4u; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, sw * 4u - off), &bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const FilterK K = filter_consts(J);
        uint4* pass4 = reinterpret_cast<uint4*>(J.pass);
        for (uint32_t slab = c_begin + 32u * warp; slab < c_end; slab += 32u * Q2_WARPS) {
            const uint32_t last = min(slab + 31u, c_end - 1u);
            const bool uniform = slab >= 1u && ndigits_u32(slab) == ndigits_u32(last) && K.fk >= 1u && K.fm.fast;
            if (uniform) {
                switch (make_century(slab).kind) {
                case K_4B: query_slab_dispatch<K_4B, (HYBRID ? 1 : 0), ALG>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_8B: query_slab_dispatch<K_8B, (HYBRID ? 1 : 0), ALG>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_44: query_slab_dispatch<K_44, (HYBRID ? 1 : 0), ALG>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_88: query_slab_dispatch<K_88, (HYBRID ? 1 : 0), ALG>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                default:   query_slab_dispatch<K_BB, (HYBRID ? 1 : 0), ALG>(K, sb_addr, 0u, J.bits, sw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                }
            } else {                                    // century 0, a digit-count boundary, or floor_k == 0
                const uint32_t c = slab + lane;
                if (c < c_end) {
                    BitView bv; bv.sm = sbits; bv.gl = J.bits; bv.sm_words = sw;
                    const Bits128 r = query_century(bv, K, c, min(100u, J.n - 100u * c));
                    pass4[c] = make_uint4((uint32_t)r.lo, (uint32_t)(r.lo >> 32), (uint32_t)r.hi, (uint32_t)(r.hi >> 32));
                }
            }
        }
        g = seg_end;
    }
}

// ------------------------------------------------------------------------------------------
// K3 (cluster): the same staged query for Bloom arrays that do not fit one CTA's shared memory.
// A cluster of two CTAs (two SMs) shares one frame: each CTA stages HALF of the bit array with TMA
// into its own shared memory and probes the other half through distributed shared memory
// (ld.shared::cluster), so no probe goes to L2.  The two CTAs interleave the slabs of the cluster's
// century range; two cluster barriers per frame segment order the re-staging.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q2_THREADS, 1)
k_query2c(const FrameJob* __restrict__ jobs, const uint32_t* __restrict__ cent_prefix, int F, uint32_t half_words_cap) {
    extern __shared__ __align__(128) uint32_t dyn[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;   // warp-uniform for the compiler
    const uint32_t rank = cluster_ctarank();
    const uint32_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    uint32_t* wq = dyn + warp * Q2_WARP_WORDS;
    const uint32_t qb = smem_u32(wq), qc = smem_u32(wq + Q2_RING * 4), pacc = smem_u32(wq + Q2_RING * 4 + Q2_RING * 2);
    uint32_t* sbits = dyn + Q2_WARPS * Q2_WARP_WORDS;
    const uint32_t base0 = mapa_shared(smem_u32(sbits), 0u), base1 = mapa_shared(smem_u32(sbits), 1u);
    sts128_if(pacc + 16u * lane, 0u, 0u, 0u, 0u, true);
    const uint32_t total = cent_prefix[F];
    const uint32_t lo = (uint32_t)(((uint64_t)total * cid) / ncl);
    const uint32_t hi = (uint32_t)(((uint64_t)total * (cid + 1)) / ncl);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    int f = 0;
    { int a = 0, b = F; while (a < b) { int m = (a + b) >> 1; if (cent_prefix[m + 1] > lo) b = m; else a = m + 1; } f = a; }
    uint32_t parity = 0, g = lo;
    while (g < hi) {                                                  // both CTAs of the cluster walk the same segments
        while (cent_prefix[f + 1] <= g) f++;
        const FrameJob& J = jobs[f];
        const uint32_t seg_end = min(hi, cent_prefix[f + 1]);
        const uint32_t c_begin = g - cent_prefix[f], c_end = seg_end - cent_prefix[f];
        const uint32_t nwords = ((J.l + 31u) >> 5);
        const uint32_t hw = min(((((nwords + 1u) >> 1) + 3u) & ~3u), half_words_cap);   // words held by rank 0
        const uint32_t mine_begin = rank ? hw : 0u;
        const uint32_t mine_words = rank ? ((nwords > hw ? nwords - hw : 0u) + 3u) & ~3u : hw;
        cluster_sync_all();                                           // nobody still probes the previous array
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, mine_words * 4u);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(J.bits + mine_begin);
            uint8_t* dst = reinterpret_cast<uint8_t*>(sbits);
            for (uint32_t off = 0; off < mine_words * 4u; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, mine_words * 4u - off), &bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        cluster_sync_all();                                           // both halves are in place
        const FilterK K = filter_consts(J);
        uint4* pass4 = reinterpret_cast<uint4*>(J.pass);
        const uint32_t a1 = base1 - 4u * hw;
        for (uint32_t slab = c_begin + 32u * (warp + Q2_WARPS * rank); slab < c_end; slab += 64u * Q2_WARPS) {
            const uint32_t last = min(slab + 31u, c_end - 1u);
            const bool uniform = slab >= 1u && ndigits_u32(slab) == ndigits_u32(last) && K.fk >= 1u && K.fm.fast;
            if (uniform) {
                switch (make_century(slab).kind) {
                case K_4B: query_slab_dispatch<K_4B, 2>(K, base0, a1, J.bits, hw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_8B: query_slab_dispatch<K_8B, 2>(K, base0, a1, J.bits, hw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_44: query_slab_dispatch<K_44, 2>(K, base0, a1, J.bits, hw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                case K_88: query_slab_dispatch<K_88, 2>(K, base0, a1, J.bits, hw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                default:   query_slab_dispatch<K_BB, 2>(K, base0, a1, J.bits, hw, J.mask, J.n, slab, c_end, pass4, qb, qc, pacc); break;
                }
            } else {                                    // rare slabs: probe the global copy
                const uint32_t c = slab + lane;
                if (c < c_end) {
                    BitView bv; bv.sm = sbits; bv.gl = J.bits; bv.sm_words = 0u;
                    const Bits128 r = query_century(bv, K, c, min(100u, J.n - 100u * c));
                    pass4[c] = make_uint4((uint32_t)r.lo, (uint32_t)(r.lo >> 32), (uint32_t)r.hi, (uint32_t)(r.hi >> 32));
                }
            }
        }
        g = seg_end;
    }
    cluster_sync_all();                                               // a peer may still be reading this CTA's half
}

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan helper (1024 threads)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? s_warp[lane] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        s_warp[lane] = wi - w;                          // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;                // block total
    }
    __syncthreads();
    const uint32_t res = s_warp[warp] + inc - v;
    block_total = s_warp[32];
    __syncthreads();
    return res;
}

// ------------------------------------------------------------------------------------------
// K3b: witness.  One CTA per frame walks the centuries in order; witness = mask bits at the
// passing positions (ivc:253), concatenated.  Finishes by converting witness and bit array
// to np.packbits order in place (ivc:945, ivc:951).
// ------------------------------------------------------------------------------------------
// pass counts per (frame, chunk of centuries): lets several CTAs work on one frame, each knowing how many passing
// positions precede its chunk
__global__ void __launch_bounds__(256) k_pass_count(const FrameJob* __restrict__ jobs, uint32_t chunks, uint32_t* __restrict__ counts) {
    const FrameJob& J = jobs[blockIdx.y];
    __shared__ uint32_t s_c[8];
    uint32_t c = 0;
    if (J.l != 0) {
        const uint32_t ncent = (J.n + 99u) / 100u;
        const uint32_t c0 = (uint32_t)(((uint64_t)ncent * blockIdx.x) / chunks), c1 = (uint32_t)(((uint64_t)ncent * (blockIdx.x + 1)) / chunks);
        const uint4* pass4 = reinterpret_cast<const uint4*>(J.pass);
        for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const uint4 p = pass4[i];
            c += __popc(p.x) + __popc(p.y) + __popc(p.z) + __popc(p.w);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s_c[i];
        counts[blockIdx.y * chunks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_witness(const FrameJob* __restrict__ jobs, uint32_t chunks,
                                                  const uint32_t* __restrict__ counts, uint32_t* __restrict__ wlen_out) {
    const FrameJob& J = jobs[blockIdx.y];
    __shared__ uint32_t s_warp[33];
    if (J.l == 0) { if (threadIdx.x == 0 && blockIdx.x == 0) wlen_out[blockIdx.y] = 0; return; }
    const uint32_t ncent = (J.n + 99u) / 100u;
    const uint32_t c0 = (uint32_t)(((uint64_t)ncent * blockIdx.x) / chunks), c1 = (uint32_t)(((uint64_t)ncent * (blockIdx.x + 1)) / chunks);
    const uint4* pass4 = reinterpret_cast<const uint4*>(J.pass);
    uint32_t base = 0;
    for (uint32_t i = 0; i < blockIdx.x; i++) base += counts[blockIdx.y * chunks + i];   // passes before this chunk
    for (uint32_t r0 = c0; r0 < c1; r0 += blockDim.x) {
        const uint32_t c = r0 + threadIdx.x;
        uint64_t wlo = 0, whi = 0;
        uint32_t cnt = 0;
        if (c < c1) {
            const uint4 p = pass4[c];
            const Bits128 mb = load_bits100(J.mask, c, min(100u, J.n - 100u * c));
            const uint32_t P[4] = {p.x, p.y, p.z, p.w};
            const uint32_t M[4] = {(uint32_t)mb.lo, (uint32_t)(mb.lo >> 32), (uint32_t)mb.hi, (uint32_t)(mb.hi >> 32)};
            // witness bit k = mask bit of the k-th passing position: walk the (few) members, not the passes --
            // a member's k is its rank among the passing positions
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t mw = M[j] & P[j];
                while (mw) {
                    const uint32_t b = __ffs(mw) - 1;
                    mw &= mw - 1u;
                    const uint32_t k = cnt + __popc(P[j] & ((1u << b) - 1u));
                    if (k < 64u) wlo |= 1ull << k; else whi |= 1ull << (k - 64u);
                }
                cnt += __popc(P[j]);
            }
        }
        uint32_t tot;
        const uint32_t off = block_excl_scan(cnt, s_warp, tot);
        if (wlo | whi) or_bits128(J.witness, (uint64_t)base + off, wlo, whi);
        base += tot;
    }
    if (threadIdx.x == 0 && blockIdx.x == chunks - 1) wlen_out[blockIdx.y] = base;
}

// witness and bit array: LSB-first words -> np.packbits order, in place (ivc:945, ivc:951)
__global__ void __launch_bounds__(256) k_finalize(const FrameJob* __restrict__ jobs, const uint32_t* __restrict__ wlen) {
    const FrameJob& J = jobs[blockIdx.y];
    if (J.l == 0) return;
    const uint32_t wwords = (wlen[blockIdx.y] + 31u) >> 5, bwords = (J.l + 31u) >> 5;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < wwords; i += gridDim.x * blockDim.x) J.witness[i] = bitrev_bytes(J.witness[i]);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < bwords; i += gridDim.x * blockDim.x) J.bits[i] = bitrev_bytes(J.bits[i]);
}

// ------------------------------------------------------------------------------------------
// K4b: decode expand.  out[i] = witness[rank of i among passing positions] (ivc:299-304).
// Witness here is LSB-first (the host converts the packbits input once with k_bitrev).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_expand(const FrameJob* __restrict__ jobs, uint32_t chunks, const uint32_t* __restrict__ counts,
                                                 uint32_t* __restrict__ consumed) {
    const FrameJob& J = jobs[blockIdx.y];
    __shared__ uint32_t s_warp[33];
    if (J.l == 0) { if (threadIdx.x == 0 && blockIdx.x == 0) consumed[blockIdx.y] = 0; return; }
    const uint32_t ncent = (J.n + 99u) / 100u;
    const uint32_t c0 = (uint32_t)(((uint64_t)ncent * blockIdx.x) / chunks), c1 = (uint32_t)(((uint64_t)ncent * (blockIdx.x + 1)) / chunks);
    const uint4* pass4 = reinterpret_cast<const uint4*>(J.pass);
    uint32_t base = 0;
    for (uint32_t i = 0; i < blockIdx.x; i++) base += counts[blockIdx.y * chunks + i];
    for (uint32_t r0 = c0; r0 < c1; r0 += blockDim.x) {
        const uint32_t c = r0 + threadIdx.x;
        uint4 p = make_uint4(0, 0, 0, 0);
        uint32_t cnt = 0;
        if (c < c1) { p = pass4[c]; cnt = __popc(p.x) + __popc(p.y) + __popc(p.z) + __popc(p.w); }
        uint32_t tot;
        const uint32_t off = base + block_excl_scan(cnt, s_warp, tot);
        if (cnt) {
            // fetch cnt (<=100) witness bits starting at bit `off`; bits at or beyond wlen_in read as 0
            const uint32_t w = off >> 5, sh = off & 31u;
            const uint32_t lim = (J.wlen_in + 31u) >> 5;
            uint32_t a[5];
#pragma unroll
            for (int j = 0; j < 5; j++) a[j] = (w + j < lim) ? __ldg(J.witness + w + j) : 0u;
            uint32_t s[4];
#pragma unroll
            for (int j = 0; j < 4; j++) s[j] = __funnelshift_r(a[j], a[j + 1], sh);
            uint64_t wl = (uint64_t)s[0] | ((uint64_t)s[1] << 32), wh = (uint64_t)s[2] | ((uint64_t)s[3] << 32);
            // bits beyond wlen_in are zero by construction of the padded buffer tail (host zero-fills)
            const uint32_t P[4] = {p.x, p.y, p.z, p.w};
            uint32_t O[4] = {0, 0, 0, 0};
            uint32_t k = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t pw = P[j];
                while (pw) {
                    const uint32_t b = __ffs(pw) - 1;
                    pw &= pw - 1u;
                    const uint32_t bit = (uint32_t)(((k < 64u) ? (wl >> k) : (wh >> (k - 64u))) & 1ull);
                    const uint32_t valid = (off + k < J.wlen_in) ? 1u : 0u;
                    O[j] |= (bit & valid) << b;
                    k++;
                }
            }
            or_bits128(J.out_mask, 100ull * c, (uint64_t)O[0] | ((uint64_t)O[1] << 32), (uint64_t)O[2] | ((uint64_t)O[3] << 32));
        }
        base += tot;
    }
    if (threadIdx.x == 0 && blockIdx.x == chunks - 1) consumed[blockIdx.y] = base;
}

// ------------------------------------------------------------------------------------------
// N1 / N2: ordered gather of the changed pixels' values (ivc:810-842) and the scatter
// that rebuilds the next frame (ivc:849-909).  One CTA per pair walks the mask words in order;
// a block scan of the popcounts gives every set pixel its rank.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_gather_scatter(const GatherJob* __restrict__ jobs, int scatter,
                                                          uint32_t* __restrict__ counts) {
    const GatherJob J = jobs[blockIdx.x];
    __shared__ uint32_t s_warp[33];
    const uint32_t nwords = (J.npix + 31u) >> 5;
    const uint32_t pb = J.pix_bytes;
    uint32_t base = 0;
    for (uint32_t w0 = 0; w0 < nwords; w0 += blockDim.x) {
        const uint32_t w = w0 + threadIdx.x;
        uint32_t m = (w < nwords) ? __ldg(J.mask + w) : 0u;
        uint32_t tot;
        uint32_t rank = base + block_excl_scan(__popc(m), s_warp, tot);
        while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1u;
            const size_t px = ((size_t)w << 5) + b;
            if (scatter) {
                for (uint32_t q = 0; q < pb; q++) J.out_frame[px * pb + q] = J.values[(size_t)rank * pb + q];
            } else {
                for (uint32_t q = 0; q < pb; q++) J.values[(size_t)rank * pb + q] = J.frame[px * pb + q];
            }
            rank++;
        }
        base += tot;
    }
    if (threadIdx.x == 0 && counts) counts[blockIdx.x] = base;
}

// ------------------------------------------------------------------------------------------
// N3: the 5x5 median of cv2.medianBlur (ivc:738) -- replicated border, 13th smallest of
// the 25 samples -- on channel 0 of an interleaved frame.  Rank selection (25 x 25 comparisons) is
// exact for any sample type; the float32 std that follows (ivc:741-744) stays in numpy on the host so
// that the noise estimate is bit-identical to the reference's.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_median5(const T* __restrict__ in, uint32_t pix_stride, uint32_t H, uint32_t W,
                                                  T* __restrict__ out) {
    const uint32_t x = blockIdx.x * 32u + (threadIdx.x & 31u), y = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    uint32_t v[25];
#pragma unroll
    for (int dy = -2; dy <= 2; dy++) {
        const uint32_t yy = (uint32_t)min(max((int)y + dy, 0), (int)H - 1);
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) {
            const uint32_t xx = (uint32_t)min(max((int)x + dx, 0), (int)W - 1);
            v[(dy + 2) * 5 + (dx + 2)] = (uint32_t)__ldg(in + ((size_t)yy * W + xx) * pix_stride);
        }
    }
    uint32_t med = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) {
        uint32_t rank = 0;
#pragma unroll
        for (int j = 0; j < 25; j++) rank += (v[j] < v[i] || (v[j] == v[i] && j < i)) ? 1u : 0u;
        if (rank == 12u) med = v[i];
    }
    out[(size_t)y * W + x] = (T)med;
}

// ------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------
__global__ void k_bitrev(uint32_t* __restrict__ w, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        w[i] = bitrev_bytes(w[i]);
}
// LSB-first packed bits -> one byte per bit (np.uint8 0/1)
__global__ void k_unpack_bits(const uint32_t* __restrict__ w, uint8_t* __restrict__ out, size_t nbits) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nbits; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)((w[i >> 5] >> (i & 31)) & 1u);
}
// one byte per position (== 1 is set, as `binary_input[i] == 1`, ivc:236) -> LSB-first packed words
__global__ void k_pack_bytes(const uint8_t* __restrict__ in, uint32_t* __restrict__ w, size_t nbits) {
    const size_t nwords = (nbits + 31) >> 5;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        const size_t b0 = i << 5;
        for (uint32_t k = 0; k < 32u && b0 + k < nbits; k++) v |= (in[b0 + k] == 1 ? 1u : 0u) << k;
        w[i] = v;
    }
}


// MSB-first (np.packbits order) packed bits -> one byte per bit
__global__ void k_unpack_bits_msb(const uint32_t* __restrict__ w, uint8_t* __restrict__ out, size_t nbits) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nbits; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)((w[i >> 5] >> ((i & 31) ^ 7)) & 1u);
}
__global__ void k_popcount(const uint32_t* __restrict__ w, size_t n, uint32_t* __restrict__ out) {
    uint32_t c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += __popc(w[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
// per frame: number of 32-bit words in which two packed bit arrays differ
__global__ void k_count_diff(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t stride_words,
                             size_t nwords, uint32_t* __restrict__ out) {
    const uint32_t* pa = a + (size_t)blockIdx.y * stride_words;
    const uint32_t* pb = b + (size_t)blockIdx.y * stride_words;
    uint32_t c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
        c += (pa[i] != pb[i]) ? 1u : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out + blockIdx.y, c);
}

// ------------------------------------------------------------------------------------------
// explicit-item kernels: RationalBloomFilter.add_index / check_index on a list of indices
// (ivc:99-138), and the string-keyed twin rbf.RationalBloomFilter / StandardBloomFilter
// (rbf:25-41, rbf:103-182).  Bit array is LSB-first in global memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t check_hashes_global(const uint32_t* __restrict__ bits, const FilterK& K, uint64_t h1,
                                                        uint64_t h2, uint64_t hA) {
    uint32_t idx = mod_u64(h1, K.fm);
    const uint32_t step = mod_u64(h2, K.fm);
    for (uint32_t i = 0; i < K.fk; i++) {
        if (!((bits[idx >> 5] >> (idx & 31u)) & 1u)) return 0u;
        idx = addmod(idx, step, K.fm.m);
    }
    if (K.has_act && hA < K.T) { if (!((bits[idx >> 5] >> (idx & 31u)) & 1u)) return 0u; }
    return 1u;
}

__global__ void k_items_u32(const FrameJob* __restrict__ job, const uint32_t* __restrict__ items, uint32_t count,
                            uint8_t* __restrict__ result, int insert) {
    const FrameJob& J = *job;
    const FilterK K = filter_consts(J);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint32_t i = items[t];
        const Century cen = make_century(i / 100u);
        const uint32_t x = (i / 10u) % 10u, y = i % 10u;
        const uint64_t h1 = finish(cen.kind, decade_state(cen, century_state(cen, K.s1), K.s1, x), K.s1, y);
        const uint64_t h2 = finish(cen.kind, decade_state(cen, century_state(cen, K.s2), K.s2, x), K.s2, y);
        const uint64_t hA = K.has_act ? finish(cen.kind, decade_state(cen, century_state(cen, K.sA), K.sA, x), K.sA, y) : 0ull;
        if (insert) insert_hashes(J.bits, K, h1, h2, hA);
        else result[t] = (uint8_t)check_hashes_global(J.bits, K, h1, h2, hA);
    }
}

__global__ void k_items_str(const FrameJob* __restrict__ job, const uint8_t* __restrict__ blob,
                            const uint64_t* __restrict__ offs, uint32_t count, uint8_t* __restrict__ result, int insert,
                            int standard_k) {
    const FrameJob& J = *job;
    const FilterK K = filter_consts(J);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint8_t* s = blob + offs[t];
        const uint32_t len = (uint32_t)(offs[t + 1] - offs[t]);
        if (standard_k > 0) {           // StandardBloomFilter: k independent hashes, seed = i (rbf:25-41)
            uint32_t ok = 1u;
            for (int i = 0; i < standard_k; i++) {
                const uint32_t idx = mod_u64(xxh64_bytes(s, len, (uint64_t)i), K.fm);
                if (insert) atomicOr(J.bits + (idx >> 5), 1u << (idx & 31u));
                else if (!((J.bits[idx >> 5] >> (idx & 31u)) & 1u)) { ok = 0u; break; }
            }
            if (!insert) result[t] = (uint8_t)ok;
        } else {
            const uint64_t h1 = xxh64_bytes(s, len, K.s1), h2 = xxh64_bytes(s, len, K.s2);
            const uint64_t hA = K.has_act ? xxh64_bytes(s, len, K.sA) : 0ull;
            if (insert) insert_hashes(J.bits, K, h1, h2, hA);
            else result[t] = (uint8_t)check_hashes_global(J.bits, K, h1, h2, hA);
        }
    }
}

// KAT / debug: mode 0 = xxh64_decimal(item), mode 1 = century/decade/finish route
__global__ void k_hash_debug(const uint32_t* __restrict__ items, uint32_t count, uint64_t seed, uint64_t* __restrict__ out,
                             int mode) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint32_t i = items[t];
        if (mode == 0) out[t] = xxh64_decimal(i, seed);
        else {
            const Century cen = make_century(i / 100u);
            out[t] = finish(cen.kind, decade_state(cen, century_state(cen, seed), seed, (i / 10u) % 10u), seed, i % 10u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int PB, int S>
static cudaError_t launch_threshold_t(const PairJob* d_pairs, int F, uint32_t npix, int thr, int any_mode, uint32_t* d_ones,
                                      uint32_t* d_resid, int variant, int sm_count, cudaStream_t st) {
    const uint32_t nwords = (npix + 31u) >> 5;
    constexpr uint32_t TP = TMA_TILE_BYTES / PB;
    if (variant == 1 && npix >= TP && (TP / 32u / (TMA_THREADS / 32)) <= 32u) {
        const int smem = 2 * TMA_STAGES * TMA_TILE_BYTES;
        cudaError_t e = cudaFuncSetAttribute(k_threshold_tma<PB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const uint64_t total = (uint64_t)(npix / TP) * (uint64_t)F;
        int grid = (int)((total < (uint64_t)(2 * sm_count)) ? total : (uint64_t)(2 * sm_count));
        k_threshold_tma<PB, S><<<grid, TMA_THREADS, smem, st>>>(d_pairs, F, npix, thr, any_mode, d_ones, d_resid);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const uint32_t px_begin = (npix / TP) * TP;
        if (px_begin < npix) {
            const uint32_t tail_words = nwords - (px_begin >> 5);
            dim3 g((tail_words + 255u) / 256u, (unsigned)F);
            k_threshold_tail<PB, S><<<g, 256, 0, st>>>(d_pairs, npix, px_begin, thr, any_mode, d_ones, d_resid);
        }
        return cudaGetLastError();
    }
    uint32_t bx = (nwords + 255u) / 256u;
    const uint32_t cap = (uint32_t)(sm_count * 32);
    if ((uint64_t)bx * (uint64_t)F > cap) { bx = (cap + (uint32_t)F - 1u) / (uint32_t)F; if (bx < 1u) bx = 1u; }
    dim3 grid(bx, (unsigned)F);
    k_threshold<PB, S><<<grid, 256, 0, st>>>(d_pairs, npix, thr, any_mode, d_ones, d_resid);
    return cudaGetLastError();
}

cudaError_t launch_threshold(const PairJob* d_pairs, int F, uint32_t npix, int channels, int sample_bytes, int thr_int,
                             int any_mode, uint32_t* d_ones, uint32_t* d_resid, int variant, int sm_count, cudaStream_t st) {
    if (F <= 0 || npix == 0) return cudaSuccess;
    const int pb = channels * sample_bytes;
    if (pb == 3 && sample_bytes == 1) return launch_threshold_t<3, 1>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, st);
    if (pb == 6 && sample_bytes == 2) return launch_threshold_t<6, 2>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, st);
    if (pb == 1 && sample_bytes == 1) return launch_threshold_t<1, 1>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, st);
    if (pb == 2 && sample_bytes == 2) return launch_threshold_t<2, 2>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_insert(const FrameJob* d_jobs, int F, uint32_t max_centuries, int variant, int sm_count, cudaStream_t st) {
    if (F <= 0 || max_centuries == 0) return cudaSuccess;
    if (variant == 1) {                                   // dense, warp-compacted insert
        const uint32_t nslab = (max_centuries + 31u) / 32u;
        uint32_t bx = (nslab + I2_WARPS - 1) / I2_WARPS;
        const uint32_t cap = (uint32_t)(sm_count * 12);
        if ((uint64_t)bx * (uint64_t)F > cap) { bx = (cap + (uint32_t)F - 1u) / (uint32_t)F; if (bx < 1u) bx = 1u; }
        dim3 grid(bx, (unsigned)F);
        k_insert2<<<grid, I2_WARPS * 32, 0, st>>>(d_jobs);
        return cudaGetLastError();
    }
    uint32_t bx = (max_centuries + 255u) / 256u;
    const uint32_t cap = (uint32_t)(sm_count * 16);
    if ((uint64_t)bx * (uint64_t)F > cap) { bx = (cap + (uint32_t)F - 1u) / (uint32_t)F; if (bx < 1u) bx = 1u; }
    dim3 grid(bx, (unsigned)F);
    k_insert<<<grid, 256, 0, st>>>(d_jobs);
    return cudaGetLastError();
}


int query_max_smem_bytes() { return 232448 - 1024; }    // 227 KB opt-in minus static shared memory + slack

template <bool HYBRID, int ALG>
static cudaError_t launch_query2_t(const FrameJob* d_jobs, const uint32_t* d_cent_prefix, int F, uint32_t total_centuries,
                                   int sm_count, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_query2<HYBRID, ALG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    uint32_t grid = (uint32_t)sm_count;
    const uint32_t max_useful = (total_centuries + Q2_THREADS - 1) / Q2_THREADS;
    if (grid > max_useful) grid = max_useful;
    if (grid < 1u) grid = 1u;
    k_query2<HYBRID, ALG><<<grid, Q2_THREADS, smem, st>>>(d_jobs, d_cent_prefix, F,
                                                     (uint32_t)((smem - Q2_WARPS * Q2_WARP_WORDS * 4) / 4));
    return cudaGetLastError();
}

cudaError_t launch_query(const FrameJob* d_jobs, const uint32_t* d_cent_prefix, int F, uint32_t total_centuries,
                         uint32_t max_l_bits, int variant, int sm_count, int smem_bytes_cap, cudaStream_t st) {
    if (F <= 0 || total_centuries == 0) return cudaSuccess;
    int cap = smem_bytes_cap & ~15;
    if (cap > query_max_smem_bytes()) cap = query_max_smem_bytes() & ~15;
    const uint32_t need_words = (((max_l_bits + 31u) >> 5) + 3u) & ~3u;
    if (variant == 4 && max_l_bits <= (1u << 23)) {       // decade tiles (records carry 23-bit indices)
        const int qbytes = Q3_WARPS * Q3_WARP_WORDS * 4;
        if (cap < qbytes + 1024) cap = qbytes + 1024;
        const int bits_cap = cap - qbytes;
        const bool fits = (size_t)need_words * 4 <= (size_t)bits_cap;
        const int smem = qbytes + (fits ? (int)(need_words * 4 < 16 ? 16 : need_words * 4) : bits_cap);
        uint32_t grid = (uint32_t)sm_count;
        const uint32_t max_useful = (total_centuries + Q3_THREADS - 1) / Q3_THREADS;
        if (grid > max_useful) grid = max_useful;
        if (grid < 1u) grid = 1u;
        cudaError_t e;
        if (fits) {
            e = cudaFuncSetAttribute(k_query3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            k_query3<0><<<grid, Q3_THREADS, smem, st>>>(d_jobs, d_cent_prefix, F, (uint32_t)((smem - qbytes) / 4));
        } else {
            e = cudaFuncSetAttribute(k_query3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            k_query3<1><<<grid, Q3_THREADS, smem, st>>>(d_jobs, d_cent_prefix, F, (uint32_t)((smem - qbytes) / 4));
        }
        return cudaGetLastError();
    }
    if (variant >= 1) {                                   // staged, queue-compacted kernels
        if (variant == 4) variant = 1;
        const int qbytes = Q2_WARPS * Q2_WARP_WORDS * 4;
        if (cap < qbytes + 1024) cap = qbytes + 1024;
        const int bits_cap = cap - qbytes;
        const bool fits = (size_t)need_words * 4 <= (size_t)bits_cap;
        const uint32_t half_words = ((((need_words + 1u) >> 1) + 3u) & ~3u) + 4u;
        if (variant == 2 && !fits && (size_t)half_words * 4 <= (size_t)bits_cap && sm_count >= 2) {
            // two-CTA clusters: each CTA holds half of the array, the other half is probed through DSMEM
            const int smem = qbytes + (int)(half_words * 4);
            cudaError_t e = cudaFuncSetAttribute(k_query2c, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            uint32_t grid = (uint32_t)(sm_count & ~1);
            const uint32_t max_useful = 2u * ((total_centuries + 2 * Q2_THREADS - 1) / (2 * Q2_THREADS));
            if (grid > max_useful) grid = max_useful;
            if (grid < 2u) grid = 2u;
            k_query2c<<<grid, Q2_THREADS, smem, st>>>(d_jobs, d_cent_prefix, F, (uint32_t)((smem - qbytes) / 4));
            return cudaGetLastError();
        }
        const int smem = qbytes + (fits ? (int)(need_words * 4 < 16 ? 16 : need_words * 4) : bits_cap);
        if (variant == 3)
            return !fits ? launch_query2_t<true, 1>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st)
                         : launch_query2_t<false, 1>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st);
        return !fits ? launch_query2_t<true, 0>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st)
                     : launch_query2_t<false, 0>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st);
    }
    int smem = (size_t)need_words * 4 > (size_t)cap ? cap : (int)(need_words * 4);
    if (smem < 16) smem = 16;
    cudaError_t e = cudaFuncSetAttribute(k_query, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_query, QT, (size_t)smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(sm_count * per_sm);
    const uint32_t max_useful = (total_centuries + QT - 1) / QT;
    if (grid > max_useful) grid = max_useful;
    if (grid < 1u) grid = 1u;
    k_query<<<grid, QT, smem, st>>>(d_jobs, d_cent_prefix, F, (uint32_t)(smem / 4));
    return cudaGetLastError();
}

static uint32_t witness_chunks(int F, uint32_t max_centuries, int sm_count) {
    uint32_t ch = (uint32_t)((4 * sm_count + F - 1) / F);             // aim at ~4 CTAs per SM (two resident at a time)
    const uint32_t by_size = max_centuries / 4096u;                   // at least 4096 centuries (4 rounds) per chunk
    if (ch > by_size) ch = by_size;
    if (ch > 32u) ch = 32u;
    if (ch < 1u) ch = 1u;
    return ch;
}
// K3b: d_scratch holds F*32 uint32 (pass counts per chunk)
cudaError_t launch_witness(const FrameJob* d_jobs, int F, uint32_t max_centuries, int sm_count, uint32_t* d_scratch, uint32_t* d_wlen,
                           cudaStream_t st) {
    if (F <= 0) return cudaSuccess;
    const uint32_t ch = witness_chunks(F, max_centuries, sm_count);
    if (ch > 1u) k_pass_count<<<dim3(ch, (unsigned)F), 256, 0, st>>>(d_jobs, ch, d_scratch);
    k_witness<<<dim3(ch, (unsigned)F), 1024, 0, st>>>(d_jobs, ch, d_scratch, d_wlen);
    k_finalize<<<dim3(32, (unsigned)F), 256, 0, st>>>(d_jobs, d_wlen);
    return cudaGetLastError();
}
cudaError_t launch_expand(const FrameJob* d_jobs, int F, uint32_t max_centuries, int sm_count, uint32_t* d_scratch, uint32_t* d_consumed,
                          cudaStream_t st) {
    if (F <= 0) return cudaSuccess;
    const uint32_t ch = witness_chunks(F, max_centuries, sm_count);
    if (ch > 1u) k_pass_count<<<dim3(ch, (unsigned)F), 256, 0, st>>>(d_jobs, ch, d_scratch);
    k_expand<<<dim3(ch, (unsigned)F), 1024, 0, st>>>(d_jobs, ch, d_scratch, d_consumed);
    return cudaGetLastError();
}
static inline unsigned grid_for(size_t n, unsigned block) {
    size_t g = (n + block - 1) / block;
    if (g > 148u * 16u) g = 148u * 16u;
    if (g < 1) g = 1;
    return (unsigned)g;
}
cudaError_t launch_gather_scatter(const GatherJob* d_jobs, int F, int scatter, uint32_t* d_counts, cudaStream_t st) {
    if (F <= 0) return cudaSuccess;
    k_gather_scatter<<<F, 1024, 0, st>>>(d_jobs, scatter, d_counts);
    return cudaGetLastError();
}
cudaError_t launch_median5(const void* d_in, uint32_t pix_stride, uint32_t H, uint32_t W, int sample_bytes, void* d_out, cudaStream_t st) {
    if (H == 0 || W == 0) return cudaSuccess;
    dim3 grid((W + 31u) / 32u, (H + 7u) / 8u);
    if (sample_bytes == 1) k_median5<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)d_in, pix_stride, H, W, (uint8_t*)d_out);
    else if (sample_bytes == 2) k_median5<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)d_in, pix_stride, H, W, (uint16_t*)d_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_bitrev(uint32_t* d_words, size_t nwords, cudaStream_t st) {
    if (nwords == 0) return cudaSuccess;
    k_bitrev<<<grid_for(nwords, 256), 256, 0, st>>>(d_words, nwords);
    return cudaGetLastError();
}
cudaError_t launch_unpack_bits(const uint32_t* d_words, uint8_t* d_out, size_t nbits, cudaStream_t st) {
    if (nbits == 0) return cudaSuccess;
    k_unpack_bits<<<grid_for(nbits, 256), 256, 0, st>>>(d_words, d_out, nbits);
    return cudaGetLastError();
}
cudaError_t launch_pack_bytes(const uint8_t* d_bytes, uint32_t* d_words, size_t nbits, cudaStream_t st) {
    if (nbits == 0) return cudaSuccess;
    k_pack_bytes<<<grid_for((nbits + 31) / 32, 256), 256, 0, st>>>(d_bytes, d_words, nbits);
    return cudaGetLastError();
}

cudaError_t launch_unpack_bits_msb(const uint32_t* d_words, uint8_t* d_out, size_t nbits, cudaStream_t st) {
    if (nbits == 0) return cudaSuccess;
    k_unpack_bits_msb<<<grid_for(nbits, 256), 256, 0, st>>>(d_words, d_out, nbits);
    return cudaGetLastError();
}
cudaError_t launch_popcount(const uint32_t* d_words, size_t nwords, uint32_t* d_out, cudaStream_t st) {
    if (nwords == 0) return cudaSuccess;
    k_popcount<<<grid_for(nwords, 256), 256, 0, st>>>(d_words, nwords, d_out);
    return cudaGetLastError();
}
cudaError_t launch_count_diff(const uint32_t* a, const uint32_t* b, size_t stride_words, size_t nwords, int F,
                              uint32_t* d_out, cudaStream_t st) {
    if (nwords == 0 || F <= 0) return cudaSuccess;
    unsigned gx = grid_for(nwords, 256);
    if (gx > 64u) gx = 64u;
    dim3 g(gx, (unsigned)F);
    k_count_diff<<<g, 256, 0, st>>>(a, b, stride_words, nwords, d_out);
    return cudaGetLastError();
}
cudaError_t launch_items_u32(const FrameJob* d_job, const uint32_t* d_items, uint32_t count, uint8_t* d_result, int insert,
                             cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    k_items_u32<<<grid_for(count, 256), 256, 0, st>>>(d_job, d_items, count, d_result, insert);
    return cudaGetLastError();
}
cudaError_t launch_items_str(const FrameJob* d_job, const uint8_t* d_blob, const uint64_t* d_offs, uint32_t count,
                             uint8_t* d_result, int insert, int standard_k, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    k_items_str<<<grid_for(count, 128), 128, 0, st>>>(d_job, d_blob, d_offs, count, d_result, insert, standard_k);
    return cudaGetLastError();
}
cudaError_t launch_hash_debug(const uint32_t* d_items, uint32_t count, uint64_t seed, uint64_t* d_out, int mode,
                              cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    k_hash_debug<<<grid_for(count, 256), 256, 0, st>>>(d_items, count, seed, d_out, mode);
    return cudaGetLastError();
}

}  // namespace rbf
