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
//
// One translation unit: this file holds the PTX helpers and the launchers; the kernels live in
// rbf_k1_threshold.cuh, rbf_k2_insert.cuh, rbf_k3_query.cuh, rbf_k3b_witness.cuh and rbf_aux_kernels.cuh.
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

#include "rbf_k1_threshold.cuh"
#include "rbf_k2_insert.cuh"
#include "rbf_k3_query.cuh"
#include "rbf_k3b_witness.cuh"
#include "rbf_aux_kernels.cuh"

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int PB, int S>
static cudaError_t launch_threshold_t(const PairJob* d_pairs, int F, uint32_t npix, int thr, int any_mode, uint32_t* d_ones,
                                      uint32_t* d_resid, int variant, int sm_count, int ctas_per_sm, cudaStream_t st) {
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
    const uint32_t cap = (uint32_t)(sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 32));
    if ((uint64_t)bx * (uint64_t)F > cap) { bx = (cap + (uint32_t)F - 1u) / (uint32_t)F; if (bx < 1u) bx = 1u; }
    dim3 grid(bx, (unsigned)F);
    k_threshold<PB, S><<<grid, 256, 0, st>>>(d_pairs, npix, thr, any_mode, d_ones, d_resid);
    return cudaGetLastError();
}

// gray_mode: the mask is taken on cv2.COLOR_BGR2GRAY of the three samples instead of on sample 0 (ivc:792-795)
template <int PB, int S>
static cudaError_t launch_threshold_gray_t(const PairJob* d_pairs, int F, uint32_t npix, int thr, int any_mode, uint32_t* d_ones,
                                           uint32_t* d_resid, int sm_count, int ctas_per_sm, cudaStream_t st) {
    const uint32_t nwords = (npix + 31u) >> 5;
    uint32_t bx = (nwords + 255u) / 256u;
    const uint32_t cap = (uint32_t)(sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 32));
    if ((uint64_t)bx * (uint64_t)F > cap) { bx = (cap + (uint32_t)F - 1u) / (uint32_t)F; if (bx < 1u) bx = 1u; }
    dim3 grid(bx, (unsigned)F);
    k_threshold<PB, S, true><<<grid, 256, 0, st>>>(d_pairs, npix, thr, any_mode, d_ones, d_resid);
    return cudaGetLastError();
}

cudaError_t launch_threshold(const PairJob* d_pairs, int F, uint32_t npix, int channels, int sample_bytes, int thr_int,
                             int any_mode, int gray_mode, uint32_t* d_ones, uint32_t* d_resid, int variant, int sm_count,
                             int ctas_per_sm, cudaStream_t st) {
    if (F <= 0 || npix == 0) return cudaSuccess;
    const int pb = channels * sample_bytes;
    if (gray_mode) {
        if (pb == 3 && sample_bytes == 1) return launch_threshold_gray_t<3, 1>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, sm_count, ctas_per_sm, st);
        if (pb == 6 && sample_bytes == 2) return launch_threshold_gray_t<6, 2>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, sm_count, ctas_per_sm, st);
        return cudaErrorInvalidValue;
    }
    if (pb == 3 && sample_bytes == 1) return launch_threshold_t<3, 1>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, ctas_per_sm, st);
    if (pb == 6 && sample_bytes == 2) return launch_threshold_t<6, 2>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, ctas_per_sm, st);
    if (pb == 1 && sample_bytes == 1) return launch_threshold_t<1, 1>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, ctas_per_sm, st);
    if (pb == 2 && sample_bytes == 2) return launch_threshold_t<2, 2>(d_pairs, F, npix, thr_int, any_mode, d_ones, d_resid, variant, sm_count, ctas_per_sm, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_insert(const FrameJob* d_jobs, int F, uint32_t max_centuries, int variant, int sm_count, cudaStream_t st) {
    if (F <= 0 || max_centuries == 0) return cudaSuccess;
    if (variant >= 1) {                                   // dense, warp-compacted insert
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

template <bool HYBRID>
static cudaError_t launch_query2_t(const FrameJob* d_jobs, const uint32_t* d_cent_prefix, int F, uint32_t total_centuries,
                                   int sm_count, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_query2<HYBRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    uint32_t grid = (uint32_t)sm_count;
    const uint32_t max_useful = (total_centuries + Q2_THREADS - 1) / Q2_THREADS;
    if (grid > max_useful) grid = max_useful;
    if (grid < 1u) grid = 1u;
    k_query2<HYBRID><<<grid, Q2_THREADS, smem, st>>>(d_jobs, d_cent_prefix, F, (uint32_t)((smem - Q2_WARPS * Q2_WARP_WORDS * 4) / 4));
    return cudaGetLastError();
}

// variant: 0 per-lane kernel; 1 staged rings; 5 decade tiles (half-decade tiles for 2^23 < m <= 2^24, rings beyond); 6 half-decade
// tiles for every m <= 2^24.  (Round 1's 2 = DSMEM cluster, 3 = dense A+B and 4 = tiles without the batch carry were measured
// slower and removed; DESIGN.md keeps their numbers.  The values still select the current default.)
cudaError_t launch_query(const FrameJob* d_jobs, const uint32_t* d_cent_prefix, int F, uint32_t total_centuries,
                         uint32_t max_l_bits, int variant, int sm_count, int smem_bytes_cap, int tile_warps, cudaStream_t st) {
    if (F <= 0 || total_centuries == 0) return cudaSuccess;
    if (tile_warps < 4 || tile_warps > Q4_WARPS) tile_warps = Q4_WARPS;
    int cap = smem_bytes_cap & ~15;
    if (cap > query_max_smem_bytes()) cap = query_max_smem_bytes() & ~15;
    const uint32_t need_words = (((max_l_bits + 31u) >> 5) + 3u) & ~3u;
    if (variant >= 2 && variant <= 4) variant = 5;
    if (variant >= 5 && max_l_bits <= (1u << 24)) {       // decade tiles (full stage-B batches with carry)
        const bool half = (variant == 6) || max_l_bits > (1u << 23);      // half-decade tiles: 24-bit record indices
        const int qbytes = 4 * (Q4_TABLE_WORDS + tile_warps * (half ? Q4Cfg<5>::WARP_WORDS : Q4Cfg<10>::WARP_WORDS));
        if (cap < qbytes + 1024) cap = qbytes + 1024;
        const int bits_cap = cap - qbytes;
        const bool fits = (size_t)need_words * 4 <= (size_t)bits_cap;
        const int smem = qbytes + (fits ? (int)(need_words * 4 < 16 ? 16 : need_words * 4) : bits_cap);
        uint32_t grid = (uint32_t)sm_count;
        const uint32_t max_useful = (total_centuries + 32u * tile_warps - 1) / (32u * tile_warps);
        if (grid > max_useful) grid = max_useful;
        if (grid < 1u) grid = 1u;
        const uint32_t words_cap = (uint32_t)((smem - qbytes) / 4);
#define RBF_LAUNCH_Q4(PM, TY)                                                                                         \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(k_query4<PM, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);    \
        if (e != cudaSuccess) return e;                                                                               \
        k_query4<PM, TY><<<grid, 32 * tile_warps, smem, st>>>(d_jobs, d_cent_prefix, F, words_cap);                        \
    } while (0)
        if (fits) { if (half) RBF_LAUNCH_Q4(0, 5); else RBF_LAUNCH_Q4(0, 10); }
        else      { if (half) RBF_LAUNCH_Q4(1, 5); else RBF_LAUNCH_Q4(1, 10); }
#undef RBF_LAUNCH_Q4
        return cudaGetLastError();
    }
    if (variant >= 1) {                                   // staged rings (also the home of filters beyond 2^24 bits)
        const int qbytes = Q2_WARPS * Q2_WARP_WORDS * 4;
        if (cap < qbytes + 1024) cap = qbytes + 1024;
        const int bits_cap = cap - qbytes;
        const bool fits = (size_t)need_words * 4 <= (size_t)bits_cap;
        const int smem = qbytes + (fits ? (int)(need_words * 4 < 16 ? 16 : need_words * 4) : bits_cap);
        return !fits ? launch_query2_t<true>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st)
                     : launch_query2_t<false>(d_jobs, d_cent_prefix, F, total_centuries, sm_count, smem, st);
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
    uint32_t ch = (uint32_t)((16 * sm_count + F - 1) / F);            // >= 8 waves of the two 1024-thread CTAs an SM holds: at 2 waves
                                                                      // (round 1) the ragged last wave cost up to a third of K3b
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
uint32_t gather_chunks(uint32_t npix) {
    uint32_t ch = ((npix + 31u) >> 5) / 4096u;             // >= 4 rounds of 1024 threads per chunk
    if (ch > (uint32_t)GS_MAX_CHUNKS) ch = GS_MAX_CHUNKS;
    return ch < 1u ? 1u : ch;
}
// d_counts: F * gather_chunks(npix) uint32 of scratch; d_totals (may be NULL): set pixels per pair
cudaError_t launch_gather_scatter(const GatherJob* d_jobs, int F, int scatter, uint32_t npix, uint32_t pix_bytes, uint32_t* d_counts,
                                  uint32_t* d_totals, cudaStream_t st) {
    if (F <= 0) return cudaSuccess;
    const uint32_t ch = gather_chunks(npix);
    const dim3 grid(ch, (unsigned)F);
    k_mask_chunk_count<<<grid, 256, 0, st>>>(d_jobs, ch, d_counts);
    switch (pix_bytes) {
    case 1: k_gather_scatter<1><<<grid, 1024, 0, st>>>(d_jobs, scatter, ch, d_counts, d_totals); break;
    case 2: k_gather_scatter<2><<<grid, 1024, 0, st>>>(d_jobs, scatter, ch, d_counts, d_totals); break;
    case 3: k_gather_scatter<3><<<grid, 1024, 0, st>>>(d_jobs, scatter, ch, d_counts, d_totals); break;
    case 6: k_gather_scatter<6><<<grid, 1024, 0, st>>>(d_jobs, scatter, ch, d_counts, d_totals); break;
    default: return cudaErrorInvalidValue;
    }
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

cudaError_t launch_push_slots(const uint32_t* d_bits, size_t stride_w, uint32_t slot_w, uint32_t pairs, uint32_t* const* recv,
                              uint32_t* const* flags, int nranks, int rank, size_t dst_off_w, uint32_t seq, int sm_count, cudaStream_t st) {
    if (nranks < 1 || nranks > PEER_MAX) return cudaErrorInvalidValue;
    PeerTable pt;
    for (int r = 0; r < PEER_MAX; r++) { pt.recv[r] = r < nranks ? recv[r] : nullptr; pt.flags[r] = r < nranks ? flags[r] : nullptr; }
    if ((slot_w & 3u) || (stride_w & 3u) || (dst_off_w & 3u)) return cudaErrorInvalidValue;
    const size_t total = (size_t)pairs * (slot_w >> 2);
    size_t grid = (total + 255) / 256;
    if (grid > (size_t)sm_count * 8) grid = (size_t)sm_count * 8;
    if (grid < 1) grid = 1;
    k_push_slots<<<(unsigned)grid, 256, 0, st>>>(d_bits, stride_w, slot_w, pairs, pt, nranks, dst_off_w);
    k_peer_signal<<<1, 32, 0, st>>>(pt, nranks, rank, seq);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(const uint32_t* d_flags, int nranks, uint32_t seq, long long timeout_cycles, uint32_t* d_err, cudaStream_t st) {
    k_peer_wait<<<1, 32, 0, st>>>(d_flags, nranks, seq, timeout_cycles, d_err);
    return cudaGetLastError();
}

}  // namespace rbf
