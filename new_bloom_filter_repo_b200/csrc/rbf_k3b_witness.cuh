// rbf_k3b_witness.cuh -- K3b witness (ivc:253) and K4b decode expand (ivc:299-304).  Included by rbf_kernels.cu inside namespace rbf.
#pragma once

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
