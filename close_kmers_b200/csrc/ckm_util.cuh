// Plumbing kernels: exclusive prefix sums (CSR offsets) and compaction of the per-protein regions
// into the dense arrays that are copied back to the host.
#pragma once
#include "ckm_common.cuh"
#include "ckm_scan.cuh"

namespace ckm {

constexpr int kPsThreads = 256;
constexpr int kPsItems = 8;
constexpr int kPsTile = kPsThreads * kPsItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total) {
    __shared__ uint64_t warp_sums[kPsThreads / 32];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    uint64_t wbase = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kPsThreads / 32; k++) {
        const uint64_t s = warp_sums[k];
        if ((uint32_t)k < wid) wbase += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return wbase + incl - v;
}

__global__ void __launch_bounds__(kPsThreads) ps_block_sums(const uint32_t *__restrict__ in, uint64_t n,
                                                            uint64_t *__restrict__ block_sums) {
    const uint64_t t0 = (uint64_t)blockIdx.x * kPsTile + (uint64_t)threadIdx.x * kPsItems;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kPsItems; k++)
        if (t0 + k < n) s += in[t0 + k];
    uint64_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of block_sums in place, chunk by chunk with a running carry
__global__ void __launch_bounds__(kPsThreads) ps_spine(uint64_t *__restrict__ block_sums, uint64_t nb) {
    uint64_t carry = 0;
    for (uint64_t c0 = 0; c0 < nb; c0 += kPsThreads) {
        const uint64_t idx = c0 + threadIdx.x;
        const uint64_t v = idx < nb ? block_sums[idx] : 0;
        uint64_t tot;
        const uint64_t ex = block_exclusive_scan(v, &tot);
        if (idx < nb) block_sums[idx] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(kPsThreads) ps_finish(const uint32_t *__restrict__ in, uint64_t n,
                                                        const uint64_t *__restrict__ block_sums,
                                                        uint64_t *__restrict__ out /* n+1 */) {
    const uint64_t t0 = (uint64_t)blockIdx.x * kPsTile + (uint64_t)threadIdx.x * kPsItems;
    uint32_t v[kPsItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kPsItems; k++) {
        v[k] = (t0 + k < n) ? in[t0 + k] : 0u;
        s += v[k];
    }
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kPsItems; k++) {
        if (t0 + k < n) out[t0 + k] = ex;
        ex += v[k];
        if (t0 + k + 1 == n) out[n] = ex;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// ---- region -> CSR compaction: one warp per protein -------------------------------------------------
__global__ void __launch_bounds__(256)
export_hits_kernel(const uint64_t *__restrict__ offsets, const HitRec *__restrict__ hits,
                   const uint64_t *__restrict__ hit_keys, const uint16_t *__restrict__ hit_avg,
                   const uint64_t *__restrict__ hit_offsets, uint32_t n, ckm_hit_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const uint64_t base = offsets[w], o0 = hit_offsets[w];
    const uint32_t cnt = (uint32_t)(hit_offsets[w + 1] - o0);
    for (uint32_t k = lane; k < cnt; k += 32) {
        const HitRec h = hits[base + k];
        ckm_hit_t r;
        r.which_kmer = hit_keys[base + k];
        r.offset = h.pos;
        r.otu_index = h.oI;
        r.function_index = (int32_t)h.fI;
        r.function_wt = h.wt;
        r.avg_from_end = hit_avg[base + k];
        r.pad_ = 0;
        r.pad2_ = 0;
        out[o0 + k] = r;
    }
}

__global__ void __launch_bounds__(256)
export_calls_kernel(const uint64_t *__restrict__ offsets, const ckm_call_t *__restrict__ calls,
                    const uint64_t *__restrict__ call_offsets, uint32_t n, int min_hits, ckm_call_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t o0 = call_offsets[i];
    const uint32_t cnt = (uint32_t)(call_offsets[i + 1] - o0);
    const ckm_call_t *src = calls + call_region_base(offsets[i], i, min_hits);
    for (uint32_t k = 0; k < cnt; k++) out[o0 + k] = src[k];
}

__global__ void __launch_bounds__(256)
export_otus_kernel(const uint64_t *__restrict__ offsets, const ckm_otu_t *__restrict__ otus,
                   const uint64_t *__restrict__ otu_offsets, uint32_t n, ckm_otu_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t o0 = otu_offsets[i];
    const uint32_t cnt = (uint32_t)(otu_offsets[i + 1] - o0);
    const ckm_otu_t *src = otus + offsets[i];
    for (uint32_t k = 0; k < cnt; k++) out[o0 + k] = src[k];
}

}  // namespace ckm
