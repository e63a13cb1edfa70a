// Microbenchmark (not product code): what limits independent random reads over a large HBM buffer on a
// B200, and which load flavour sustains the most.  Run plain for rates, or under
// `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` for DRAM bytes per access.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_micro gather_micro.cu
//   ./gather_micro [mode]     mode: flavours | sizes | inflight | tma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

enum { V_LDG = 0, V_NC_NOALLOC, V_L2_64B, V_CG, V_V8_EF, V_V8, V_V8_64B, V_NC_NOALLOC_64B, V_CPASYNC, V_COUNT };
static const char *kNames[V_COUNT] = {"__ldg.v4", "ld.nc.L1::no_allocate.v4", "ld.global.L2::64B.v4", "ld.global.cg.v4",
                                      "ld.nc.L2::evict_first.v8.b32", "ld.global.v8.b32", "ld.global.L2::64B.v8.b32",
                                      "ld.nc.L1::no_alloc.L2::64B.v4", "cp.async.cg.16"};

template <int V>
__device__ __forceinline__ uint32_t load16(const uint4 *p, uint32_t smem_addr) {
    uint4 v = make_uint4(0, 0, 0, 0);
    uint32_t a = 0, b = 0, c = 0, d = 0;
    const uint4 *q = (const uint4 *)((uintptr_t)p & ~(uintptr_t)31);
    if (V == V_LDG) v = __ldg(p);
    else if (V == V_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == V_L2_64B) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == V_NC_NOALLOC_64B) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == V_CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == V_V8_EF)
        asm volatile("ld.global.nc.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(q));
    else if (V == V_V8)
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(q));
    else if (V == V_V8_64B)
        asm volatile("ld.global.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(q));
    else if (V == V_CPASYNC) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(p));
    return v.x ^ v.y ^ v.z ^ v.w ^ a ^ b ^ c ^ d;
}

template <int V, int UNROLL>
__global__ void k(const uint4 *base, uint64_t n_units, uint32_t rounds, unsigned long long *sink) {
    extern __shared__ uint4 sm[];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        uint64_t idx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) idx[u] = mix64((tid * rounds + r) * UNROLL + u + 0x9e3779b97f4a7c15ull) % n_units;
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            acc ^= load16<V>(base + idx[u], (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * UNROLL + u]));
        if (V == V_CPASYNC) {
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int u = 0; u < UNROLL; u++) acc ^= sm[threadIdx.x * UNROLL + u].x;
        }
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

static const uint4 *g_buf;
static unsigned long long *g_sink;
static int g_sms;

// Partitioned-probe emulation: record r (consumed in order by the grid) reads a random 16-byte unit inside table bin
// r / recs_per_bin -- i.e. random within a sliding window of `bin_units` units -- to see what the L2 turns a
// bin-sorted probe stream into.
template <int UNROLL>
__global__ void __launch_bounds__(256) k_window(const uint4 *base, uint64_t bin_units, uint64_t recs_per_bin, uint64_t n_recs,
                                                unsigned long long *sink) {
    uint32_t acc = 0;
    const uint64_t per_block = (uint64_t)256 * UNROLL;
    for (uint64_t r0 = (uint64_t)blockIdx.x * per_block; r0 < n_recs; r0 += (uint64_t)gridDim.x * per_block) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t r = r0 + (uint64_t)u * 256 + threadIdx.x;
            const uint64_t bin = r / recs_per_bin;
            const uint64_t idx = bin * bin_units + mix64(r + 0x9e3779b97f4a7c15ull) % bin_units;
            v[u] = __ldg(base + idx);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

static void run_window(double table_gb, double bin_mb, double recs_per_unit, int bps) {
    const uint64_t bin_units = (uint64_t)(bin_mb * 1e6 / 16);
    const uint64_t n_bins = (uint64_t)(table_gb * 1e9 / 16) / bin_units;
    const uint64_t recs_per_bin = (uint64_t)(bin_units * recs_per_unit);
    const uint64_t n_recs = recs_per_bin * n_bins;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_window<4><<<g_sms * bps, 256>>>(g_buf, bin_units, recs_per_bin, n_recs, g_sink);
    CK(cudaEventRecord(e0));
    k_window<4><<<g_sms * bps, 256>>>(g_buf, bin_units, recs_per_bin, n_recs, g_sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("window: table %.1f GB, bins of %.0f MB, %.2f probes per 16-B slot (%.0fM probes), bps %d: %8.3f ms  %7.2f G probes/s\n",
           table_gb, bin_mb, recs_per_unit, n_recs / 1e6, bps, ms, n_recs / ms / 1e6);
}

// TMA flavour: every lane issues UNROLL 1-D bulk copies of BYTES into shared memory, one mbarrier per warp
template <int UNROLL, int BYTES>
__global__ void k_tma(const uint8_t *base, uint64_t n_units, uint32_t rounds, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *bars = (uint64_t *)smem;                                      // one per warp
    uint8_t *buf = smem + 128 + (size_t)threadIdx.x * UNROLL * BYTES;  // after the barriers
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, phase = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32u * UNROLL * BYTES) : "memory");
        __syncwarp();
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t idx = mix64((tid * rounds + r) * UNROLL + u + 0x9e3779b97f4a7c15ull) % n_units;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + u * BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(base + idx * BYTES), "r"((uint32_t)BYTES), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        phase ^= 1;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= *(const uint32_t *)(buf + u * BYTES);
        __syncwarp();
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}


template <int V, int UNROLL>
static double run(uint64_t n_units, int bps, int threads, uint32_t rounds, bool print = true) {
    const unsigned blocks = g_sms * bps;
    const size_t smem = V == V_CPASYNC ? (size_t)threads * UNROLL * 16 : 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<V, UNROLL><<<blocks, threads, smem>>>(g_buf, n_units, rounds, g_sink);
    CK(cudaEventRecord(e0));
    k<V, UNROLL><<<blocks, threads, smem>>>(g_buf, n_units, rounds, g_sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double acc = (double)blocks * threads * rounds * UNROLL;
    if (print)
        printf("%-30s buf %6.2f GB  bps %2d x %4d thr x unroll %d (%6d loads in flight/SM)  %8.3f ms  %7.2f G/s\n", kNames[V],
               n_units * 16 / 1e9, bps, threads, UNROLL, bps * threads * UNROLL, ms, acc / ms / 1e6);
    return acc / ms / 1e6;
}

template <int UNROLL, int BYTES>
static void run_tma(uint64_t bytes_total, int bps, int threads, uint32_t rounds) {
    const unsigned blocks = g_sms * bps;
    const size_t smem = 128 + (size_t)threads * UNROLL * BYTES;
    CK(cudaFuncSetAttribute(k_tma<UNROLL, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_units = bytes_total / BYTES;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_tma<UNROLL, BYTES><<<blocks, threads, smem>>>((const uint8_t *)g_buf, n_units, rounds, g_sink);
    CK(cudaEventRecord(e0));
    k_tma<UNROLL, BYTES><<<blocks, threads, smem>>>((const uint8_t *)g_buf, n_units, rounds, g_sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double acc = (double)blocks * threads * rounds * UNROLL;
    printf("cp.async.bulk %3d B             buf %6.2f GB  bps %2d x %4d thr x unroll %d (%6d copies in flight/SM)  %8.3f ms  %7.2f G/s\n",
           BYTES, bytes_total / 1e9, bps, threads, UNROLL, bps * threads * UNROLL, ms, acc / ms / 1e6);
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "flavours";
    const double max_gb = 32.0;
    const uint64_t max_units = (uint64_t)(max_gb * 1e9 / 16);
    uint4 *buf;
    CK(cudaMalloc(&buf, max_units * 16));
    CK(cudaMalloc(&g_sink, 8));
    CK(cudaMemset(buf, 0x5a, max_units * 16));
    CK(cudaMemset(g_sink, 0, 8));
    g_buf = buf;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    const uint64_t u8 = (uint64_t)(8e9 / 16);
    if (!strcmp(mode, "flavours")) {
        run<V_LDG, 4>(u8, 8, 256, 32);
        run<V_NC_NOALLOC, 4>(u8, 8, 256, 32);
        run<V_L2_64B, 4>(u8, 8, 256, 32);
        run<V_NC_NOALLOC_64B, 4>(u8, 8, 256, 32);
        run<V_CG, 4>(u8, 8, 256, 32);
        run<V_V8_EF, 4>(u8, 8, 256, 32);
        run<V_V8, 4>(u8, 8, 256, 32);
        run<V_V8_64B, 4>(u8, 8, 256, 32);
        run<V_CPASYNC, 4>(u8, 8, 256, 32);
    } else if (!strcmp(mode, "sizes")) {
        for (double gb : {0.06, 0.12, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0, 32.0}) {
            run<V_LDG, 4>((uint64_t)(gb * 1e9 / 16), 8, 256, 32);
            run<V_V8_EF, 4>((uint64_t)(gb * 1e9 / 16), 8, 256, 32);
        }
    } else if (!strcmp(mode, "inflight")) {
        for (int bps : {1, 2, 4, 8}) {
            run<V_LDG, 1>(u8, bps, 256, 128);
            run<V_LDG, 2>(u8, bps, 256, 64);
            run<V_LDG, 4>(u8, bps, 256, 32);
            run<V_LDG, 8>(u8, bps, 256, 16);
        }
        for (int bps : {1, 2, 4, 8}) {
            run<V_V8_EF, 1>(u8, bps, 256, 128);
            run<V_V8_EF, 2>(u8, bps, 256, 64);
            run<V_V8_EF, 4>(u8, bps, 256, 32);
            run<V_V8_EF, 8>(u8, bps, 256, 16);
        }
        run<V_LDG, 1>(u8, 1, 32, 256);
        run<V_LDG, 1>(u8, 1, 64, 256);
        run<V_LDG, 1>(u8, 1, 128, 256);
    } else if (!strcmp(mode, "window")) {
        for (double bin_mb : {8.0, 16.0, 32.0, 64.0})
            for (int bps : {4, 8}) run_window(8.0, bin_mb, 0.57, bps);  // C2: 290M probes over 508M slots
        run_window(4.0, 32.0, 4.7, 8);                                    // C3: 1.17G probes over 248M slots
    } else if (!strcmp(mode, "tma")) {
        run_tma<1, 16>(8000000000ull, 4, 256, 64);
        run_tma<4, 16>(8000000000ull, 4, 256, 32);
        run_tma<4, 32>(8000000000ull, 4, 256, 32);
        run_tma<8, 32>(8000000000ull, 4, 256, 16);
        run_tma<4, 32>(8000000000ull, 8, 256, 32);
        run_tma<4, 128>(8000000000ull, 2, 256, 32);
    }
    return 0;
}
