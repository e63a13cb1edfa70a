// libckm.so -- context, table loader and the C ABI declared in include/ckm.h.
//
// Host side of the drop-in boundary: what a request handler used to do through a per-thread KmerGuts
// (kguts.h:334-372) it now does with one batch call on a ckm_ctx.  No CPU fallback exists: every
// compute entry point launches the sm_100a kernels or fails with CKM_ECUDA.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "ckm_common.cuh"
#include "ckm_ctx.h"
#include "ckm_probe.cuh"
#include "ckm_probe_group.cuh"
#include "ckm_chain.cuh"
#include "ckm_hint.cuh"
#include "ckm_scan.cuh"
#include "ckm_warp_scan.cuh"
#include "ckm_pc.cuh"
#include "ckm_packed.cuh"
#include "ckm_util.cuh"

using namespace ckm;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int ckm_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *ckm_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------------------------------
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) return ckm_fail(CKM_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    cap = want;
    return 0;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int PinBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return ckm_fail(CKM_ENOMEM, "cudaMallocHost(%zu): %s", want, cudaGetErrorString(e));
    cap = want;
    return 0;
}
void PinBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

// ---------------------------------------------------------------------------------------------------
// table load: verbatim upload, then repack on the device (see ckm_common.cuh for the layout)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_table_kernel(const uint64_t *__restrict__ raw /* 3 x u64 per slot */, uint64_t num_sigs, uint4 *__restrict__ packed,
                  unsigned int *__restrict__ misfit) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_sigs) return;
    const uint64_t k = raw[3 * i], a = raw[3 * i + 1], b = raw[3 * i + 2];
    uint4 v;
    if (k > CKM_MAX_ENCODED) {  // empty (kguts.cc:587, 596)
        v.x = 0;
        v.y = 0x8u;
        v.z = 0;
        v.w = 0;
    } else {
        const int32_t oI = (int32_t)(uint32_t)a;
        const uint32_t avg = (uint32_t)(a >> 32) & 0xFFFFu;
        const uint32_t fI = (uint32_t)b;
        const uint32_t wt = (uint32_t)(b >> 32);
        const uint32_t o1 = (uint32_t)(oI + 1);
        if (fI >= kPackedFieldLimit || oI < -1 || o1 >= kPackedFieldLimit) atomicOr(misfit, 1u);
        v.x = (uint32_t)k;
        v.y = (uint32_t)(k >> 32) | (avg << 4) | ((o1 & 0xFFFu) << 20);
        v.z = wt;
        v.w = (fI & (kPackedFieldLimit - 1)) | ((o1 >> 12) << 22);
    }
    packed[i] = v;
}

// The library's own table: every k-mer of the image re-inserted, by linear probing from own_home(key), into 2^hbits packed slots.
// An image may hold a k-mer twice (the reference's builder never checks, kguts.cc:202-222); lookup_hash_entry then always ends
// at the first one in probe order, so only that one is carried over and a lookup here returns what a lookup there returns.
__global__ void __launch_bounds__(256) fill_empty_kernel(uint4 *__restrict__ slots, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = make_uint4(0u, 0x8u, 0u, 0u);
}
__global__ void __launch_bounds__(256)
rehash_kernel(const uint64_t *__restrict__ raw /* 3 x u64 per slot */, uint64_t image_buckets, uint64_t image_magic, uint4 *__restrict__ slots,
              uint32_t hbits, unsigned int *__restrict__ misfit) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= image_buckets) return;
    const uint64_t k = raw[3 * i], a = raw[3 * i + 1], b = raw[3 * i + 2];
    if (k > CKM_MAX_ENCODED) return;
    // is slot i where the reference's lookup of k ends?  Not if the k-mer sits in an earlier slot of its probe sequence, and not
    // if an empty slot comes first (a builder never leaves one there; a hand-made image might: the reference would not find k)
    for (uint64_t h = fast_mod(k, image_buckets, image_magic); h != i; h = (h + 1 == image_buckets) ? 0 : h + 1) {
        const uint64_t kh = raw[3 * h];
        if (kh == k || kh > CKM_MAX_ENCODED) return;
    }
    const int32_t oI = (int32_t)(uint32_t)a;
    const uint32_t avg = (uint32_t)(a >> 32) & 0xFFFFu, fI = (uint32_t)b, wt = (uint32_t)(b >> 32), o1 = (uint32_t)(oI + 1);
    if (fI >= kPackedFieldLimit || oI < -1 || o1 >= kPackedFieldLimit) {
        atomicOr(misfit, 1u);
        return;
    }
    const uint32_t x = (uint32_t)k, y = (uint32_t)(k >> 32) | (avg << 4) | ((o1 & 0xFFFu) << 20);
    const unsigned long long mine = (unsigned long long)x | ((unsigned long long)y << 32), empty = 0x8ull << 32;
    const uint32_t mask = hbits == 32u ? 0xFFFFFFFFu : (1u << hbits) - 1u;
    for (uint32_t h = own_home(k, hbits);; h = (h + 1u) & mask) {
        if (atomicCAS(reinterpret_cast<unsigned long long *>(slots + h), empty, mine) == empty) {
            reinterpret_cast<uint2 *>(slots + h)[1] = make_uint2(wt, (fI & (kPackedFieldLimit - 1)) | ((o1 >> 12) << 22));
            return;
        }
    }
}

// bit h of occupied[] = slot h holds a k-mer; one thread per 32 slots
template <bool PACKED>
__global__ void __launch_bounds__(256)
occupancy_kernel(const void *__restrict__ slots, uint64_t num_sigs, uint32_t *__restrict__ occupied) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= num_sigs) return;
    uint32_t bits = 0;
    for (uint32_t b = 0; b < 32; b++) {
        const uint64_t h = w * 32 + b;
        if (h >= num_sigs) break;
        bool occ;
        if (PACKED) occ = !(reinterpret_cast<const uint4 *>(slots)[h].y & 0x8u);
        else occ = reinterpret_cast<const uint64_t *>(slots)[3 * h] <= CKM_MAX_ENCODED;
        bits |= occ ? (1u << b) : 0u;
    }
    occupied[w] = bits;
}

static int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    return ckm_fail(CKM_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CU(x)                                  \
    do {                                       \
        int rc_ = check_cuda((x), #x);         \
        if (rc_) return rc_;                   \
    } while (0)
#define RC(x)                \
    do {                     \
        int rc_ = (x);       \
        if (rc_) return rc_; \
    } while (0)

static int select_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return ckm_fail(CKM_ECUDA, "no CUDA device available (%s); libckm has no CPU fallback",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return ckm_fail(CKM_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return ckm_fail(CKM_ECUDA, "device %d is sm_%d%d; libckm is built for sm_100a only", device, prop.major, prop.minor);
    return 0;
}

static int install_table(ckm_ctx *c, DevBuf raw, uint64_t n);
static int build_chain(ckm_ctx *c);
static int prefix_sum(ckm_ctx *c, const uint32_t *d_in, uint64_t n, uint64_t *d_out /* n+1 */);

static int upload_table(ckm_ctx *c, const ckm_image_header_t *hdr) {
    const uint64_t n = hdr->num_sigs;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(hdr + 1);
    const size_t raw_bytes = (size_t)n * kRawSlotBytes;
    DevBuf raw;
    RC(raw.ensure(raw_bytes + 64));
    const size_t chunk = (size_t)1 << 28;
    for (size_t o = 0; o < raw_bytes; o += chunk)
        if (int rc = check_cuda(cudaMemcpyAsync((uint8_t *)raw.p + o, src + o, std::min(chunk, raw_bytes - o), cudaMemcpyHostToDevice, c->stream),
                                "cudaMemcpyAsync(table)")) {
            cudaStreamSynchronize(c->stream);
            raw.release();
            return rc;
        }
    return install_table(c, raw, n);
}

// the verbatim 24-byte slots are on the device (uploaded, or built there): repack, choose the slot format, build the
// occupancy bitmap.  Takes ownership of `raw`.
static int install_table(ckm_ctx *c, DevBuf raw, uint64_t n) {
    DevBuf packed, flag;
    unsigned int misfit = 0;
    c->image_buckets = n;
    c->hbits = 0;
    // The library's own table (rehash_kernel) where it applies: 16-byte slots fit the image's fields, fewer than 2^32 buckets,
    // and room for it beside the image for a moment.  Otherwise the image's order and hashing stay (packed or verbatim slots).
    if (!c->force_raw && !c->reference_hash && n) {
        uint32_t hbits = 6;
        while (hbits < 31u && (1ull << hbits) < n) hbits++;
        if ((1ull << hbits) >= n) {
            const uint64_t nb = 1ull << hbits;
            DevBuf own;
            auto body = [&]() -> int {
                RC(flag.ensure(256));
                CU(cudaMemsetAsync(flag.p, 0, 4, c->stream));
                fill_empty_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>((uint4 *)own.p, nb);
                rehash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const uint64_t *)raw.p, n, (uint64_t)((((unsigned __int128)1) << 64) / n),
                                                                                   (uint4 *)own.p, hbits, (unsigned int *)flag.p);
                c->launches += 2;
                CU(cudaMemcpyAsync(&misfit, flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaGetLastError());
                return 0;
            };
            if (own.ensure((size_t)nb * kPackedSlotBytes + 64) == 0) {
                const int rc = body();
                flag.release();
                if (rc) {
                    cudaStreamSynchronize(c->stream);
                    own.release();
                    raw.release();
                    return rc;
                }
                if (!misfit) {
                    raw.release();
                    c->table = own;
                    c->slot_bytes = kPackedSlotBytes;
                    c->hbits = hbits;
                    n = nb;  // from here on: the buckets of the table in HBM
                } else {
                    own.release();  // a field does not fit 22 bits: the verbatim slots below
                }
            } else {
                (void)cudaGetLastError();
            }
        }
    }
    if (!c->hbits) {
    // raw + packed need 40 B per bucket for a moment: without room for the packed copy the verbatim slots serve (RAW kernels)
    bool no_room = !c->force_raw && packed.ensure((size_t)n * kPackedSlotBytes + 64) != 0;
    if (no_room) {
        (void)cudaGetLastError();
        fprintf(stderr, "libckm: no device memory for the 16-byte slots (%s); keeping the 24-byte slots of the image\n", ckm_last_error());
    }
    if (!no_room && !c->force_raw) {
        auto body = [&]() -> int {
            RC(flag.ensure(256));
            CU(cudaMemsetAsync(flag.p, 0, 4, c->stream));
            if (n) {
                const uint64_t blocks = (n + 255) / 256;
                pack_table_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>((const uint64_t *)raw.p, n, (uint4 *)packed.p,
                                                                            (unsigned int *)flag.p);
                c->launches++;
            }
            CU(cudaMemcpyAsync(&misfit, flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaGetLastError());
            return 0;
        };
        const int rc = body();
        flag.release();
        if (rc) {
            cudaStreamSynchronize(c->stream);
            packed.release();
            raw.release();
            return rc;
        }
    } else if (int rc = check_cuda(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize(table upload)")) {
        packed.release();
        raw.release();
        return rc;
    }
    if (misfit || c->force_raw || no_room) {
        packed.release();
        c->table = raw;
        c->slot_bytes = kRawSlotBytes;
    } else {
        raw.release();
        c->table = packed;
        c->slot_bytes = kPackedSlotBytes;
    }
    }
    c->num_sigs = n;
    c->magic = n ? (uint64_t)((((unsigned __int128)1) << 64) / n) : 0;
    // occupancy bitmap: only worth an extra L2 access per probe when the table itself cannot live in L2
    const size_t table_bytes = (size_t)n * c->slot_bytes;
    const char *ob = getenv("CKM_OCCUPANCY_BITMAP");  // "0" disables, "1" forces
    const bool want_bitmap = ob ? ob[0] == '1' : table_bytes > (size_t)c->l2_bytes;
    c->occupied.release();
    if (want_bitmap) {
        const uint64_t words = (n + 31) / 32;
        RC(c->occupied.ensure(words * 4 + 64));
        const unsigned blocks = (unsigned)((words + 255) / 256);
        if (c->slot_bytes == kPackedSlotBytes)
            occupancy_kernel<true><<<blocks, 256, 0, c->stream>>>(c->table.p, n, (uint32_t *)c->occupied.p);
        else
            occupancy_kernel<false><<<blocks, 256, 0, c->stream>>>(c->table.p, n, (uint32_t *)c->occupied.p);
        c->launches++;
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        // keep the bitmap resident: persisting L2 window on the ctx stream (best effort)
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t bytes = words * 4;
            const size_t carve = std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof attr);
                attr.accessPolicyWindow.base_ptr = c->occupied.p;
                attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)bytes);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (!getenv("CKM_NO_L2_PERSIST")) {
                    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess)
                        (void)cudaGetLastError();
                    c->l2_window = attr.accessPolicyWindow;  // also applied to the second pipeline stream when it is created
                    c->has_l2_window = true;
                }
            } else {
                (void)cudaGetLastError();
            }
        }
    }
    // neighbour-ordered copy (ckm_chain.cuh): like the bitmap, only pays when hits are DRAM transactions
    c->chain.release();
    c->cpos.release();
    c->cres.release();
    c->cpay.release();
    c->n_chain = 0;
    const char *ch = getenv("CKM_CHAIN");  // "0" disables, "1" forces
    const bool want_chain = ch ? ch[0] == '1' : table_bytes > (size_t)c->l2_bytes;
    if (want_chain && c->slot_bytes == kPackedSlotBytes && n > 0 && n < 0xFFFFFFF0ull) {
        // an optimisation only: without the memory for it (~32 B per bucket while building) the table works as before
        if (build_chain(c)) {
            if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)
                return ckm_fail(CKM_ECUDA, "building the neighbour copy failed: %s", ckm_last_error());
            fprintf(stderr, "libckm: neighbour copy of the table not built (%s); plain hash probing\n", ckm_last_error());
        }
    }
    return 0;
}

// Build chain[] / cpos[] from the packed table on the device (steps 1-3 of ckm_chain.cuh).
static int build_chain(ckm_ctx *c) {
    const uint64_t n = c->num_sigs;
    TableView tv;
    memset(&tv, 0, sizeof tv);
    tv.slots = c->table.p;
    tv.num_sigs = n;
    tv.magic = c->magic;
    tv.m35 = magic35(n);
    tv.hbits = c->hbits;
    tv.occupied = (const uint32_t *)c->occupied.p;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, c->stream));
    DevBuf pd, succ, claim, len, start, flag;
    int rc = 0;
    auto body = [&]() -> int {
        RC(pd.ensure(n * 8));
        RC(succ.ensure(n * 4));
        RC(claim.ensure(n * 8));
        RC(len.ensure(n * 4 + 64));
        RC(start.ensure((n + 1) * 8));
        RC(flag.ensure(64));
        RC(c->cpos.ensure(n * 8 + 64));
        const unsigned blocks = (unsigned)((n + 255) / 256);
        CU(cudaMemsetAsync(claim.p, 0xFF, n * 8, c->stream));
        CU(cudaMemsetAsync(len.p, 0, n * 4, c->stream));
        CU(cudaMemsetAsync(flag.p, 0, 64, c->stream));
        chain_propose_kernel<<<blocks, 256, 0, c->stream>>>(tv, (uint64_t *)pd.p, (uint32_t *)succ.p, (unsigned long long *)claim.p);
        chain_link_kernel<<<blocks, 256, 0, c->stream>>>(n, (const uint32_t *)succ.p, (const unsigned long long *)claim.p, (uint64_t *)pd.p);
        c->launches += 2;
        unsigned int changed = 1;
        for (int round = 0; round < 26 && changed; round++) {
            CU(cudaMemsetAsync(flag.p, 0, 4, c->stream));
            chain_jump_kernel<<<blocks, 256, 0, c->stream>>>(n, (volatile uint64_t *)pd.p, (unsigned int *)flag.p);
            c->launches++;
            CU(cudaMemcpyAsync(&changed, flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
        c->chain_cycles = changed ? 1 : 0;
        if (changed) {  // cycles: their members become chains of one (succ[] is free by now and holds the marks)
            chain_cycle_mark_kernel<<<blocks, 256, 0, c->stream>>>(n, (const uint64_t *)pd.p, (uint32_t *)succ.p);
            chain_cycle_cut_kernel<<<blocks, 256, 0, c->stream>>>(n, (const uint32_t *)succ.p, (uint64_t *)pd.p);
            c->launches += 2;
        }
        chain_length_kernel<<<blocks, 256, 0, c->stream>>>(n, (const uint64_t *)pd.p, (uint32_t *)len.p);
        chain_pad_kernel<<<blocks, 256, 0, c->stream>>>(n, (uint32_t *)len.p);
        c->launches += 2;
        RC(prefix_sum(c, (const uint32_t *)len.p, n, (uint64_t *)start.p));
        uint64_t total = 0;
        CU(cudaMemcpyAsync(&total, (const uint64_t *)start.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (total >= 0xFFFFFFF0ull) return ckm_fail(CKM_ESTATE, "chain copy: %llu entries do not fit 32-bit indices", (unsigned long long)total);
        RC(c->chain.ensure((total + 32) * sizeof(uint4)));
        CU(cudaMemsetAsync(c->chain.p, 0xFF, (total + 32) * sizeof(uint4), c->stream));
        chain_place_kernel<<<blocks, 256, 0, c->stream>>>(tv, (const uint64_t *)pd.p, (const uint64_t *)start.p, (uint4 *)c->chain.p,
                                                          (uint2 *)c->cpos.p, (unsigned long long *)flag.p + 1);
        RC(c->cres.ensure(total + 256));
        RC(c->cpay.ensure((total + 32) * sizeof(uint2)));
        CU(cudaMemsetAsync(c->cres.p, kCresNone, total + 256, c->stream));
        CU(cudaMemsetAsync(c->cpay.p, 0, (total + 32) * sizeof(uint2), c->stream));
        chain_compact_kernel<<<blocks, 256, 0, c->stream>>>(tv, (const uint64_t *)pd.p, (const uint64_t *)start.p, (const uint32_t *)len.p,
                                                            (uint8_t *)c->cres.p, (uint2 *)c->cpay.p);
        c->launches += 2;
        unsigned long long roots = 0;
        CU(cudaMemcpyAsync(&roots, (const unsigned long long *)flag.p + 1, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        c->n_chain = (uint32_t)total;
        c->n_chains = roots;
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        c->chain_build_ms = ms;
        return 0;
    };
    rc = body();
    DevBuf *tmp[] = {&pd, &succ, &claim, &len, &start, &flag};
    for (auto b : tmp) b->release();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) {
        c->chain.release();
        c->cpos.release();
        c->cres.release();
        c->cpay.release();
        c->n_chain = 0;
    }
    return rc;
}

// T2: kmer_image.cc:87-105
static int validate_image(const void *image, size_t bytes, const char *name) {
    if (!image || bytes < sizeof(ckm_image_header_t))
        return ckm_fail(CKM_EFORMAT, "Version mismatch for file %s: file size does not match", name);
    const ckm_image_header_t *h = (const ckm_image_header_t *)image;
    // the size check by division: a crafted num_sigs must not wrap sizeof(slot) * num_sigs around to the file size
    const size_t body = bytes - sizeof(ckm_image_header_t);
    if (body % sizeof(ckm_sig_kmer_t) != 0 || (unsigned long long)(body / sizeof(ckm_sig_kmer_t)) != (unsigned long long)h->num_sigs)
        return ckm_fail(CKM_EFORMAT, "Version mismatch for file %s: file size does not match", name);
    if (h->version != 1)
        return ckm_fail(CKM_EFORMAT, "Version mismatch for file %s: file has %lld code has %lld", name, (long long)h->version, 1LL);
    if (h->entry_size != sizeof(ckm_sig_kmer_t))
        return ckm_fail(CKM_EFORMAT, "Version mismatch for file %s: file has entry size %lld code has %lld", name,
                        (long long)h->entry_size, (long long)sizeof(ckm_sig_kmer_t));
    // one bucket: floor(2^64 / 1) does not fit the 64-bit multiplier of fast_mod (ckm_common.cuh); no real image is that small
    if (h->num_sigs < 2) return ckm_fail(CKM_EFORMAT, "image %s has fewer than two buckets", name);
    return 0;
}

static int ctx_create(int device, ckm_ctx **out) {
    RC(select_device(device));
    ckm_ctx *c = new ckm_ctx();
    c->device = device;
    if (const char *pc = getenv("CKM_PIPELINE_CHUNK_KB")) c->pipeline_chunk_bytes = std::max<uint64_t>(1, (uint64_t)atol(pc)) << 10;
    if (const char *pr = getenv("CKM_PIPELINE_RAMP_DIV")) c->pipeline_ramp_div = std::max(1, atoi(pr));
    if (const char *pt = getenv("CKM_PIPELINE_TAIL_DIV")) c->pipeline_tail_div = std::max(1, atoi(pt));
    if (const char *pm = getenv("CKM_PIPELINE_MIN_KB")) c->pipeline_min_bytes = (uint64_t)atol(pm) << 10;
    if (const char *rh = getenv("CKM_REFERENCE_HASH")) c->reference_hash = rh[0] == '1';
    const char *fr = getenv("CKM_FORCE_RAW_SLOTS");
    c->force_raw = fr && fr[0] == '1';
    if (const char *pg = getenv("CKM_PROBE_GROUP")) {
        const int g = atoi(pg);
        if (g == 4 || g == 8 || g == 16 || g == 32) c->probe_group_override = (uint32_t)g;
    }
    c->staged_upload = !getenv("CKM_NO_STAGED_UPLOAD");
    if (check_cuda(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
        delete c;
        return CKM_ECUDA;
    }
    if (cudaFuncSetAttribute(probe_pc_kernel<kPcProducers, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pc_smem_bytes(kPcProducers)) != cudaSuccess) {
        const int rc = ckm_fail(CKM_ECUDA, "cudaFuncSetAttribute(probe_pc_kernel): %s", cudaGetErrorString(cudaGetLastError()));
        cudaStreamDestroy(c->stream);
        delete c;
        return rc;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = prop.l2CacheSize;
    // The probe is a random 16-byte gather: with the default L2 fetch granularity every miss drags a whole
    // 128-byte line out of HBM (measured: 132 B of DRAM reads per probe, profiles/r1a).  Ask for single
    // 32-byte sectors instead.  It is a device-wide hint; CKM_L2_FETCH=64|128 restores coarser fetches.
    {
        size_t gran = 32;
        if (const char *g = getenv("CKM_L2_FETCH")) gran = (size_t)atoi(g);
        if (gran == 32 || gran == 64 || gran == 128) {
            if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) (void)cudaGetLastError();
        }
        size_t got = 0;
        if (cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity) == cudaSuccess) c->l2_fetch = (int)got;
        else (void)cudaGetLastError();
    }
    ckm_set_default_params(c);
    *out = c;
    return 0;
}

extern "C" int ckm_open_image(const void *image, size_t image_bytes, int device, const char *const *function_names,
                              int32_t n_functions, const char *const *otu_names, int32_t n_otus, ckm_ctx **out) {
    if (!out) return ckm_fail(CKM_EINVAL, "out is NULL");
    *out = nullptr;
    RC(validate_image(image, image_bytes, "<memory>"));
    ckm_ctx *c = nullptr;
    RC(ctx_create(device, &c));
    int rc = upload_table(c, (const ckm_image_header_t *)image);
    if (rc) {
        ckm_close(c);
        return rc;
    }
    for (int32_t i = 0; i < n_functions; i++) c->functions.emplace_back(function_names[i]);
    for (int32_t i = 0; i < n_otus; i++) c->otu_names.emplace_back(otu_names[i]);
    *out = c;
    return 0;
}

// load_indexed_ar, kguts.cc:544-575: "<idx>\t<text>\n", dense and in order; the last character of
// every line (the newline fgets leaves) is dropped.
static int load_index_file(const std::string &path, std::vector<std::string> &out) {
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) return ckm_fail(CKM_EIO, "could not open %s", path.c_str());
    char line[1000];
    int j;
    while (fscanf(fp, "%d\t", &j) == 1 && fgets(line, 1000, fp)) {
        if ((int)out.size() != j) {
            fclose(fp);
            return ckm_fail(CKM_EFORMAT, "Your index must be dense and in order (%s, should be %d)", path.c_str(), (int)out.size());
        }
        size_t l = strlen(line);
        out.emplace_back(line, l ? l - 1 : 0);
        if (out.size() >= 1000000) {  // MAX_FUNC_OI_INDEX, kguts.cc:541
            fclose(fp);
            return ckm_fail(CKM_EFORMAT, "index %s has more than 1000000 entries", path.c_str());
        }
    }
    fclose(fp);
    return 0;
}

extern "C" int ckm_open(const char *kmer_dir, int device, ckm_ctx **out) {
    if (!out || !kmer_dir) return ckm_fail(CKM_EINVAL, "NULL argument");
    *out = nullptr;
    const std::string dir(kmer_dir), file = dir + "/kmer.table.mem_map";
    int fd = open(file.c_str(), O_RDONLY);
    if (fd < 0) return ckm_fail(CKM_EIO, "open %s: %s", file.c_str(), strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) < 0) {
        close(fd);
        return ckm_fail(CKM_EIO, "stat %s failed: %s", file.c_str(), strerror(errno));
    }
    void *img = mmap(0, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (img == MAP_FAILED) return ckm_fail(CKM_EIO, "mmap of kmer_table %s failed: %s", file.c_str(), strerror(errno));
    int rc = validate_image(img, (size_t)sb.st_size, file.c_str());
    ckm_ctx *c = nullptr;
    if (!rc) rc = ctx_create(device, &c);
    if (!rc) rc = upload_table(c, (const ckm_image_header_t *)img);
    munmap(img, (size_t)sb.st_size);
    if (!rc) rc = load_index_file(dir + "/function.index", c->functions);
    if (!rc) rc = load_index_file(dir + "/otu.index", c->otu_names);
    if (rc) {
        if (c) ckm_close(c);
        return rc;
    }
    *out = c;
    return 0;
}

// One KmerGuts per worker thread over one shared KmerImage (threadpool.cc:33, kguts.h:312): a second context on the same
// device with its own streams, parameters and work buffers, reading the parent's signature table, occupancy bitmap and
// (if loaded) family tables in place.  The parent must outlive its clones.
extern "C" int ckm_clone(ckm_ctx *parent, ckm_ctx **out) {
    if (!parent || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    *out = nullptr;
    ckm_ctx *c = nullptr;
    RC(ctx_create(parent->device, &c));
    c->shares_tables = true;
    c->table = parent->table;
    c->occupied = parent->occupied;
    c->chain = parent->chain;
    c->cpos = parent->cpos;
    c->cres = parent->cres;
    c->cpay = parent->cpay;
    c->n_chain = parent->n_chain;
    c->n_chains = parent->n_chains;
    c->num_sigs = parent->num_sigs;
    c->image_buckets = parent->image_buckets;
    c->hbits = parent->hbits;
    c->magic = parent->magic;
    c->slot_bytes = parent->slot_bytes;
    c->force_raw = parent->force_raw;
    c->tuning = parent->tuning;
    c->probe_group_override = parent->probe_group_override;
    c->staged_upload = parent->staged_upload;
    c->functions = parent->functions;
    c->otu_names = parent->otu_names;
    c->prm = parent->prm;
    if (parent->has_l2_window) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof attr);
        attr.accessPolicyWindow = parent->l2_window;
        if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) (void)cudaGetLastError();
        c->l2_window = parent->l2_window;
        c->has_l2_window = true;
    }
    ckm_ctx::Family &F = c->fam, &P = parent->fam;
    if (P.loaded) {
        F.loaded = true;
        F.table = P.table;
        F.ids = P.ids;
        F.fam_func = P.fam_func;
        F.fam_pgf = P.fam_pgf;
        F.func_sid = P.func_sid;
        F.mask = P.mask;
        F.n_fams = P.n_fams;
        F.n_functions = P.n_functions;
        F.hypo_sid = P.hypo_sid;
        F.pgf_names = P.pgf_names;
        F.plf = P.plf;
    }
    *out = c;
    return 0;
}

extern "C" void ckm_close(ckm_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    c->free_all();
    if (c->stream_h2d) cudaStreamSynchronize(c->stream_h2d);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_done2) cudaEventDestroy(c->ev_done2);
    for (auto &e : c->ev_copy)
        if (e) cudaEventDestroy(e);
    if (c->stream_h2d) cudaStreamDestroy(c->stream_h2d);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char *ckm_function_at_index(const ckm_ctx *c, int32_t i) {
    if (i < 0 || (size_t)i >= c->functions.size()) return "INVALID_OFFSET";
    return c->functions[i].c_str();
}
extern "C" const char *ckm_otu_at_index(const ckm_ctx *c, int32_t i) {
    if (i < 0 || (size_t)i >= c->otu_names.size()) return "INVALID_OFFSET";
    return c->otu_names[i].c_str();
}
extern "C" int32_t ckm_function_count(const ckm_ctx *c) { return (int32_t)c->functions.size(); }
extern "C" int32_t ckm_otu_count(const ckm_ctx *c) { return (int32_t)c->otu_names.size(); }
extern "C" uint64_t ckm_num_sigs(const ckm_ctx *c) { return c->image_buckets; }
extern "C" uint64_t ckm_table_buckets(const ckm_ctx *c) { return c->num_sigs; }
extern "C" int ckm_table_slot_bytes(const ckm_ctx *c) { return c->slot_bytes; }
extern "C" int ckm_l2_fetch_granularity(const ckm_ctx *c) { return c->l2_fetch; }
extern "C" void ckm_set_tuning(ckm_ctx *c, uint32_t bits) { c->tuning = bits; }
extern "C" int ckm_last_batch_was_fused(const ckm_ctx *c) { return c->last_fused ? 1 : 0; }
extern "C" int ckm_experiments_enabled(void) {
#ifdef CKM_EXPERIMENTS
    return 1;
#else
    return 0;
#endif
}
extern "C" int ckm_has_occupancy_bitmap(const ckm_ctx *c) { return c->occupied.p != nullptr; }
extern "C" int ckm_chain_info(ckm_ctx *c, uint64_t info[4]) {
    if (!c || !info) return ckm_fail(CKM_EINVAL, "NULL argument");
    info[0] = c->n_chain - (uint64_t)kChainPad * c->n_chains;  // k-mers in the copy (every chain is followed by kChainPad unused indices)
    info[1] = c->n_chains;
    info[2] = (uint64_t)(c->chain_build_ms * 1000.0);
    info[3] = 0;
    if (c->totals.p && c->n_chain && c->last_used_copy) {
        CU(cudaMemcpyAsync(&info[3], (const uint64_t *)c->totals.p + 4, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}
extern "C" int ckm_copy_state(const ckm_ctx *c, uint32_t state[3]) {
    if (!c || !state) return ckm_fail(CKM_EINVAL, "NULL argument");
    state[0] = c->copy_suspended ? 1u : 0u;
    state[1] = c->copy_retry_in;
    state[2] = c->copy_suspensions;
    return 0;
}
extern "C" void *ckm_stream(ckm_ctx *c) { return (void *)c->stream; }
extern "C" uint64_t ckm_launch_count(const ckm_ctx *c) { return c->launches; }
extern "C" int ckm_synchronize(ckm_ctx *c) {
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// parameters (Q1)
// ---------------------------------------------------------------------------------------------------
extern "C" void ckm_set_default_params(ckm_ctx *c) {
    c->prm.order_constraint = 0;
    c->prm.min_hits = 5;
    c->prm.min_weighted_hits = 0;
    c->prm.max_gap = 200;
}
extern "C" int ckm_set_params(ckm_ctx *c, int order_constraint, int min_hits, int min_weighted_hits, int max_gap) {
    // min_hits < 1 makes the reference read before its hit buffer (kguts.cc:772 with num_hits == 0):
    // undefined there, rejected here.
    if (min_hits < 1) return ckm_fail(CKM_EINVAL, "min_hits=%d: the reference's behaviour is undefined below 1", min_hits);
    c->prm.order_constraint = order_constraint;
    c->prm.min_hits = min_hits;
    c->prm.min_weighted_hits = min_weighted_hits;
    c->prm.max_gap = max_gap;
    return 0;
}
extern "C" void ckm_get_params(const ckm_ctx *c, int *oc, int *mh, int *mwh, int *mg) {
    if (oc) *oc = c->prm.order_constraint;
    if (mh) *mh = c->prm.min_hits;
    if (mwh) *mwh = c->prm.min_weighted_hits;
    if (mg) *mg = c->prm.max_gap;
}

// ---------------------------------------------------------------------------------------------------
// statics (E1/E2 on the host, for formatters and builders)
// ---------------------------------------------------------------------------------------------------
static const char kProtAlpha[21] = "ACDEFGHIKLMNPQRSTVWY";
static inline int aa_off(char ch) {
    const char *p = ch ? strchr(kProtAlpha, ch) : nullptr;
    return p ? (int)(p - kProtAlpha) : 20;
}
extern "C" uint64_t ckm_encoded_aa_kmer(const char *p) {
    uint64_t k = 0;
    for (int j = 0; j < CKM_KMER_SIZE; j++) {
        const int o = aa_off(p[j]);
        if (o >= 20) return CKM_MAX_ENCODED + 1;
        k = k * 20 + (uint64_t)o;
    }
    return k;
}
extern "C" void ckm_decoded_kmer(uint64_t k, char decoded[9]) {
    decoded[CKM_KMER_SIZE] = 0;
    for (int i = CKM_KMER_SIZE - 1; i >= 0; i--) {
        decoded[i] = kProtAlpha[k % 20];
        k /= 20;
    }
}

// ---------------------------------------------------------------------------------------------------
// image builder: KmerGuts(dir, nbuckets) + insert_kmer + save_kmer_hash_table (kguts.cc:77-115,
// 166-171, 188-234) on the host.  Produces the reference's file bytes; not part of the query path.
// ---------------------------------------------------------------------------------------------------
extern "C" int ckm_image_build(uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *fI, const int32_t *oI,
                               const uint16_t *avg, const float *wt, void *image_out, size_t image_bytes) {
    const size_t need = sizeof(ckm_image_header_t) + (size_t)nbuckets * sizeof(ckm_sig_kmer_t);
    if (!image_out || image_bytes != need) return ckm_fail(CKM_EINVAL, "image buffer must be exactly %zu bytes", need);
    memset(image_out, 0, need);
    ckm_image_header_t *h = (ckm_image_header_t *)image_out;
    h->num_sigs = nbuckets;
    h->entry_size = sizeof(ckm_sig_kmer_t);
    h->version = 1;
    ckm_sig_kmer_t *s = (ckm_sig_kmer_t *)(h + 1);
    for (uint64_t i = 0; i < nbuckets; i++) s[i].which_kmer = CKM_MAX_ENCODED + 1;
    uint64_t loaded = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (keys[i] > CKM_MAX_ENCODED) continue;  // kguts.cc:206-210
        uint64_t e = keys[i] % nbuckets;
        while (s[e].which_kmer <= CKM_MAX_ENCODED) e = (e + 1) % nbuckets;
        loaded++;
        if ((long long)loaded >= (long long)nbuckets / 2)  // kguts.cc:213-216 (the reference exits)
            return ckm_fail(CKM_EINVAL, "Your Kmer hash is half-full; use a larger bucket count");
        s[e].which_kmer = keys[i];
        s[e].avg_from_end = avg[i];
        s[e].function_index = fI[i];
        s[e].otu_index = oI[i];
        s[e].function_wt = wt[i];
    }
    return 0;
}

extern "C" int ckm_host_alloc(void **p, size_t bytes) {
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess) return ckm_fail(CKM_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return 0;
}
extern "C" void ckm_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------
// the calling path
// ---------------------------------------------------------------------------------------------------
static int prefix_sum(ckm_ctx *c, const uint32_t *d_in, uint64_t n, uint64_t *d_out /* n+1 */) {
    const uint64_t nb = std::max<uint64_t>(1, (n + kPsTile - 1) / kPsTile);
    RC(c->ps_blocks.ensure(nb * 8));
    ps_block_sums<<<(unsigned)nb, kPsThreads, 0, c->stream>>>(d_in, n, (uint64_t *)c->ps_blocks.p);
    ps_spine<<<1, kPsThreads, 0, c->stream>>>((uint64_t *)c->ps_blocks.p, nb);
    ps_finish<<<(unsigned)nb, kPsThreads, 0, c->stream>>>(d_in, n, (const uint64_t *)c->ps_blocks.p, d_out);
    c->launches += 3;
    return 0;
}

constexpr uint32_t kWorkSlots = 64;  // work counters of probe_pc_kernel launches in flight (two streams, a handful each)

struct RunPlan {
    bool want_scan, general, want_keys, want_avg;
    bool fused;            // the scoring scan runs inside K1 (ckm_warp_scan.cuh): no hit records, no scan_kernel
    uint32_t probe_group;  // lanes per sequence in K1: 32 (probe_kernel) or 4 / 8 / 16 (probe_group_kernel, short sequences)
};

// the fused K1 serves requests for calls and / or the best call of proteins that cannot saturate the hit window
static bool plan_fused(const ckm_ctx *c, const RunPlan &plan, uint32_t flags) {
    return plan.want_scan && !plan.general && !(flags & (CKM_WANT_HITS | CKM_WANT_OTU)) && plan.probe_group >= 32u &&
           c->slot_bytes == kPackedSlotBytes && c->num_sigs < 0xFFFFFFF0ull && !(c->tuning & CKM_TUNE_UNFUSED);
}

// size the per-batch regions (indexed by residue offset / sequence index, so chunked launches can share them)
static int prepare_regions(ckm_ctx *c, uint32_t n, uint64_t total, uint32_t max_len, uint32_t flags, RunPlan *plan) {
    c->cur_n = n;
    c->cur_total = total;
    c->cur_flags = flags;
    plan->want_scan = flags & (CKM_WANT_CALLS | CKM_WANT_OTU | CKM_WANT_BEST);
    {
        // by mean length: a group step covers 4*G start positions
        const uint64_t mean = n ? total / n : 0;
        plan->probe_group = n < 64 ? 32u : mean <= 24 ? 4u : mean <= 48 ? 8u : mean <= 96 ? 16u : 32u;
        if (c->probe_group_override) plan->probe_group = c->probe_group_override;
    }
    plan->general = plan->want_scan && (c->prm.order_constraint != 0 || max_len == 0 || max_len > kHitCap + CKM_KMER_SIZE);
    plan->want_keys = flags & CKM_WANT_HITS;
    plan->want_avg = (flags & CKM_WANT_HITS) || (plan->want_scan && c->prm.order_constraint != 0);
    plan->fused = plan_fused(c, *plan, flags);
    const uint64_t ncall_slots = total / (uint64_t)std::max(1, c->prm.min_hits) + n + 1;
    RC(c->totals.ensure(64));
    if (!plan->fused) RC(c->hits.ensure((total + 1) * sizeof(HitRec)));
    RC(c->n_hits.ensure(((size_t)n + 1) * 4));
    if (c->n_chain && plan->probe_group >= 32u) RC(c->hints.ensure(((total >> kHintShift) + n + 2) * 4));
    if (plan->want_keys) RC(c->hit_keys.ensure((total + 1) * 8));
    if (plan->want_avg) RC(c->hit_avg.ensure((total + 1) * 2));
    if (plan->want_scan) {
        RC(c->calls.ensure(ncall_slots * sizeof(ckm_call_t)));
        RC(c->n_calls.ensure(((size_t)n + 1) * 4));
        if (plan->general) RC(c->stored_idx.ensure((total + 1) * 4));
        if (flags & CKM_WANT_BEST) {
            RC(c->calls_work.ensure(ncall_slots * sizeof(ckm_call_t)));
            RC(c->best.ensure(((size_t)n + 1) * sizeof(ckm_best_t)));
        }
        if (flags & CKM_WANT_OTU) {
            RC(c->otus.ensure((total + 1) * sizeof(ckm_otu_t)));
            RC(c->n_otus.ensure(((size_t)n + 1) * 4));
        }
    }
    return 0;
}

static TableView table_view(const ckm_ctx *c) {
    TableView tv;
    tv.slots = c->table.p;
    tv.num_sigs = c->num_sigs;
    tv.magic = c->magic;
    tv.occupied = (const uint32_t *)c->occupied.p;
    tv.tuning = c->tuning;
    tv.chain = (const uint4 *)c->chain.p;
    tv.cpos = (const uint2 *)c->cpos.p;
    tv.cres = (const uint8_t *)c->cres.p;
    tv.cpay = (const uint2 *)c->cpay.p;
    tv.n_chain = c->n_chain;
    tv.m35 = magic35(c->num_sigs);
    tv.hbits = c->hbits;
    return tv;
}

// does K1 go through the neighbour copy (hint_kernel + probe_hint_kernel) for this ctx at present?
static bool use_neighbour_copy(const ckm_ctx *c) {
    return c->slot_bytes == kPackedSlotBytes && c->n_chain && !(c->tuning & CKM_TUNE_PLAIN_PROBE) && !c->copy_suspended;
}

// Automatic fall-back from the neighbour copy to plain hash probing.  The copy pays when a good part of a batch's probes is
// answered from it (dense signature sets: 0.59 of C2's probes, K1 4.0 ms against 6.7); when the signatures are a sparse
// subset of the windows the chains are a k-mer or two long, hints lead nowhere and the hinted step only adds work
// (profiles/r2/sparse_worlds.jsonl: break-even at ~0.12 of the probes).  So every batch that went through the copy and whose
// counters reach the host is judged: below kCopyMinShare the copy is suspended for `backoff` batches (16, doubling up to
// 1024 while it keeps failing), then tried again on one batch.  Results never depend on the path; CKM_TUNE_NO_FALLBACK pins it.
constexpr uint64_t kCopyJudgeMinProbes = 200000;  // smaller batches say too little
constexpr double kCopyMinShare = 0.12;
static void adapt_probe_path(ckm_ctx *c, uint64_t probes, uint64_t from_copy) {
    if (!c->n_chain || (c->tuning & (CKM_TUNE_NO_FALLBACK | CKM_TUNE_PLAIN_PROBE))) return;
    if (c->copy_suspended) {
        if (c->copy_retry_in == 0 || --c->copy_retry_in == 0) c->copy_suspended = false;  // the next batch is the retry
        return;
    }
    if (!c->last_used_copy || probes < kCopyJudgeMinProbes) return;
    if ((double)from_copy < kCopyMinShare * (double)probes) {
        c->copy_suspended = true;
        c->copy_retry_in = c->copy_backoff;
        c->copy_backoff = std::min<uint32_t>(c->copy_backoff * 2, 1024);
        c->copy_suspensions++;
    } else {
        c->copy_backoff = 16;
    }
}

// K1 (+ K2) for sequences [i0, i0+cnt) of the batch on `stream`
static int launch_range(ckm_ctx *c, cudaStream_t stream, const uint8_t *d_res, const uint64_t *d_off, uint32_t i0, uint32_t cnt,
                        uint32_t flags, const RunPlan &plan, ckm_ctx::ProfEv *pe) {
    if (cnt == 0) return 0;
    c->last_fused = plan.fused;
    c->last_used_copy = plan.probe_group >= 32u && use_neighbour_copy(c);
    const TableView tv = table_view(c);
    const bool packed = c->slot_bytes == kPackedSlotBytes;
    FusedArgs fa;
    memset(&fa, 0, sizeof fa);
    if (plan.fused) {
        fa.calls = (ckm_call_t *)c->calls.p;
        fa.calls_work = (flags & CKM_WANT_BEST) ? (ckm_call_t *)c->calls_work.p : nullptr;
        fa.n_calls = (uint32_t *)c->n_calls.p + i0;
        fa.best = (flags & CKM_WANT_BEST) ? (ckm_best_t *)c->best.p + i0 : nullptr;
        fa.prm = c->prm;
        fa.call_magic = call_region_magic(c->prm.min_hits);
    }
    {
        const uint32_t warps_per_block = kProbeThreads / 32;
        uint64_t blocks = ((uint64_t)cnt + warps_per_block - 1) / warps_per_block;
        // 64 blocks per SM in the grid (3 are resident at 80 registers): a finer grid than the residency evens out the
        // tail between long and short proteins (measured: 7.5 -> 7.1 ms on C2, profiles/r1/tune3.jsonl)
        uint32_t bps = 8, gshift = 3;
#ifdef CKM_EXPERIMENTS
        if ((c->tuning >> 8) & 0xFu) bps = (c->tuning >> 8) & 0xFu;
        if ((c->tuning >> 12) & 0xFu) gshift = ((c->tuning >> 12) & 0xFu) - 1;
#endif
        blocks = std::min<uint64_t>(blocks, ((uint64_t)c->sm_count * bps) << gshift);
        HitRec *hp = (HitRec *)c->hits.p;
        uint64_t *keys = plan.want_keys ? (uint64_t *)c->hit_keys.p : nullptr;
        uint16_t *avg = plan.want_avg ? (uint16_t *)c->hit_avg.p : nullptr;
        uint32_t *nh = (uint32_t *)c->n_hits.p + i0;
        unsigned long long *tot = (unsigned long long *)c->totals.p;
        // short sequences (fastq fragments, peptides): a group of 4, 8 or 16 lanes per sequence instead of a warp
        const uint32_t group = plan.probe_group;
        if (group < 32u) {
            const uint32_t per_block = warps_per_block * (32u / group);
            const unsigned gb = (unsigned)std::min<uint64_t>(((uint64_t)cnt + per_block - 1) / per_block, (uint64_t)c->sm_count * 64);
            if (packed && group == 4u) probe_group_kernel<true, 4><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
            else if (group == 4u) probe_group_kernel<false, 4><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
            else if (packed && group == 8u) probe_group_kernel<true, 8><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
            else if (packed) probe_group_kernel<true, 16><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
            else if (group == 8u) probe_group_kernel<false, 8><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
            else probe_group_kernel<false, 16><<<gb, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
        } else if (plan.fused) {
            // probing warps + one scan warp per block (ckm_pc.cuh); through the neighbour copy when there is one
            const uint32_t *hints = nullptr;
            if (use_neighbour_copy(c)) {
                const uint64_t per_block = 256 / kHintLanes;
                const unsigned hb = (unsigned)std::min<uint64_t>(((uint64_t)cnt + per_block - 1) / per_block, (uint64_t)c->sm_count * 64);
                hint_kernel<<<hb, 256, 0, stream>>>(tv, d_res, d_off + i0, cnt, i0, (uint32_t *)c->hints.p);
                c->launches++;
                hints = (const uint32_t *)c->hints.p;
            }
            RC(c->work.ensure(kWorkSlots * 8));
            unsigned long long *work = (unsigned long long *)c->work.p + (c->pc_seq++ % kWorkSlots);
            CU(cudaMemsetAsync(work, 0, 8, stream));
            {
                constexpr int P = kPcProducers;
                const unsigned gb = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->sm_count, ((uint64_t)cnt + P * kPcChunk - 1) / (P * kPcChunk)));
                probe_pc_kernel<P, 1><<<gb, (P + 1) * 32, pc_smem_bytes(P), stream>>>(tv, d_res, d_off + i0, cnt, i0, hints, nh, tot, fa, work);
            }
            if (fa.best) {  // find_best_call of the proteins with several calls
                c->launches++;
                best_fixup_kernel<<<(cnt + 255) / 256, 256, 0, stream>>>(d_off + i0, cnt, i0, fa);
            }
        } else if (use_neighbour_copy(c)) {
            bool done = false;
#ifdef CKM_EXPERIMENTS
            if (c->tuning & (64u | 128u)) {  // the walking probe_chain_kernel at 2 / 3 blocks per SM
                if (c->tuning & 64u) probe_chain_kernel<2><<<(unsigned)blocks, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
                else probe_chain_kernel<3><<<(unsigned)blocks, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
                done = true;
            }
#endif
            if (!done) {  // hints first (one sample window in 64), then the probe proper (ckm_hint.cuh)
                const uint64_t per_block = 256 / kHintLanes;
                const unsigned hb = (unsigned)std::min<uint64_t>(((uint64_t)cnt + per_block - 1) / per_block, (uint64_t)c->sm_count * 64);
                hint_kernel<<<hb, 256, 0, stream>>>(tv, d_res, d_off + i0, cnt, i0, (uint32_t *)c->hints.p);
                c->launches++;
#define CKM_HINT_LAUNCH(T, B, S)                                                                                                      \
    probe_hint_kernel<T, B, S><<<(unsigned)std::min<uint64_t>(((uint64_t)cnt + T / 32 - 1) / (T / 32), blocks * (kProbeThreads / T)), T, 0, \
                              stream>>>(tv, d_res, d_off + i0, cnt, i0, (const uint32_t *)c->hints.p, hp, keys, avg, nh, tot)
#ifdef CKM_EXPERIMENTS
                const uint32_t variant = (c->tuning >> 16) & 7u;  // A/B: block shape, blocks per SM, hit payload staged or in registers
                if (variant == 1u) CKM_HINT_LAUNCH(256, 4, true);
                else if (variant == 2u) CKM_HINT_LAUNCH(128, 8, true);
                else if (variant == 3u) CKM_HINT_LAUNCH(128, 6, true);
                else if (variant == 4u) CKM_HINT_LAUNCH(256, 3, false);
                else if (variant == 5u) CKM_HINT_LAUNCH(256, 3, true);
                else
#endif
                CKM_HINT_LAUNCH(128, 7, true);  // 72 registers, 28 warps per SM (profiles/r1/tune_hint_v13_1M.jsonl)
#undef CKM_HINT_LAUNCH
            }
        } else if (packed) {
            probe_kernel<true><<<(unsigned)blocks, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
        } else {
            probe_kernel<false><<<(unsigned)blocks, kProbeThreads, 0, stream>>>(tv, d_res, d_off + i0, cnt, hp, keys, avg, nh, tot);
        }
        c->launches++;
    }
    if (pe) CU(cudaEventRecord(pe->e1, stream));
    if (plan.want_scan && !plan.fused) {
        ScanArgs a;
        a.offsets = d_off + i0;
        a.hits = (const HitRec *)c->hits.p;
        a.hit_avg = (c->prm.order_constraint != 0) ? (const uint16_t *)c->hit_avg.p : nullptr;
        a.n_hits = (const uint32_t *)c->n_hits.p + i0;
        a.stored_idx = plan.general ? (uint32_t *)c->stored_idx.p : nullptr;
        a.calls = (ckm_call_t *)c->calls.p;
        a.calls_work = (flags & CKM_WANT_BEST) ? (ckm_call_t *)c->calls_work.p : nullptr;
        a.n_calls = (uint32_t *)c->n_calls.p + i0;
        a.otus = (flags & CKM_WANT_OTU) ? (ckm_otu_t *)c->otus.p : nullptr;
        a.n_otus = (flags & CKM_WANT_OTU) ? (uint32_t *)c->n_otus.p + i0 : nullptr;
        a.best = (flags & CKM_WANT_BEST) ? (ckm_best_t *)c->best.p + i0 : nullptr;
        a.totals = (unsigned long long *)c->totals.p;
        a.n = cnt;
        a.index_base = i0;
        a.prm = c->prm;
        const unsigned blocks = (cnt + kScanThreads - 1) / kScanThreads;
        if (plan.general)
            scan_kernel<true><<<blocks, kScanThreads, 0, stream>>>(a);
        else
            scan_kernel<false><<<blocks, kScanThreads, 0, stream>>>(a);
        c->launches++;
    }
    return 0;
}

// K1 + K2 over a batch that is already in HBM.  Leaves regions / per-protein counters on the device.
static int run_device(ckm_ctx *c, const uint8_t *d_res, const uint64_t *d_off, uint32_t n, uint64_t total, uint32_t max_len,
                      uint32_t flags) {
    RunPlan plan;
    RC(prepare_regions(c, n, total, max_len, flags, &plan));
    c->cur_off = d_off;
    CU(cudaMemsetAsync(c->totals.p, 0, 64, c->stream));
    if (n == 0) return 0;
    ckm_ctx::ProfEv pe = {nullptr, nullptr, nullptr, false};
    if (c->profiling) {
        CU(cudaEventCreate(&pe.e0));
        CU(cudaEventCreate(&pe.e1));
        CU(cudaEventCreate(&pe.e2));
        CU(cudaEventRecord(pe.e0, c->stream));
    }
    RC(launch_range(c, c->stream, d_res, d_off, 0, n, flags, plan, c->profiling ? &pe : nullptr));
    if (c->profiling) {
        CU(cudaEventRecord(pe.e2, c->stream));
        pe.has_scan = plan.want_scan;
        c->prof.push_back(pe);
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" void ckm_profile_enable(ckm_ctx *c, int on) { c->profiling = on != 0; }

extern "C" int ckm_profile_read(ckm_ctx *c, double *probe_ms, double *scan_ms, uint64_t *batches) {
    CU(cudaStreamSynchronize(c->stream));
    double p = 0, s = 0;
    for (auto &e : c->prof) {
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, e.e0, e.e1));
        CU(cudaEventElapsedTime(&b, e.e1, e.e2));
        p += a;
        if (e.has_scan) s += b;
        cudaEventDestroy(e.e0);
        cudaEventDestroy(e.e1);
        cudaEventDestroy(e.e2);
    }
    if (probe_ms) *probe_ms = p;
    if (scan_ms) *scan_ms = s;
    if (batches) *batches = c->prof.size();
    c->prof.clear();
    return 0;
}

extern "C" int ckm_call_batch_device(ckm_ctx *c, const void *d_residues, const uint64_t *d_offsets, uint32_t n,
                                     uint64_t total_residues, uint32_t max_len, uint32_t flags) {
    if (!c) return ckm_fail(CKM_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    return run_device(c, (const uint8_t *)d_residues, d_offsets, n, total_residues, max_len, flags);
}

extern "C" int ckm_device_results(ckm_ctx *c, ckm_device_out_t *out) {
    if (!c || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    memset(out, 0, sizeof *out);
    out->n = c->cur_n;
    out->d_n_hits = (const uint32_t *)c->n_hits.p;
    out->d_n_calls = (const uint32_t *)c->n_calls.p;
    out->d_calls = (const ckm_call_t *)c->calls.p;
    out->d_best = (const ckm_best_t *)c->best.p;
    out->d_totals = (const uint64_t *)c->totals.p;
    out->min_hits_for_call_base = c->prm.min_hits;
    return 0;
}

extern "C" int ckm_read_totals(ckm_ctx *c, uint64_t totals[3]) {
    if (!c || !totals) return ckm_fail(CKM_EINVAL, "NULL argument");
    if (!c->totals.p) return ckm_fail(CKM_ESTATE, "no batch has been run on this ctx");
    uint64_t t[8];
    CU(cudaMemcpyAsync(t, c->totals.p, 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (t[7]) return ckm_fail(CKM_ECUDA, "probe_pc_kernel: a hand-off between probing and scan warps timed out");
    if (t[6]) return ckm_fail(CKM_EINVAL, "a sequence is longer than the max_len the batch was announced with");
    adapt_probe_path(c, t[0], t[4]);
    totals[0] = t[0];
    totals[1] = t[1];
    totals[2] = t[2];
    return 0;
}

// H2D of one batch: residues (+32 zeroed bytes of slack) and offsets rebased to 0
// Host-to-device copy of a large PAGEABLE buffer.  cudaMemcpyAsync from pageable memory is staged by the driver on the
// calling thread at ~10 GB/s; here kStageThreads workers copy alternate 8 MB chunks into their own page-locked buffers and
// queue the DMA behind them, so the host-side copies run in parallel and overlap the transfers.  ctx->stream is made to
// wait for every chunk.  (Page-locked caller memory goes straight to cudaMemcpyAsync.)
static int staged_upload(ckm_ctx *c, void *dst, const char *src, size_t bytes) {
    constexpr int T = ckm_ctx::kStageThreads;
    constexpr size_t CH = ckm_ctx::kStageChunk;
    for (int t = 0; t < T; t++) {
        if (!c->stage_stream[t]) CU(cudaStreamCreateWithFlags(&c->stage_stream[t], cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            RC(c->stage_buf[t][k].ensure(CH));
            if (!c->stage_ev[t][k]) CU(cudaEventCreateWithFlags(&c->stage_ev[t][k], cudaEventDisableTiming));
        }
    }
    // the destination may still be read by work queued on ctx->stream
    if (!c->ev_ready) CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_ready, c->stream));
    const size_t n_chunks = (bytes + CH - 1) / CH;
    std::atomic<int> failed{0};
    auto worker = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed = 1; return; }
        if (cudaStreamWaitEvent(c->stage_stream[t], c->ev_ready, 0) != cudaSuccess) { failed = 1; return; }
        int k = 0;
        for (size_t i = (size_t)t; i < n_chunks; i += T, k ^= 1) {
            const size_t o = i * CH, len = std::min(CH, bytes - o);
            if (cudaEventSynchronize(c->stage_ev[t][k]) != cudaSuccess) { failed = 1; return; }  // buffer k free again
            memcpy(c->stage_buf[t][k].p, src + o, len);
            if (cudaMemcpyAsync((char *)dst + o, c->stage_buf[t][k].p, len, cudaMemcpyHostToDevice, c->stage_stream[t]) != cudaSuccess ||
                cudaEventRecord(c->stage_ev[t][k], c->stage_stream[t]) != cudaSuccess) { failed = 1; return; }
        }
    };
    std::thread th[T];
    for (int t = 1; t < T; t++) th[t] = std::thread(worker, t);
    worker(0);
    for (int t = 1; t < T; t++) th[t].join();
    if (failed) {
        (void)cudaGetLastError();
        return ckm_fail(CKM_ECUDA, "staged host-to-device copy failed");
    }
    for (int t = 0; t < T; t++)
        for (int k = 0; k < 2; k++) CU(cudaStreamWaitEvent(c->stream, c->stage_ev[t][k], 0));
    return 0;
}

static bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static int upload_batch(ckm_ctx *c, const char *residues, const uint64_t *offsets, uint32_t n, uint64_t *total_out,
                        uint32_t *max_len_out) {
    if (!offsets || (n && !residues && offsets[n] != offsets[0])) return ckm_fail(CKM_EINVAL, "NULL argument");
    CU(cudaSetDevice(c->device));
    uint32_t max_len = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return ckm_fail(CKM_EINVAL, "offsets must be non-decreasing (at %u)", i);
        const uint64_t l = offsets[i + 1] - offsets[i];
        if (l > 500000000ull) return ckm_fail(CKM_EINVAL, "sequence %u longer than MAX_SEQ_LEN", i);  // kmer_params.h:6
        max_len = std::max<uint32_t>(max_len, (uint32_t)l);
    }
    const uint64_t total = offsets[n] - offsets[0];
    RC(c->in_res.ensure(total + 32));
    RC(c->in_off.ensure(((size_t)n + 1) * 8));
    if (total >= (32u << 20) && c->staged_upload && is_pageable(residues + offsets[0]))
        RC(staged_upload(c, c->in_res.p, residues + offsets[0], total));
    else if (total)
        CU(cudaMemcpyAsync(c->in_res.p, residues + offsets[0], total, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync((uint8_t *)c->in_res.p + total, 0, 32, c->stream));
    const uint64_t *h_off = offsets;
    if (offsets[0] != 0) {
        RC(c->h_off.ensure(((size_t)n + 1) * 8));
        uint64_t *t = (uint64_t *)c->h_off.p;
        for (uint32_t i = 0; i <= n; i++) t[i] = offsets[i] - offsets[0];
        h_off = t;
    }
    CU(cudaMemcpyAsync(c->in_off.p, h_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    *total_out = total;
    *max_len_out = max_len;
    return 0;
}

// error exit of the pipelined path after work was enqueued: nothing of the aborted batch may still run when the next one starts
static int pipeline_abort(ckm_ctx *c, int rc) {
    if (c->stream_h2d) cudaStreamSynchronize(c->stream_h2d);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream);
    (void)cudaGetLastError();
    return rc;
}

// find_best_call-only batches (fixed-size results) are streamed: the batch is cut into chunks of about
// pipeline_chunk_bytes residues that alternate between two CUDA streams, so that the H2D copy of chunk k+1 and the
// D2H copy of chunk k-1 overlap the kernels of chunk k.  All chunks share the per-batch regions (indexed by residue
// offset / sequence index, hence disjoint), so nothing is double-buffered.
// `packed` != NULL: `offsets` are word offsets into the packed words (ckm_packed.cuh); every chunk is unpacked on the
// device behind its copy, and one sequence takes 8 residue slots per word there.
static int call_batch_pipelined(ckm_ctx *c, const char *residues, const uint32_t *packed, const uint64_t *offsets, uint32_t n,
                                ckm_batch_out_t *out) {
    const uint64_t unit = packed ? 8ull : 1ull;  // device residue slots per unit of `offsets`
    CU(cudaSetDevice(c->device));
    const uint64_t base0 = offsets[0];
    if (offsets[n] < base0) return ckm_fail(CKM_EINVAL, "offsets must be non-decreasing");
    const uint64_t total = (offsets[n] - base0) * unit;
    RunPlan plan;
    // max_len is only known chunk by chunk (validation is overlapped with the copies): size for the general kernel
    // lazily, below, if a chunk turns out to need it
    RC(prepare_regions(c, n, total, 1u, CKM_WANT_BEST, &plan));
    RC(c->in_res.ensure(total + 32));
    RC(c->in_off.ensure(((size_t)n + 1) * 8));
    if (packed) {
        RC(c->in_packed.ensure((offsets[n] - base0 + 2) * 4));
        RC(c->in_woff.ensure(((size_t)n + 1) * 8));
    }
    RC(c->h_best.ensure(((size_t)n + 1) * sizeof(ckm_best_t)));
    RC(c->h_totals.ensure(64));
    if (!c->stream2) {
        CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        if (c->has_l2_window) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow = c->l2_window;
            if (cudaStreamSetAttribute(c->stream2, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) (void)cudaGetLastError();
        }
    }
    if (!c->ev_ready) CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    if (!c->ev_done2) CU(cudaEventCreateWithFlags(&c->ev_done2, cudaEventDisableTiming));
    // The copies of all chunks go down a stream of their own, back to back: on the compute streams a chunk's copy would sit behind
    // the kernels of the chunk two before it (a stream is in-order) although the copy engine is idle.  A chunk's kernels wait for
    // its event.
    if (!c->stream_h2d) CU(cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking));
    for (auto &e : c->ev_copy)
        if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->cur_off = (const uint64_t *)c->in_off.p;
    // stream 0: counters and slack; stream 1 waits for them.  Offsets travel with their chunk.
    CU(cudaMemsetAsync(c->totals.p, 0, 64, c->stream));
    CU(cudaMemsetAsync((uint8_t *)c->in_res.p + total, 0, 32, c->stream));
    CU(cudaEventRecord(c->ev_ready, c->stream));
    CU(cudaStreamWaitEvent(c->stream2, c->ev_ready, 0));
    CU(cudaStreamWaitEvent(c->stream_h2d, c->ev_ready, 0));  // (whatever the context's stream still had queued reads the old input)
    uint64_t *reb = nullptr;
    if (base0 != 0) {
        RC(c->h_off.ensure(((size_t)n + 1) * 8));
        reb = (uint64_t *)c->h_off.p;
        reb[0] = 0;
    }
    // chunk schedule: ramp up from 1/6 of the chunk size so the first kernels start early, and end on a short chunk
    // so little compute is left once the last copy lands
    const uint64_t cap = std::max<uint64_t>(c->pipeline_chunk_bytes, 1024), tail = cap / c->pipeline_tail_div;
    uint64_t want = std::max<uint64_t>(cap / c->pipeline_ramp_div, 1024);
    uint32_t i0 = 0;
    int k = 0;
    while (i0 < n) {
        const uint64_t start = (offsets[i0] - base0) * unit, left = total - start;
        uint64_t goal = want;
        if (left <= goal + tail) goal = left > 2 * tail ? left - tail : left;
        uint32_t i1 = i0, max_len = 0;
        while (i1 < n) {  // validate and size the chunk in one pass over its offsets
            if (offsets[i1 + 1] < offsets[i1]) return pipeline_abort(c, ckm_fail(CKM_EINVAL, "offsets must be non-decreasing (at %u)", i1));
            const uint64_t l = (offsets[i1 + 1] - offsets[i1]) * unit;
            if (l > 500000000ull) return pipeline_abort(c, ckm_fail(CKM_EINVAL, "sequence %u longer than MAX_SEQ_LEN", i1));
            if (i1 > i0 && (offsets[i1 + 1] - base0) * unit - start > goal) break;
            max_len = std::max<uint32_t>(max_len, (uint32_t)l);
            if (reb) reb[i1 + 1] = offsets[i1 + 1] - base0;
            i1++;
        }
        const uint64_t bytes = (offsets[i1] - base0) * unit - start;
        RunPlan cp = plan;
        cp.general = max_len > kHitCap + CKM_KMER_SIZE || c->prm.order_constraint != 0;
        cp.fused = plan_fused(c, cp, CKM_WANT_BEST);
        if (cp.general) {
            if (c->stored_idx.cap < (total + 1) * 4) {  // first general chunk: nothing in flight uses this buffer yet
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaStreamSynchronize(c->stream2));
                RC(c->stored_idx.ensure((total + 1) * 4));
            }
            if (c->hits.cap < (total + 1) * sizeof(HitRec)) {  // the fused chunks before this one left no hit regions
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaStreamSynchronize(c->stream2));
                RC(c->hits.ensure((total + 1) * sizeof(HitRec)));
            }
            cp.want_avg = c->prm.order_constraint != 0;
            if (cp.want_avg && c->hit_avg.cap < (total + 1) * 2) {
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaStreamSynchronize(c->stream2));
                RC(c->hit_avg.ensure((total + 1) * 2));
            }
        }
        cudaStream_t st = (k & 1) ? c->stream2 : c->stream;
        cudaStream_t sc = c->stream_h2d;
        cudaEvent_t landed = c->ev_copy[k % ckm_ctx::kCopyEvents];
        if (k >= ckm_ctx::kCopyEvents) CU(cudaStreamSynchronize(k & 1 ? c->stream2 : c->stream));  // the event's last user has passed it
        const uint64_t *src_off = reb ? reb : offsets;
        if (packed) {
            const uint64_t w0 = offsets[i0] - base0, nw = offsets[i1] - offsets[i0];
            CU(cudaMemcpyAsync((uint64_t *)c->in_woff.p + i0, src_off + i0, ((size_t)(i1 - i0) + 1) * 8, cudaMemcpyHostToDevice, sc));
            if (nw) CU(cudaMemcpyAsync((uint32_t *)c->in_packed.p + w0, packed + base0 + w0, nw * 4, cudaMemcpyHostToDevice, sc));
            CU(cudaEventRecord(landed, sc));
            CU(cudaStreamWaitEvent(st, landed, 0));
            const unsigned ub = (unsigned)std::min<uint64_t>(((uint64_t)(i1 - i0) + 1 + 7) / 8, (uint64_t)c->sm_count * 32);
            unpack7_kernel<<<ub, kUnpackThreads, 0, st>>>((const uint32_t *)c->in_packed.p, (const uint64_t *)c->in_woff.p + i0, i1 - i0, 0ull,
                                               (uint8_t *)c->in_res.p, (uint64_t *)c->in_off.p + i0);
            c->launches++;
        } else {
            CU(cudaMemcpyAsync((uint64_t *)c->in_off.p + i0, src_off + i0, ((size_t)(i1 - i0) + 1) * 8, cudaMemcpyHostToDevice, sc));
            if (bytes) CU(cudaMemcpyAsync((uint8_t *)c->in_res.p + start, residues + base0 + start, bytes, cudaMemcpyHostToDevice, sc));
            CU(cudaEventRecord(landed, sc));
            CU(cudaStreamWaitEvent(st, landed, 0));
        }
        RC(launch_range(c, st, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, i0, i1 - i0, CKM_WANT_BEST, cp, nullptr));
        CU(cudaMemcpyAsync((ckm_best_t *)c->h_best.p + i0, (const ckm_best_t *)c->best.p + i0, (size_t)(i1 - i0) * sizeof(ckm_best_t),
                           cudaMemcpyDeviceToHost, st));
        i0 = i1;
        want = std::min(cap, want * 2);
        k++;
    }
    // join stream 1 into stream 0, then read the batch counters
    CU(cudaEventRecord(c->ev_done2, c->stream2));
    CU(cudaStreamWaitEvent(c->stream, c->ev_done2, 0));
    uint64_t *ht = (uint64_t *)c->h_totals.p;
    CU(cudaMemcpyAsync(ht, c->totals.p, 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (ht[7]) return ckm_fail(CKM_ECUDA, "probe_pc_kernel: a hand-off between probing and scan warps timed out");
    if (ht[6]) return ckm_fail(CKM_EINVAL, "a sequence is longer than the max_len the batch was announced with");
    adapt_probe_path(c, ht[0], ht[4]);
    out->n_probes = ht[0];
    out->n_hits = ht[1];
    out->best = (const ckm_best_t *)c->h_best.p;
    return 0;
}

static int finish_batch(ckm_ctx *c, uint32_t n, uint32_t flags, ckm_batch_out_t *out);

extern "C" int ckm_call_batch(ckm_ctx *c, const char *residues, const uint64_t *offsets, uint32_t n, uint32_t flags,
                              ckm_batch_out_t *out) {
    if (!c || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    memset(out, 0, sizeof *out);
    out->n = n;
    uint64_t total = 0;
    uint32_t max_len = 0;
    if (flags == CKM_WANT_BEST && offsets && n > 1 && offsets[n] - offsets[0] >= c->pipeline_min_bytes)
        return call_batch_pipelined(c, residues, nullptr, offsets, n, out);
    RC(upload_batch(c, residues, offsets, n, &total, &max_len));
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(max_len, 1u), flags));
    return finish_batch(c, n, flags, out);
}

// ---- packed input: seven residues per 32-bit word, base 22 (ckm_packed.cuh) ----
extern "C" uint64_t ckm_packed_words(uint64_t n_residues) { return (n_residues + kPackPerWord - 1) / kPackPerWord; }

extern "C" int ckm_pack_residues(const char *residues, const uint64_t *offsets, uint32_t n, uint32_t *packed, uint64_t packed_capacity_words,
                                 uint64_t *word_offsets) {
    if (!offsets || !word_offsets || (n && !residues && offsets[n] != offsets[0])) return ckm_fail(CKM_EINVAL, "NULL argument");
    uint8_t code[256];
    for (int ch = 0; ch < 256; ch++) code[ch] = (uint8_t)kPackInvalid;
    for (int k = 0; k < 20; k++) code[(unsigned char)kProtAlpha[k]] = (uint8_t)k;
    code[0] = (uint8_t)kPackEnd;  // an embedded NUL ends the sequence (strlen, kguts.cc:791)
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return ckm_fail(CKM_EINVAL, "offsets must be non-decreasing (at %u)", i);
        const uint64_t len = offsets[i + 1] - offsets[i], words = ckm_packed_words(len);
        word_offsets[i] = w;
        if (w + words > packed_capacity_words) return ckm_fail(CKM_EINVAL, "packed buffer too small (%llu words needed so far)", (unsigned long long)(w + words));
        const unsigned char *p = (const unsigned char *)residues + offsets[i];
        uint32_t *dst = packed + w;
        bool ended = false;
        for (uint64_t wi = 0; wi < words; wi++) {  // digit k of a word = residue 7 wi + k; behind the last residue: "end"
            uint32_t x = 0, mul = 1;
            for (uint32_t k = 0; k < kPackPerWord; k++, mul *= kPackBase) {
                const uint64_t r = kPackPerWord * wi + k;
                uint32_t cd = kPackEnd;
                if (r < len && !ended) {
                    cd = code[p[r]];
                    ended = cd == kPackEnd;
                }
                x += cd * mul;
            }
            dst[wi] = x;
        }
        w += words;
    }
    word_offsets[n] = w;
    return 0;
}

extern "C" int ckm_call_batch_packed(ckm_ctx *c, const uint32_t *packed, const uint64_t *word_offsets, uint32_t n, uint32_t flags,
                                     ckm_batch_out_t *out) {
    if (!c || !out || !word_offsets || (n && !packed && word_offsets[n] != word_offsets[0])) return ckm_fail(CKM_EINVAL, "NULL argument");
    memset(out, 0, sizeof *out);
    out->n = n;
    if (flags == CKM_WANT_BEST && n > 1 && (word_offsets[n] - word_offsets[0]) * 4 >= c->pipeline_min_bytes)
        return call_batch_pipelined(c, nullptr, packed, word_offsets, n, out);
    CU(cudaSetDevice(c->device));
    uint32_t max_words = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (word_offsets[i + 1] < word_offsets[i]) return ckm_fail(CKM_EINVAL, "word offsets must be non-decreasing (at %u)", i);
        const uint64_t l = word_offsets[i + 1] - word_offsets[i];
        if (8 * l > 500000000ull) return ckm_fail(CKM_EINVAL, "sequence %u longer than MAX_SEQ_LEN", i);  // kmer_params.h:6
        max_words = std::max<uint32_t>(max_words, (uint32_t)l);
    }
    const uint64_t base0 = word_offsets[0], nw = word_offsets[n] - base0, total = 8 * nw;
    RC(c->in_res.ensure(total + 32));
    RC(c->in_off.ensure(((size_t)n + 1) * 8));
    RC(c->in_packed.ensure((nw + 2) * 4));
    RC(c->in_woff.ensure(((size_t)n + 1) * 8));
    const uint64_t *h_woff = word_offsets;
    if (base0 != 0) {
        RC(c->h_off.ensure(((size_t)n + 1) * 8));
        uint64_t *t = (uint64_t *)c->h_off.p;
        for (uint32_t i = 0; i <= n; i++) t[i] = word_offsets[i] - base0;
        h_woff = t;
    }
    if (nw) CU(cudaMemcpyAsync(c->in_packed.p, packed + base0, nw * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_woff.p, h_woff, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync((uint8_t *)c->in_res.p + total, 0, 32, c->stream));
    {
        const unsigned ub = (unsigned)std::min<uint64_t>(((uint64_t)n + 1 + 7) / 8, (uint64_t)c->sm_count * 32);
        unpack7_kernel<<<ub, kUnpackThreads, 0, c->stream>>>((const uint32_t *)c->in_packed.p, (const uint64_t *)c->in_woff.p, n, 0ull, (uint8_t *)c->in_res.p,
                                                  (uint64_t *)c->in_off.p);
        c->launches++;
    }
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(8u * max_words, 1u), flags));
    return finish_batch(c, n, flags, out);
}

// D2H of what `flags` asks for, after run_device
static int finish_batch(ckm_ctx *c, uint32_t n, uint32_t flags, ckm_batch_out_t *out) {
    // totals decide the size of the compacted outputs
    RC(c->h_totals.ensure(64));
    uint64_t *ht = (uint64_t *)c->h_totals.p;
    CU(cudaMemcpyAsync(ht, c->totals.p, 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (ht[7]) return ckm_fail(CKM_ECUDA, "probe_pc_kernel: a hand-off between probing and scan warps timed out");
    if (ht[6]) return ckm_fail(CKM_EINVAL, "a sequence is longer than the max_len the batch was announced with");
    adapt_probe_path(c, ht[0], ht[4]);
    out->n_probes = ht[0];
    out->n_hits = ht[1];
    const uint64_t n_calls_total = ht[2];
    const bool want_scan = flags & (CKM_WANT_CALLS | CKM_WANT_OTU | CKM_WANT_BEST);
    const unsigned tb = 256;

    if (flags & CKM_WANT_HITS) {
        RC(c->hit_off.ensure(((size_t)n + 1) * 8));
        RC(prefix_sum(c, (const uint32_t *)c->n_hits.p, n, (uint64_t *)c->hit_off.p));
        RC(c->hits_out.ensure((out->n_hits + 1) * sizeof(ckm_hit_t)));
        if (n) {
            export_hits_kernel<<<(unsigned)(((uint64_t)n * 32 + tb - 1) / tb), tb, 0, c->stream>>>(
                (const uint64_t *)c->in_off.p, (const HitRec *)c->hits.p, (const uint64_t *)c->hit_keys.p,
                (const uint16_t *)c->hit_avg.p, (const uint64_t *)c->hit_off.p, n, (ckm_hit_t *)c->hits_out.p);
            c->launches++;
        }
        RC(c->h_hit_off.ensure(((size_t)n + 1) * 8));
        RC(c->h_hits.ensure((out->n_hits + 1) * sizeof(ckm_hit_t)));
        CU(cudaMemcpyAsync(c->h_hit_off.p, c->hit_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        if (out->n_hits)
            CU(cudaMemcpyAsync(c->h_hits.p, c->hits_out.p, out->n_hits * sizeof(ckm_hit_t), cudaMemcpyDeviceToHost, c->stream));
        out->hit_offsets = (const uint64_t *)c->h_hit_off.p;
        out->hits = (const ckm_hit_t *)c->h_hits.p;
    }
    if (flags & CKM_WANT_CALLS) {
        RC(c->call_off.ensure(((size_t)n + 1) * 8));
        RC(prefix_sum(c, (const uint32_t *)c->n_calls.p, n, (uint64_t *)c->call_off.p));
        RC(c->calls_out.ensure((n_calls_total + 1) * sizeof(ckm_call_t)));
        if (n) {
            export_calls_kernel<<<(n + tb - 1) / tb, tb, 0, c->stream>>>((const uint64_t *)c->in_off.p,
                                                                         (const ckm_call_t *)c->calls.p,
                                                                         (const uint64_t *)c->call_off.p, n, c->prm.min_hits,
                                                                         (ckm_call_t *)c->calls_out.p);
            c->launches++;
        }
        RC(c->h_call_off.ensure(((size_t)n + 1) * 8));
        RC(c->h_calls.ensure((n_calls_total + 1) * sizeof(ckm_call_t)));
        CU(cudaMemcpyAsync(c->h_call_off.p, c->call_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        if (n_calls_total)
            CU(cudaMemcpyAsync(c->h_calls.p, c->calls_out.p, n_calls_total * sizeof(ckm_call_t), cudaMemcpyDeviceToHost, c->stream));
        out->call_offsets = (const uint64_t *)c->h_call_off.p;
        out->calls = (const ckm_call_t *)c->h_calls.p;
    }
    if (flags & CKM_WANT_OTU) {
        RC(c->otu_off.ensure(((size_t)n + 1) * 8));
        RC(prefix_sum(c, (const uint32_t *)c->n_otus.p, n, (uint64_t *)c->otu_off.p));
        RC(c->h_otu_off.ensure(((size_t)n + 1) * 8));
        CU(cudaMemcpyAsync(c->h_otu_off.p, c->otu_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint64_t n_otus_total = ((const uint64_t *)c->h_otu_off.p)[n];
        RC(c->otus_out.ensure((n_otus_total + 1) * sizeof(ckm_otu_t)));
        if (n) {
            export_otus_kernel<<<(n + tb - 1) / tb, tb, 0, c->stream>>>((const uint64_t *)c->in_off.p, (const ckm_otu_t *)c->otus.p,
                                                                        (const uint64_t *)c->otu_off.p, n, (ckm_otu_t *)c->otus_out.p);
            c->launches++;
        }
        RC(c->h_otus.ensure((n_otus_total + 1) * sizeof(ckm_otu_t)));
        if (n_otus_total)
            CU(cudaMemcpyAsync(c->h_otus.p, c->otus_out.p, n_otus_total * sizeof(ckm_otu_t), cudaMemcpyDeviceToHost, c->stream));
        out->otu_offsets = (const uint64_t *)c->h_otu_off.p;
        out->otus = (const ckm_otu_t *)c->h_otus.p;
    }
    if (flags & CKM_WANT_BEST) {
        RC(c->h_best.ensure(((size_t)n + 1) * sizeof(ckm_best_t)));
        if (n) CU(cudaMemcpyAsync(c->h_best.p, c->best.p, (size_t)n * sizeof(ckm_best_t), cudaMemcpyDeviceToHost, c->stream));
        out->best = (const ckm_best_t *)c->h_best.p;
    }
    (void)want_scan;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Calibration: what this GPU sustains for INDEPENDENT random sector reads over the resident table --
// the physical ceiling of the hash probe, reported beside the streaming-copy roofline.
// Each thread issues `UNROLL` independent loads per round at hashed slot indices; nothing depends on
// the loaded values except a final XOR that keeps the loads alive.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int UNROLL, int BYTES>
__global__ void __launch_bounds__(256)
gather_calib_kernel(const uint8_t *__restrict__ base, uint64_t n_units /* BYTES-sized units */, uint32_t rounds,
                    uint64_t magic, unsigned long long *__restrict__ sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        uint64_t idx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t h = mix64((tid * rounds + r) * UNROLL + u + 0x9e3779b97f4a7c15ull) >> 28;  // < 2^36
            idx[u] = fast_mod(h, n_units, magic);
        }
        if (BYTES == 16) {
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) v[u] = __ldg(reinterpret_cast<const uint4 *>(base) + idx[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) acc ^= v[u].x ^ v[u].w;
        } else {  // 32 B = one whole sector as two 16 B halves
            uint4 v[UNROLL], w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                v[u] = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * idx[u]);
                w[u] = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * idx[u] + 1);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) acc ^= v[u].x ^ w[u].w;
        }
    }
    if ((uint32_t)acc == 0x12345678u) atomicAdd(sink, 1ull);  // never true in practice; keeps the loads alive
}

// returns accesses/s through *rate; unroll in {1,4,8}, bytes in {16,32}
extern "C" int ckm_calibrate_gather(ckm_ctx *c, int bytes, int unroll, uint32_t rounds, int blocks_per_sm, double *rate,
                                    double *ms_out) {
    if (!c || !c->table.p) return ckm_fail(CKM_ESTATE, "no table loaded");
    CU(cudaSetDevice(c->device));
    RC(c->totals.ensure(64));
    const uint64_t table_bytes = c->num_sigs * (uint64_t)c->slot_bytes;
    const uint64_t n_units = table_bytes / (uint64_t)bytes;
    if (n_units == 0) return ckm_fail(CKM_EINVAL, "table too small");
    const uint64_t magic = (uint64_t)((((unsigned __int128)1) << 64) / n_units);
    const unsigned blocks = (unsigned)(c->sm_count * std::max(1, blocks_per_sm));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    auto launch = [&]() {
#define CKM_CAL(U, B)                                                                                        \
    gather_calib_kernel<U, B><<<blocks, 256, 0, c->stream>>>((const uint8_t *)c->table.p, n_units, rounds, magic, \
                                                             (unsigned long long *)c->totals.p + 4)
        if (bytes == 16 && unroll == 1) CKM_CAL(1, 16);
        else if (bytes == 16 && unroll == 4) CKM_CAL(4, 16);
        else if (bytes == 16 && unroll == 8) CKM_CAL(8, 16);
        else if (bytes == 32 && unroll == 1) CKM_CAL(1, 32);
        else if (bytes == 32 && unroll == 4) CKM_CAL(4, 32);
        else CKM_CAL(8, 32);
#undef CKM_CAL
        c->launches++;
    };
    launch();  // warm-up
    CU(cudaEventRecord(e0, c->stream));
    launch();
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double accesses = (double)blocks * 256.0 * rounds * (double)(unroll == 1 || unroll == 4 || unroll == 8 ? unroll : 8);
    if (rate) *rate = accesses / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}

#include "ckm_family.cuh"
#include "ckm_fq.cuh"
#include "ckm_matrix.cuh"
#include "ckm_family_nr.cuh"
#include "ckm_build.cuh"
