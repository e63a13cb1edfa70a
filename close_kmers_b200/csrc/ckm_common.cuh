// Shared device-side definitions for the signature-k-mer calling path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ckm.h"

namespace ckm {

// ---------------------------------------------------------------------------------------------------
// Table layouts in HBM.  The reference file stores 24-byte slots behind a 24-byte header
// (kmer_image.h:11-23), so half of them straddle a 32-byte DRAM sector.  At load time the slots are
// repacked, in the SAME order and bucket count (so probe sequences and therefore results are
// unchanged), into 16-byte slots -- two per sector, never straddling:
//
//   w0 (u64): [0,35)  which_kmer      (valid keys are < 20^8 < 2^35)
//             [35]    empty flag      (reference: which_kmer > MAX_ENCODED, kguts.cc:587,596)
//             [36,52) avg_from_end
//             [52,64) (otu_index+1) low 12 bits
//   w1 (u64): [0,32)  function_wt (f32 bits)
//             [32,54) function_index  (22 bits; function.index holds < 1e6 entries, kguts.cc:541)
//             [54,64) (otu_index+1) high 10 bits
//
// If any occupied slot does not fit (function_index outside [0,2^22) or otu_index outside
// [-1,2^22-1)) the loader keeps the verbatim 24-byte slots and the kernels run their RAW
// instantiation; results are identical either way.
// ---------------------------------------------------------------------------------------------------
constexpr int kPackedSlotBytes = 16;
constexpr int kRawSlotBytes = 24;
constexpr uint32_t kPackedFieldLimit = 1u << 22;

struct TableView {
    const void *slots;   // packed: uint4[num_sigs]; raw: 24-byte ckm_sig_kmer_t[num_sigs] (8-byte aligned)
    uint64_t num_sigs;
    uint64_t magic;      // floor(2^64 / num_sigs), for key % num_sigs without a divide
    // Occupancy bitmap (bit h = slot h holds a k-mer), 1/128 of the packed table: small enough to live in L2
    // (63.5 MB for 508M buckets) while the table itself does not.  A probe whose slot is empty -- most misses at
    // the reference's <= 1/3 load factor -- is answered from L2 and never becomes a DRAM transaction, which is
    // what bounds this kernel.  NULL when the table itself fits L2.
    const uint32_t *occupied;
    // ckm_set_tuning bits (include/ckm.h); the kernels themselves read them only in CKM_EXPERIMENTS builds
    uint32_t tuning;
    // Neighbour-ordered copy of the occupied slots and each slot's index in it (ckm_chain.cuh); NULL when not built.
    const uint4 *chain;
    // per table slot: (low 32 bits of its k-mer, its index in the copy) -- what hint_kernel needs of a slot, in ONE 8-byte read
    // instead of the slot and the index from two arrays; .y == 0xFFFFFFFF: an empty slot.  Hints steer, keys decide: a k-mer
    // that differs from the sampled one in its top three bits only yields a useless hint, never a wrong answer.
    const uint2 *cpos;
    // the compact form of the copy (chain_compact_kernel): residue string and (weight, function word) per index
    const uint8_t *cres;
    const uint2 *cpay;
    uint32_t n_chain;
    uint32_t m35;  // floor(2^35 / num_sigs) when 64 <= num_sigs < 2^32 (fast_mod35), else 0
    // != 0: the table is the library's own -- 2^hbits 16-byte slots, home bucket = top hbits bits of a multiplicative hash of the
    // key (own_home), every key of the image re-inserted by linear probing (rehash_kernel) -- instead of the image's slots in the
    // image's order under key % num_sigs.  A lookup is a function of the key alone, so it returns the same fields either way;
    // the hash costs 3 instructions where the modulo by the image's (odd, non-power-of-two) bucket count costs 13, four times
    // per lane and step of K1.
    uint32_t hbits;
};
__host__ __device__ __forceinline__ uint32_t own_home(uint64_t key, uint32_t hbits) {
    return ((uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u) >> (32u - hbits);
}

// One table hit as the ordered scoring scan consumes it (KmerHit, kguts.h:154-163, minus the key).
struct __align__(16) HitRec {
    uint32_t pos;  // from0_in_prot
    uint32_t fI;
    float wt;
    int32_t oI;
};

constexpr int kTile = 128;  // start positions per warp step of K1 (4 per lane)

struct Params {  // kguts.h:290-293
    int order_constraint, min_hits, min_weighted_hits, max_gap;
};

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg_v4_hint(const uint4 *p, uint64_t policy) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint4 ldg_v4_l2_64(const uint4 *p) {  // L2 fetches 64 B instead of the whole 128-byte line
    uint4 v;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const uint32_t *p, uint64_t policy) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void stg_v4_hint(uint4 *p, const uint4 &v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}

// asynchronous global -> shared copies (LDGSTS): no registers are held while the data is in flight
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc, uint64_t policy) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// key % d with d = num_sigs: q = mulhi(key, floor(2^64/d)) is floor(key/d) or one less.
__device__ __forceinline__ uint64_t fast_mod(uint64_t key, uint64_t d, uint64_t magic) {
    uint64_t q = __umul64hi(key, magic);
    uint64_t r = key - q * d;
    return r >= d ? r - d : r;
}

// The same for keys below 2^35 (every valid 8-mer: 20^8 < 2^35) and 64 <= d < 2^32, in 32-bit pieces: with m = floor(2^35/d),
// q = (key * m) >> 35 is floor(key/d) or one less (key * m < 2^64 because m <= 2^29), so key - q*d < 2d.  A third of the
// instructions of fast_mod, which is a fifth of K1's instruction stream.
__device__ __forceinline__ uint32_t fast_mod35(uint64_t key, uint32_t d, uint32_t m35) {
    const uint64_t prod = (uint64_t)(uint32_t)key * m35 + ((uint64_t)((uint32_t)(key >> 32) * m35) << 32);
    const uint32_t q = (uint32_t)(prod >> 35);
    const uint64_t r = key - (uint64_t)q * d;
    return r >= d ? (uint32_t)(r - d) : (uint32_t)r;
}
// home bucket of a valid 8-mer key
__device__ __forceinline__ uint64_t table_home(const TableView &tv, uint64_t key) {
    if (tv.hbits) return own_home(key, tv.hbits);
    return tv.m35 ? (uint64_t)fast_mod35(key, (uint32_t)tv.num_sigs, tv.m35) : fast_mod(key, tv.num_sigs, tv.magic);
}
// the same without fast_mod35 (probe_kernel: the 32-bit path costs it sixteen registers, profiles/r1/tune_hint_v11_1M.jsonl)
__device__ __forceinline__ uint64_t table_home_wide(const TableView &tv, uint64_t key) {
    return tv.hbits ? (uint64_t)own_home(key, tv.hbits) : fast_mod(key, tv.num_sigs, tv.magic);
}
static inline uint32_t magic35(uint64_t num_sigs) {
    return (num_sigs >= 64 && num_sigs < 0xFFFFFFFFull) ? (uint32_t)((1ull << 35) / num_sigs) : 0u;
}

// ASCII -> 0..19 for ACDEFGHIKLMNPQRSTVWY, everything else invalid (kguts.cc:273-339).
// Invalid is encoded as 0x80 so that "any invalid residue in a word" is one AND with 0x80808080.
constexpr uint8_t kInvalidCode = 0x80;

__device__ __forceinline__ void fill_aa_lut(uint8_t *lut /*256 B shared*/) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        uint8_t v = kInvalidCode;
        switch (c) {
            case 'A': v = 0; break;  case 'C': v = 1; break;  case 'D': v = 2; break;  case 'E': v = 3; break;
            case 'F': v = 4; break;  case 'G': v = 5; break;  case 'H': v = 6; break;  case 'I': v = 7; break;
            case 'K': v = 8; break;  case 'L': v = 9; break;  case 'M': v = 10; break; case 'N': v = 11; break;
            case 'P': v = 12; break; case 'Q': v = 13; break; case 'R': v = 14; break; case 'S': v = 15; break;
            case 'T': v = 16; break; case 'V': v = 17; break; case 'W': v = 18; break; case 'Y': v = 19; break;
            default: break;
        }
        lut[c] = v;
    }
}

// four ASCII bytes -> four codes
__device__ __forceinline__ uint32_t codes_of_word(const uint8_t *lut, uint32_t w) {
    uint32_t c0 = lut[w & 0xFF], c1 = lut[(w >> 8) & 0xFF], c2 = lut[(w >> 16) & 0xFF], c3 = lut[w >> 24];
    return c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
}

}  // namespace ckm
