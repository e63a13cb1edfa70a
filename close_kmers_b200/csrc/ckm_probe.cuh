// K1: warp-cooperative 8-mer encode + signature-table probe + ballot/prefix compaction of hits.
//
// Replaces, for a whole batch, the per-residue loop of KmerGuts::gather_hits up to and including the
// hit callback (kguts.cc:783-815, 857-871) together with to_amino_acid_off (273-339), encoded_kmer
// (438-455), advance_past_ambig (682-732) and lookup_hash_entry (585-602).
//
// One warp per protein.  A warp step covers 128 consecutive start positions: every lane loads one
// aligned 32-bit word of residues (128 B coalesced per warp), converts it to 5-bit codes through a
// shared-memory table, borrows the next two words from its neighbours by shuffle, and owns the four
// windows that start in its word.  The four keys are built from two 4-residue halves
// (key = hi * 20^4 + lo), reduced modulo the bucket count by a multiply-high, and the four table
// slots are fetched as four independent 16-byte loads (one DRAM sector each) before any is examined,
// so each lane keeps four random HBM accesses in flight.  Hits are compacted in position order with a
// warp prefix sum over per-lane hit counts and appended to the protein's hit region
// hits[offsets[i] ...) -- a protein of length L has at most L-8 hits, so regions indexed by residue
// offset never overlap and no allocation pass is needed.
#pragma once
#include "ckm_common.cuh"

namespace ckm {

struct SlotFields {
    uint32_t fI;
    float wt;
    int32_t oI;
    uint32_t avg;
};

// returns 1 = hit (fields filled), 0 = keep probing, -1 = empty slot (miss)
template <bool PACKED>
struct SlotIO;

template <>
struct SlotIO<true> {
    typedef uint4 raw_t;
    static __device__ __forceinline__ raw_t load(const void *slots, uint64_t h) {
        return __ldg(reinterpret_cast<const uint4 *>(slots) + h);
    }
    static __device__ __forceinline__ raw_t load_hint(const void *slots, uint64_t h, uint64_t policy) {
        return policy ? ldg_v4_hint(reinterpret_cast<const uint4 *>(slots) + h, policy)
                      : ldg_v4_l2_64(reinterpret_cast<const uint4 *>(slots) + h);
    }
    static __device__ __forceinline__ int test(const raw_t &v, uint64_t key, SlotFields &f) {
        if (v.x == (uint32_t)key && (v.y & 0xFu) == (uint32_t)(key >> 32)) {
            f.fI = v.w & (kPackedFieldLimit - 1);
            f.wt = __uint_as_float(v.z);
            f.avg = (v.y >> 4) & 0xFFFFu;
            f.oI = (int32_t)(((v.y >> 20) & 0xFFFu) | ((v.w >> 22) << 12)) - 1;
            return 1;
        }
        return (v.y & 0x8u) ? -1 : 0;
    }
};

struct Raw3 {
    uint64_t k, a, b;
};
template <>
struct SlotIO<false> {
    typedef Raw3 raw_t;
    static __device__ __forceinline__ raw_t load(const void *slots, uint64_t h) {
        const uint64_t *s = reinterpret_cast<const uint64_t *>(slots) + 3 * h;
        Raw3 r;
        r.k = __ldg(s);
        r.a = __ldg(s + 1);
        r.b = __ldg(s + 2);
        return r;
    }
    static __device__ __forceinline__ raw_t load_hint(const void *slots, uint64_t h, uint64_t) { return load(slots, h); }
    static __device__ __forceinline__ int test(const raw_t &v, uint64_t key, SlotFields &f) {
        if (v.k == key) {
            f.oI = (int32_t)(uint32_t)v.a;
            f.avg = (uint32_t)(v.a >> 32) & 0xFFFFu;
            f.fI = (uint32_t)v.b;
            f.wt = __uint_as_float((uint32_t)(v.b >> 32));
            return 1;
        }
        return v.k > CKM_MAX_ENCODED ? -1 : 0;
    }
};


// The four 8-mer keys a lane owns in one 128-position warp step (positions t0 + 4*lane + j), with bit j of `act` set
// when window j is probed: all eight residues valid (kguts.cc:273-339, 682-732) and j-th start < nwin.  `nwin` (= len-8:
// the last window is never probed, kguts.cc:792) shrinks when an embedded NUL ends the protein (strlen, kguts.cc:791).
struct TileKeys {
    uint64_t key[4];
    uint32_t act;
    uint32_t c0, c1, c2;  // the lane's residue codes t0 + 4*lane .. +10 (one per byte, invalid = 0x80; the top byte of c2 is unused)
};

// the residue words of one step: 32 words (one per lane) + 3 spill words (lanes 0-2)
__device__ __forceinline__ void tile_words(const uint32_t *__restrict__ wb, uint32_t nwords, uint32_t t0, uint32_t lane, uint32_t &w,
                                           uint32_t &x) {
    const uint32_t wi = (t0 >> 2) + lane;
    w = wi < nwords ? __ldg(wb + wi) : 0u;
    x = (lane < 3u && wi + 32u < nwords) ? __ldg(wb + wi + 32u) : 0u;
}

__device__ __forceinline__ TileKeys tile_keys_from(const uint8_t *lut, uint32_t w, uint32_t x, uint32_t sh, uint32_t t0, uint32_t lane,
                                                   uint32_t len, uint32_t &nwin);

__device__ __forceinline__ TileKeys tile_keys(const uint8_t *lut, const uint32_t *__restrict__ wb, uint32_t nwords, uint32_t sh,
                                              uint32_t t0, uint32_t lane, uint32_t len, uint32_t &nwin) {
    uint32_t w, x;
    tile_words(wb, nwords, t0, lane, w, x);
    return tile_keys_from(lut, w, x, sh, t0, lane, len, nwin);
}

__device__ __forceinline__ TileKeys tile_keys_from(const uint8_t *lut, uint32_t w, uint32_t x, uint32_t sh, uint32_t t0, uint32_t lane,
                                                   uint32_t len, uint32_t &nwin) {
    // ---- residues re-aligned to the protein start ----
    uint32_t w_next = __shfl_down_sync(0xffffffffu, w, 1);
    const uint32_t x0 = __shfl_sync(0xffffffffu, x, 0);
    if (lane == 31u) w_next = x0;
    const uint32_t x_next = __shfl_down_sync(0xffffffffu, x, 1);
    const uint32_t a = __funnelshift_r(w, w_next, sh);   // residues t0+4*lane .. +3
    const uint32_t e = __funnelshift_r(x, x_next, sh);   // lanes 0,1: residues t0+128.. / t0+132..

    // the reference scans strlen(seq) residues (kguts.cc:791): an embedded NUL ends the protein
    {
        const uint32_t za = (a - 0x01010101u) & ~a & 0x80808080u;
        const uint32_t ze = (e - 0x01010101u) & ~e & 0x80808080u;
        uint32_t r = 0xffffffffu;
        if (za) r = t0 + 4u * lane + ((__ffs(za) - 1) >> 3);
        else if (lane < 2u && ze) r = t0 + kTile + 4u * lane + ((__ffs(ze) - 1) >> 3);
        r = __reduce_min_sync(0xffffffffu, r);
        if (r < len) nwin = min(nwin, r > CKM_KMER_SIZE ? r - CKM_KMER_SIZE : 0u);
    }

    const uint32_t c0 = codes_of_word(lut, a);
    const uint32_t ce = codes_of_word(lut, e);
    uint32_t c1 = __shfl_down_sync(0xffffffffu, c0, 1);
    uint32_t c2 = __shfl_down_sync(0xffffffffu, c0, 2);
    const uint32_t e0 = __shfl_sync(0xffffffffu, ce, 0);
    const uint32_t e1 = __shfl_sync(0xffffffffu, ce, 1);
    if (lane == 31u) { c1 = e0; c2 = e1; }
    if (lane == 30u) c2 = e0;

    // ---- four keys per lane ----
    uint32_t b[11];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        b[k] = (c0 >> (8 * k)) & 0xFFu;
        b[4 + k] = (c1 >> (8 * k)) & 0xFFu;
        if (k < 3) b[8 + k] = (c2 >> (8 * k)) & 0xFFu;
    }
    // bit i of inv set <=> residue i of the lane's 11 is not one of the 20 amino acids
    const uint32_t i0 = c0 & 0x80808080u, i1 = c1 & 0x80808080u, i2 = c2 & 0x00808080u;
    const uint32_t inv = ((i0 >> 7) & 1u) | ((i0 >> 14) & 2u) | ((i0 >> 21) & 4u) | ((i0 >> 28) & 8u) |
                         ((i1 >> 3) & 16u) | ((i1 >> 10) & 32u) | ((i1 >> 17) & 64u) | ((i1 >> 24) & 128u) |
                         ((i2 << 1) & 256u) | ((i2 >> 6) & 512u) | ((i2 >> 13) & 1024u);
    uint32_t g[8];
    g[0] = ((b[0] * 20u + b[1]) * 20u + b[2]) * 20u + b[3];
#pragma unroll
    for (int q = 0; q < 7; q++) g[q + 1] = (g[q] - b[q] * 8000u) * 20u + b[q + 4];

    const uint32_t q0 = t0 + 4u * lane;
    TileKeys tk;
    tk.act = 0;
    tk.c0 = c0;
    tk.c1 = c1;
    tk.c2 = c2;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool ok = (((inv >> j) & 0xFFu) == 0u) && (q0 + j < nwin);
        tk.key[j] = (uint64_t)g[j] * 160000ull + g[j + 4];
        tk.act |= ok ? (1u << j) : 0u;
    }
    return tk;
}

constexpr int kProbeThreads = 256;

template <bool PACKED>
__global__ void __launch_bounds__(kProbeThreads)
probe_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n,
             HitRec *__restrict__ hits, uint64_t *__restrict__ hit_keys, uint16_t *__restrict__ hit_avg,
             uint32_t *__restrict__ n_hits, unsigned long long *__restrict__ totals) {
    __shared__ uint8_t lut[256];
    fill_aa_lut(lut);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t my_probes = 0, my_hits = 0;
#ifdef CKM_EXPERIMENTS  // cache-policy A/B (profiles/r1/tune1_cache_policy.jsonl: no effect or worse)
    const bool t_tab = tv.tuning & 17u, t_bm = tv.tuning & 2u, t_st = tv.tuning & 4u;
    const uint64_t pol_first = (tv.tuning & 16u) ? 0ull : policy_evict_first(), pol_last = policy_evict_last();
#else
    constexpr bool t_tab = false, t_bm = false, t_st = false;
    constexpr uint64_t pol_first = 0, pol_last = 0;
#endif

    for (uint32_t i = warp0; i < n; i += n_warps) {
        const uint64_t base = __ldg(offsets + i);
        const uint32_t len = (uint32_t)(__ldg(offsets + i + 1) - base);
        uint32_t count = 0;
        if (len > CKM_KMER_SIZE) {
            // probed starts are p < len-8: the last window is never probed (kguts.cc:792, 798)
            uint32_t nwin = len - CKM_KMER_SIZE;
            const uint8_t *p0 = residues + base;
            const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
            const uint32_t *wb = reinterpret_cast<const uint32_t *>(p0 - s);
            const uint32_t nwords = (len + s + 3u) >> 2;
            const uint32_t sh = 8u * s;
            HitRec *out = hits + base;

            for (uint32_t t0 = 0; t0 < nwin; t0 += kTile) {
                const TileKeys tk = tile_keys(lut, wb, nwords, sh, t0, lane, len, nwin);
                const uint32_t q0 = t0 + 4u * lane;
                const uint32_t act = tk.act;
                uint64_t key[4], h[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    key[j] = tk.key[j];
                    h[j] = table_home_wide(tv, key[j]);
                }

                // ---- probe: occupancy bits from L2 first, then the sector loads that are still needed ----
                uint32_t need = act;
                if (tv.occupied) {
                    uint32_t bw[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (act & (1u << j)) bw[j] = t_bm ? ldg_u32_hint(tv.occupied + (h[j] >> 5), pol_last) : __ldg(tv.occupied + (h[j] >> 5));
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((act & (1u << j)) && !((bw[j] >> (h[j] & 31u)) & 1u)) need &= ~(1u << j);  // empty slot: miss
                }
                typename SlotIO<PACKED>::raw_t v[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (need & (1u << j)) v[j] = t_tab ? SlotIO<PACKED>::load_hint(tv.slots, h[j], pol_first) : SlotIO<PACKED>::load(tv.slots, h[j]);

                SlotFields f[4];
                uint32_t hm = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (need & (1u << j)) {
                        int r = SlotIO<PACKED>::test(v[j], key[j], f[j]);
                        uint64_t guard = 0;
                        while (r == 0) {  // linear probing: h = (h+1) % size_hash (kguts.cc:589)
                            h[j] = (h[j] + 1 == tv.num_sigs) ? 0 : h[j] + 1;
                            if (++guard >= tv.num_sigs) { r = -1; break; }  // table without an empty slot
                            if (tv.occupied && !((__ldg(tv.occupied + (h[j] >> 5)) >> (h[j] & 31u)) & 1u)) { r = -1; break; }
                            v[j] = SlotIO<PACKED>::load(tv.slots, h[j]);
                            r = SlotIO<PACKED>::test(v[j], key[j], f[j]);
                        }
                        if (r > 0) hm |= 1u << j;
                    }
                }
                my_probes += __popc(act);

                // ---- ordered compaction: exclusive prefix of per-lane hit counts ----
                const uint32_t cnt = __popc(hm);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += t;
                }
                const uint32_t tile_hits = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t o = count + incl - cnt;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (hm & (1u << j)) {
                        HitRec rec;
                        rec.pos = q0 + j;
                        rec.fI = f[j].fI;
                        rec.wt = f[j].wt;
                        rec.oI = f[j].oI;
                        if (t_st) stg_v4_hint(reinterpret_cast<uint4 *>(out + o), *reinterpret_cast<const uint4 *>(&rec), pol_first);
                        else out[o] = rec;
                        if (hit_keys) hit_keys[base + o] = key[j];
                        if (hit_avg) hit_avg[base + o] = (uint16_t)f[j].avg;
                        o++;
                    }
                }
                count += tile_hits;
            }
        }
        if (lane == 0) n_hits[i] = count;
        if (lane == 0) my_hits += count;
    }

    // batch totals: one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_probes += __shfl_down_sync(0xffffffffu, my_probes, d);
        my_hits += __shfl_down_sync(0xffffffffu, my_hits, d);
    }
    if (lane == 0) {
        atomicAdd(totals + 0, (unsigned long long)my_probes);
        atomicAdd(totals + 1, (unsigned long long)my_hits);
    }
}

}  // namespace ckm
