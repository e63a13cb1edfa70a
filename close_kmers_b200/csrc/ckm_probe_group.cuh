// K1 for batches of SHORT sequences (fastq fragments, peptides): the same encode + probe + ordered compaction as
// probe_kernel (ckm_probe.cuh), with a group of G lanes (4, 8 or 16) per sequence instead of the whole warp, so that a warp
// works on 32/G sequences side by side.  A 50-residue fragment has 42 probed windows: one warp step of probe_kernel offers
// 128 window slots for them, a group of 8 lanes offers 32 per step.  Every warp-wide primitive becomes segment-wide
// (shuffles of width G, prefix sums within the group); the groups of a warp iterate in lock step until the longest of
// their sequences is done.  Results are identical to probe_kernel's.
#pragma once
#include "ckm_probe.cuh"

namespace ckm {

template <int G>
__device__ __forceinline__ TileKeys tile_keys_group(const uint8_t *lut, const uint32_t *__restrict__ wb, uint32_t nwords, uint32_t sh,
                                                    uint32_t t0, uint32_t gl, uint32_t len, uint32_t &nwin, bool active) {
    constexpr uint32_t T = 4u * G;  // start positions per group step
    // ---- residues: G words + 3 spill words, re-aligned to the sequence start ----
    const uint32_t wi = (t0 >> 2) + gl;
    uint32_t w = (active && wi < nwords) ? __ldg(wb + wi) : 0u;
    uint32_t x = (active && gl < 3u && wi + G < nwords) ? __ldg(wb + wi + G) : 0u;
    uint32_t w_next = __shfl_down_sync(0xffffffffu, w, 1, G);
    const uint32_t x0 = __shfl_sync(0xffffffffu, x, 0, G);
    if (gl == G - 1u) w_next = x0;
    const uint32_t x_next = __shfl_down_sync(0xffffffffu, x, 1, G);
    const uint32_t a = __funnelshift_r(w, w_next, sh);  // residues t0+4*gl .. +3
    const uint32_t e = __funnelshift_r(x, x_next, sh);  // group lanes 0,1: residues t0+T.. / t0+T+4..

    // the reference scans strlen(seq) residues (kguts.cc:791): an embedded NUL ends the sequence
    {
        const uint32_t za = (a - 0x01010101u) & ~a & 0x80808080u;
        const uint32_t ze = (e - 0x01010101u) & ~e & 0x80808080u;
        uint32_t r = 0xffffffffu;
        if (za) r = t0 + 4u * gl + ((__ffs(za) - 1) >> 3);
        else if (gl < 2u && ze) r = t0 + T + 4u * gl + ((__ffs(ze) - 1) >> 3);
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, d, G));
        if (r < len) nwin = min(nwin, r > CKM_KMER_SIZE ? r - CKM_KMER_SIZE : 0u);
    }

    const uint32_t c0 = codes_of_word(lut, a);
    const uint32_t ce = codes_of_word(lut, e);
    uint32_t c1 = __shfl_down_sync(0xffffffffu, c0, 1, G);
    uint32_t c2 = __shfl_down_sync(0xffffffffu, c0, 2, G);
    const uint32_t e0 = __shfl_sync(0xffffffffu, ce, 0, G);
    const uint32_t e1 = __shfl_sync(0xffffffffu, ce, 1, G);
    if (gl == G - 1u) { c1 = e0; c2 = e1; }
    if (gl == G - 2u) c2 = e0;

    // ---- four keys per lane (as in tile_keys) ----
    uint32_t b[11];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        b[k] = (c0 >> (8 * k)) & 0xFFu;
        b[4 + k] = (c1 >> (8 * k)) & 0xFFu;
        if (k < 3) b[8 + k] = (c2 >> (8 * k)) & 0xFFu;
    }
    const uint32_t i0 = c0 & 0x80808080u, i1 = c1 & 0x80808080u, i2 = c2 & 0x00808080u;
    const uint32_t inv = ((i0 >> 7) & 1u) | ((i0 >> 14) & 2u) | ((i0 >> 21) & 4u) | ((i0 >> 28) & 8u) |
                         ((i1 >> 3) & 16u) | ((i1 >> 10) & 32u) | ((i1 >> 17) & 64u) | ((i1 >> 24) & 128u) |
                         ((i2 << 1) & 256u) | ((i2 >> 6) & 512u) | ((i2 >> 13) & 1024u);
    uint32_t g[8];
    g[0] = ((b[0] * 20u + b[1]) * 20u + b[2]) * 20u + b[3];
#pragma unroll
    for (int q = 0; q < 7; q++) g[q + 1] = (g[q] - b[q] * 8000u) * 20u + b[q + 4];

    const uint32_t q0 = t0 + 4u * gl;
    TileKeys tk;
    tk.act = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool ok = active && (((inv >> j) & 0xFFu) == 0u) && (q0 + j < nwin);
        tk.key[j] = (uint64_t)g[j] * 160000ull + g[j + 4];
        tk.act |= ok ? (1u << j) : 0u;
    }
    return tk;
}

template <bool PACKED, int G>
__global__ void __launch_bounds__(kProbeThreads)
probe_group_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n,
                   HitRec *__restrict__ hits, uint64_t *__restrict__ hit_keys, uint16_t *__restrict__ hit_avg,
                   uint32_t *__restrict__ n_hits, unsigned long long *__restrict__ totals) {
    static_assert(G == 4 || G == 8 || G == 16, "group size");
    constexpr uint32_t kPerWarp = 32u / G, T = 4u * G;
    __shared__ uint8_t lut[256];
    fill_aa_lut(lut);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u, gl = lane % G, grp = lane / G;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t my_probes = 0, my_hits = 0;

    for (uint32_t i_first = warp0 * kPerWarp; i_first < n; i_first += n_warps * kPerWarp) {
        const uint32_t i = i_first + grp;
        const bool valid = i < n;
        const uint64_t base = valid ? __ldg(offsets + i) : 0ull;
        const uint32_t len = valid ? (uint32_t)(__ldg(offsets + i + 1) - base) : 0u;
        // probed starts are p < len-8: the last window is never probed (kguts.cc:792, 798)
        uint32_t nwin = len > CKM_KMER_SIZE ? len - CKM_KMER_SIZE : 0u;
        const uint8_t *p0 = residues + base;
        const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
        const uint32_t *wb = reinterpret_cast<const uint32_t *>(p0 - s);
        const uint32_t nwords = (len + s + 3u) >> 2;
        const uint32_t sh = 8u * s;
        HitRec *out = hits + base;
        uint32_t count = 0;

        for (uint32_t t0 = 0; __any_sync(0xffffffffu, t0 < nwin); t0 += T) {
            const bool active = t0 < nwin;
            const TileKeys tk = tile_keys_group<G>(lut, wb, nwords, sh, t0, gl, len, nwin, active);
            const uint32_t q0 = t0 + 4u * gl;
            const uint32_t act = tk.act;
            uint64_t key[4], h[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                key[j] = tk.key[j];
                h[j] = table_home(tv, key[j]);
            }
            // ---- probe: occupancy bits from L2 first, then the sector loads that are still needed ----
            uint32_t need = act;
            if (tv.occupied) {
                uint32_t bw[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (act & (1u << j)) bw[j] = __ldg(tv.occupied + (h[j] >> 5));
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((act & (1u << j)) && !((bw[j] >> (h[j] & 31u)) & 1u)) need &= ~(1u << j);  // empty slot: miss
            }
            typename SlotIO<PACKED>::raw_t v[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (need & (1u << j)) v[j] = SlotIO<PACKED>::load(tv.slots, h[j]);

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

            // ---- ordered compaction within the group: exclusive prefix of per-lane hit counts ----
            const uint32_t cnt = __popc(hm);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d, G);
                if (gl >= (uint32_t)d) incl += t;
            }
            const uint32_t tile_hits = __shfl_sync(0xffffffffu, incl, G - 1, G);
            uint32_t o = count + incl - cnt;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (hm & (1u << j)) {
                    HitRec rec;
                    rec.pos = q0 + j;
                    rec.fI = f[j].fI;
                    rec.wt = f[j].wt;
                    rec.oI = f[j].oI;
                    out[o] = rec;
                    if (hit_keys) hit_keys[base + o] = key[j];
                    if (hit_avg) hit_avg[base + o] = (uint16_t)f[j].avg;
                    o++;
                }
            }
            count += tile_hits;
        }
        if (valid && gl == 0) {
            n_hits[i] = count;
            my_hits += count;
        }
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
