// K1 through the neighbour copy with precomputed hints: two kernels, no dependency between the 128-window steps.
//
// probe_chain_kernel (ckm_chain.cuh) cuts the DRAM transactions per probe by a factor of three and is still no faster
// than plain hash probing, because it learns where a protein sits in chain[] *while* walking it: every warp step waits
// for the previous step's offset and then runs up to three anchor passes, each a chain of dependent DRAM round trips
// (slot -> cpos -> chain entry).  A dozen serial round trips per protein with ~20 warps per SM is latency, not bandwidth.
//
// Here the dependency is taken out of the walk:
//
//   hint_kernel        one thread per *sample* window (one window in 64): plain lookup_hash_entry (kguts.cc:585-602); a hit
//                      stores hint = cpos[slot] - position, "where window 0 of this protein would sit in chain[] if the
//                      protein followed this chain".  All samples of a batch are independent: two or three round trips at
//                      full occupancy, 1.6 % of the probes.
//   probe_hint_kernel  probe_kernel (ckm_probe.cuh) with one extra source: every window first compares its key with
//                      chain[hint + position], the hint being that of its own 64-window segment or else of the nearest
//                      segment that has one.  Those reads are issued together with the occupancy words, are coalesced (the
//                      128 windows of a step read 2 KB = 16 lines of chain[] when they share a hint) and answer ~98 % of the
//                      hits of a protein that is a homologue of a signature source.  What is left -- windows whose home slot
//                      is occupied and whose chain entry did not match: misses that collide with another k-mer, hits off
//                      the chain, windows without a hint -- is hash-probed exactly as probe_kernel does it.
//
// A hint only chooses where to look first; the key is compared in full and every unresolved window takes the reference's
// probe sequence, so results are bit-identical to probe_kernel whatever the hints hold (tests/test_gpu_chain.py runs worlds
// with duplicate k-mers, cycles, isolated k-mers, indels and 39 000-residue proteins side by side with plain probing and the
// oracle, through every build of the kernel).
//
// Measured: 4.2-4.4 ms per 1M-protein C2 step against 7.1 ms for probe_kernel, 16.3 GB of DRAM
// traffic against 30.9 GB.  The kernel is latency-bound (three dependent memory phases per step: residues, occupancy words +
// chain entries, left-over slots), so its shape is chosen for resident warps without spills: 128-thread blocks, 72 registers,
// the slot behind each hit parked in shared memory (STAGE) -- 28 warps per SM.
#pragma once
#include "ckm_chain.cuh"

namespace ckm {

#ifndef CKM_HINT_SHIFT
#define CKM_HINT_SHIFT 6  // measured on C2: one sample per 32 / 64 / 128 windows -> K1 4.23 / 4.12 / 4.24 ms (tune_hint_v18_c2_*)
#endif
constexpr uint32_t kHintShift = CKM_HINT_SHIFT;    // one sample per 64 windows
constexpr int kSegsPerTile = 128 >> kHintShift;    // hint segments per 128-window step
constexpr uint32_t kHintSeg = 1u << kHintShift;
constexpr uint32_t kNoHint = 0xFFFFFFFFu;
constexpr uint32_t kHintLanes = 8;                 // lanes per protein in hint_kernel (a 300-residue protein has 5 samples)

// hints of protein i (global index gi) live at hints[(offsets[i] >> kHintShift) + gi ...): a protein of length L has at most
// (L-1)/64 + 1 segments and consecutive regions start at least that far apart, so regions never overlap.
__device__ __forceinline__ uint64_t hint_region(uint64_t seq_base, uint32_t gi) { return (seq_base >> kHintShift) + gi; }

__global__ void __launch_bounds__(256)
hint_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n, uint32_t index_base,
            uint32_t *__restrict__ hints) {
    __shared__ uint8_t lut[256];
    fill_aa_lut(lut);
    __syncthreads();
    const uint32_t sub = threadIdx.x & (kHintLanes - 1);
    const uint32_t g0 = (blockIdx.x * blockDim.x + threadIdx.x) / kHintLanes;
    const uint32_t n_groups = (gridDim.x * blockDim.x) / kHintLanes;
    const uint32_t nsig = (uint32_t)tv.num_sigs;
    for (uint32_t i = g0; i < n; i += n_groups) {
        const uint64_t seq_base = __ldg(offsets + i);
        const uint32_t len = (uint32_t)(__ldg(offsets + i + 1) - seq_base);
        if (len <= CKM_KMER_SIZE) continue;
        const uint32_t nwin = len - CKM_KMER_SIZE;
        const uint32_t nseg = (nwin + kHintSeg - 1) >> kHintShift;
        uint32_t *out = hints + hint_region(seq_base, index_base + i);
        for (uint32_t k = sub; k < nseg; k += kHintLanes) {
            uint32_t p = (k << kHintShift) + kHintSeg / 2;  // middle of the segment, or its first window when it is short
            if (p >= nwin) p = k << kHintShift;
            const uint8_t *r = residues + seq_base + p;
            uint64_t key = 0;
            uint32_t bad = 0;
#pragma unroll
            for (int q = 0; q < CKM_KMER_SIZE; q++) {
                const uint32_t code = lut[__ldg(r + q)];
                bad |= code;
                key = key * 20ull + (code & 0x1Fu);
            }
            uint32_t hint = kNoHint;
            if (!(bad & kInvalidCode)) {
                uint32_t h = (uint32_t)table_home(tv, key);
                for (uint32_t steps = 0; steps < nsig; steps++) {
                    if (tv.occupied && !((__ldg(tv.occupied + (h >> 5)) >> (h & 31u)) & 1u)) break;
                    const uint2 t = __ldg(tv.cpos + h);  // (k-mer low word, index in the copy): one DRAM access per slot looked at
                    if (t.y == kNoSlot) break;           // an empty slot (or one no lookup ends at)
                    if (t.x == (uint32_t)key) {
                        hint = t.y - p;
                        break;
                    }
                    h = (h + 1u == nsig) ? 0u : h + 1u;
                }
            }
            out[k] = hint;
        }
    }
}

// stage[] index of a step's window e (STAGE variant): 16-byte entries, four consecutive ones per lane; XOR-ing the low three
// bits with the next three spreads a warp's accesses over all banks
__device__ __forceinline__ uint32_t stage_at(uint32_t e) { return e ^ ((e >> 3) & 7u); }

// STAGE: the slot behind every hit of a step waits in shared memory (one entry per window) instead of in twelve registers per
// lane, which is what keeps the register variant at 80 registers / 24 warps per SM.
template <int THREADS, int MINB, bool STAGE>
__global__ void __launch_bounds__(THREADS, MINB)
probe_hint_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n, uint32_t index_base,
                  const uint32_t *__restrict__ hints, HitRec *__restrict__ hits, uint64_t *__restrict__ hit_keys,
                  uint16_t *__restrict__ hit_avg, uint32_t *__restrict__ n_hits, unsigned long long *__restrict__ totals) {
    __shared__ uint8_t lut[256];
    __shared__ uint4 queues[THREADS / 32][kTile];  // per warp: windows left for the hash probe, then their results
    __shared__ uint4 stages[STAGE ? THREADS / 32 : 1][kTile];
    fill_aa_lut(lut);
    __syncthreads();
    uint4 *queue = queues[threadIdx.x >> 5];
    uint4 *stage = stages[STAGE ? threadIdx.x >> 5 : 0];

    constexpr uint32_t full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint4 *__restrict__ slots = reinterpret_cast<const uint4 *>(tv.slots);
    const uint32_t nsig = (uint32_t)tv.num_sigs;  // the neighbour copy is only built for tables below 2^32 buckets
    uint32_t my_probes = 0, my_hits = 0, my_chain = 0;
    // L2 policies: the read-once streams (chain entries, slots) are loaded evict_first, which keeps more of the occupancy bitmap
    // in L2 (4.41 -> 4.22 ms on C2, profiles/r1/tune_hint_v15_c2.jsonl).  CKM_EXPERIMENTS builds can switch that and the L2
    // prefetches off, load the occupancy words evict_last or store hit records evict_first (measured: no effect).
#ifdef CKM_EXPERIMENTS
    const bool pf = !(tv.tuning & 0x80000u);
    const bool t_tab = !(tv.tuning & 1u), t_bm = tv.tuning & 2u, t_st = tv.tuning & 4u;
    const uint64_t pol_last = policy_evict_last();
#else
    constexpr bool pf = true, t_tab = true, t_bm = false, t_st = false;
    constexpr uint64_t pol_last = 0;
#endif
    const uint64_t pol_first = policy_evict_first();

    for (uint32_t i = warp0; i < n; i += n_warps) {
        const uint64_t seq_base = __ldg(offsets + i);
        const uint32_t len = (uint32_t)(__ldg(offsets + i + 1) - seq_base);
        uint32_t count = 0;
        if (len > CKM_KMER_SIZE) {
            uint32_t nwin = len - CKM_KMER_SIZE;  // the last window is never probed (kguts.cc:792, 798)
            const uint32_t nseg = (nwin + kHintSeg - 1) >> kHintShift;
            const uint32_t *hp = hints + hint_region(seq_base, index_base + i);
            const uint8_t *p0 = residues + seq_base;
            const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
            const uint32_t *wb = reinterpret_cast<const uint32_t *>(p0 - s);
            const uint32_t nwords = (len + s + 3u) >> 2;
            const uint32_t sh = 8u * s;
            HitRec *out = hits + seq_base;
            uint32_t hv = kNoHint;     // hint of segment (32 * round + lane), reloaded every 32 segments
            uint32_t carry = kNoHint;  // last hint seen in front of the current segment (the first one there is, to begin with)
            uint32_t rw, rx;           // residue words of the step, fetched one step ahead
            tile_words(wb, nwords, 0, lane, rw, rx);

            for (uint32_t t0 = 0; t0 < nwin; t0 += kTile) {
                const uint32_t seg0 = t0 >> kHintShift;  // first of the kSegsPerTile segments of this step
                if ((seg0 & 31u) == 0u) {
                    hv = (seg0 + lane < nseg) ? __ldg(hp + seg0 + lane) : kNoHint;
                    if (carry == kNoHint) {
                        const uint32_t m = __ballot_sync(full, hv != kNoHint);
                        if (m) carry = __shfl_sync(full, hv, __ffs(m) - 1);
                    }
                }
                const uint32_t q0 = t0 + 4u * lane;

                // ---- this lane's hint: its segment's, else the nearest one in front, else the first one of the protein ----
                uint32_t mh = kNoHint;
#pragma unroll
                for (int k = 0; k < kSegsPerTile; k++) {
                    uint32_t fk = __shfl_sync(full, hv, (seg0 & 31u) + k);
                    if (fk == kNoHint) fk = carry;
                    else carry = fk;
                    if ((int)(lane * kSegsPerTile >> 5) == k) mh = fk;
                }
                // the chain entries the hint predicts (coalesced) are requested before the keys are even built: their
                // addresses depend on the position only
                uint4 cv[4];
                uint32_t ok = 0;
                if (mh != kNoHint) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t idx = mh + q0 + j;
                        if (q0 + j < nwin && idx < tv.n_chain) {
                            cv[j] = t_tab ? ldg_v4_hint(tv.chain + idx, pol_first) : __ldg(tv.chain + idx);
                            ok |= 1u << j;
                        }
                    }
                    if (pf && t0 + kTile < nwin && mh + q0 + kTile < tv.n_chain) prefetch_l2(tv.chain + (mh + q0 + kTile));
                }

                const TileKeys tk = tile_keys_from(lut, rw, rx, sh, t0, lane, len, nwin);
                if (t0 + kTile < nwin) tile_words(wb, nwords, t0 + kTile, lane, rw, rx);
                const uint32_t act = tk.act;
                if (pf) {
                    // pull into L2 what the next step (or, from the first step, the next protein of this warp) starts with
                    if (lane < 2u && t0 + 2u * kTile + 128u * lane < len) prefetch_l2(p0 + t0 + 2u * kTile + 128u * lane);
                    if (t0 == 0 && i + n_warps < n && lane >= 2u && lane < 5u) {
                        const uint64_t nb = __ldg(offsets + i + n_warps);
                        if (lane < 4u) prefetch_l2(residues + nb + 128u * (lane - 2u));
                        else prefetch_l2(hints + hint_region(nb, index_base + i + n_warps));
                    }
                }
                uint32_t h[4];
#pragma unroll
                for (int j = 0; j < 4; j++) h[j] = (uint32_t)table_home(tv, tk.key[j]);

                // ---- occupancy words (L2) ----
                uint32_t bw[4] = {0u, 0u, 0u, 0u};
                if (tv.occupied) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (act & (1u << j)) bw[j] = t_bm ? ldg_u32_hint(tv.occupied + (h[j] >> 5), pol_last) : __ldg(tv.occupied + (h[j] >> 5));
                }
                HitWords w[4];
                uint32_t hm = 0;
                {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((ok & act & (1u << j)) && packed_match(cv[j], tk.key[j])) {
                            if (STAGE) {
                                stage[stage_at(4u * lane + j)] = cv[j];
                            } else {
                                w[j].y = cv[j].y;
                                w[j].z = cv[j].z;
                                w[j].w = cv[j].w;
                            }
                            hm |= 1u << j;
                        }
                    }
                }
                my_chain += __popc(hm);
                uint32_t need = act & ~hm;
                if (tv.occupied) {  // a window whose home slot is empty is a miss
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((need & (1u << j)) && !((bw[j] >> (h[j] & 31u)) & 1u)) need &= ~(1u << j);
                }

                // ---- what is left: lookup_hash_entry (kguts.cc:585-602).  The windows left are few (a tenth) and scattered over
                //      the lanes, so they are queued in shared memory and probed one per lane: one dense pass instead of four
                //      sparse ones, and one slot in registers instead of four. ----
                if (__any_sync(full, need != 0u)) {
                    uint32_t qp[4], n_left = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b = __ballot_sync(full, (need >> j) & 1u);
                        qp[j] = n_left + __popc(b & lt);
                        n_left += __popc(b);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (need & (1u << j)) {
                            queue[qp[j]] = make_uint4((uint32_t)tk.key[j], (uint32_t)(tk.key[j] >> 32) | ((4u * lane + j) << 8), h[j], bw[j]);
                            if (STAGE) stage[stage_at(4u * lane + j)].y = 0x8u;  // "no hit" until the probe says otherwise
                        }
                    }
                    __syncwarp();
                    for (uint32_t k = lane; k < n_left; k += 32u) {
                        const uint4 it = queue[k];
                        uint32_t hh = it.z;
                        uint4 v = t_tab ? ldg_v4_hint(slots + hh, pol_first) : __ldg(slots + hh);
                        // a window gets here because its home slot is taken -- nearly always by another k-mer, so the probe
                        // sequence goes on: when the same occupancy word says the next slot is taken too, fetch it now
                        if ((hh & 31u) != 31u && hh + 1u < nsig && ((it.w >> ((hh & 31u) + 1u)) & 1u) && tv.occupied) {
                            const uint4 v1 = __ldg(slots + hh + 1u);
                            if (!(v.x == it.x && (v.y & 0xFu) == (it.y & 0xFFu)) && !(v.y & 0x8u)) {
                                v = v1;
                                hh++;
                            }
                        }
                        uint32_t found = 0;
                        for (;;) {
                            if (v.x == it.x && (v.y & 0xFu) == (it.y & 0xFFu)) { found = 1u; break; }
                            if (v.y & 0x8u) break;
                            hh = (hh + 1u == nsig) ? 0u : hh + 1u;
                            if (hh == it.z) break;  // a table without an empty slot
                            if (tv.occupied) {
                                const uint32_t ow = (hh >> 5) == (it.z >> 5) ? it.w : __ldg(tv.occupied + (hh >> 5));
                                if (!((ow >> (hh & 31u)) & 1u)) break;
                            }
                            v = __ldg(slots + hh);
                        }
                        if (STAGE) {
                            if (found) stage[stage_at(it.y >> 8)] = v;
                        } else {
                            queue[k] = make_uint4(found, v.y, v.z, v.w);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (need & (1u << j)) {
                            if (STAGE) {
                                if (!(stage[stage_at(4u * lane + j)].y & 0x8u)) hm |= 1u << j;
                            } else {
                                const uint4 r = queue[qp[j]];
                                if (r.x) {
                                    w[j].y = r.y;
                                    w[j].z = r.z;
                                    w[j].w = r.w;
                                    hm |= 1u << j;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                my_probes += __popc(act);

                // ---- ordered compaction: exclusive prefix of per-lane hit counts ----
                const uint32_t cnt = __popc(hm);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(full, incl, d);
                    if (lane >= (uint32_t)d) incl += t;
                }
                const uint32_t tile_hits = __shfl_sync(full, incl, 31);
                uint32_t o = count + incl - cnt;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (hm & (1u << j)) {
                        HitWords r;
                        uint64_t key;
                        if (STAGE) {
                            const uint4 e = stage[stage_at(4u * lane + j)];
                            r.y = e.y;
                            r.z = e.z;
                            r.w = e.w;
                            key = (uint64_t)e.x | ((uint64_t)(e.y & 0x7u) << 32);
                        } else {
                            r = w[j];
                            key = tk.key[j];
                        }
                        HitRec rec;
                        rec.pos = q0 + j;
                        rec.fI = r.w & (kPackedFieldLimit - 1);
                        rec.wt = __uint_as_float(r.z);
                        rec.oI = (int32_t)(((r.y >> 20) & 0xFFFu) | ((r.w >> 22) << 12)) - 1;
                        if (t_st) stg_v4_hint(reinterpret_cast<uint4 *>(out + o), *reinterpret_cast<const uint4 *>(&rec), pol_first);
                        else out[o] = rec;
                        if (hit_keys) hit_keys[seq_base + o] = key;
                        if (hit_avg) hit_avg[seq_base + o] = (uint16_t)((r.y >> 4) & 0xFFFFu);
                        o++;
                    }
                }
                count += tile_hits;
            }
        }
        if (lane == 0) n_hits[i] = count;
        if (lane == 0) my_hits += count;
    }

    // batch totals: one atomic per warp (totals[4] = hits answered from the neighbour copy)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_probes += __shfl_down_sync(full, my_probes, d);
        my_hits += __shfl_down_sync(full, my_hits, d);
        my_chain += __shfl_down_sync(full, my_chain, d);
    }
    if (lane == 0) {
        atomicAdd(totals + 0, (unsigned long long)my_probes);
        atomicAdd(totals + 1, (unsigned long long)my_hits);
        atomicAdd(totals + 4, (unsigned long long)my_chain);
    }
}

}  // namespace ckm
