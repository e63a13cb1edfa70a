// Neighbour copy of the signature table, and the K1 instantiation that probes through it.
//
// The hash table of kmer_image.h scatters k-mers that were consecutive in their source protein over the whole image, so
// every hit of lookup_hash_entry (kguts.cc:585-602) is its own random DRAM transaction -- and the number of independent
// DRAM transactions per second, not bytes, is what bounds probe_kernel on this part.  But a query that
// hits a signature k-mer x1..x8 very often hits x2..x9 at its next position: the query is a homologue of the protein the
// signatures were cut from (build_signature_kmers walks proteins position by position).  This file rebuilds that
// adjacency from the image alone and stores a second, *neighbour-ordered* copy of the occupied slots:
//
//   chain[]  16-byte packed slots (same bit layout as the table, ckm_common.cuh); k-mers that follow each other in a
//            chain of overlapping signature k-mers are adjacent, eight per 128-byte line;
//   cpos[h]  for table slot h, the index of its copy in chain[].
//
// A lookup is a pure function of the key (first match in probe order wins; only slots a lookup can end at are
// copied), so answering a probe from chain[] -- key compared in full -- returns exactly the fields the hash probe
// returns.  probe_chain_kernel keeps, per protein, the difference `base` = chain index - position of the last
// chain-resolved hit, and for the 128 windows of a warp step
//   A. compares each window whose home slot is occupied (L2 bitmap) with chain[base + position]: one coalesced 2 KB read
//      per warp instead of up to 128 random transactions;
//   B. hash-probes only the first window of every run of still-unresolved windows ("anchors"), reads cpos[] of the anchors
//      that hit, and lets the windows behind an anchor try chain[anchor index + distance]; repeats on what is left
//      (every pass resolves at least the anchors; after three passes everything left is hash-probed).
// A protein whose hits do not follow chains (anchor hits outnumbering chain-resolved hits 2:1) drops back to plain hash
// probing for its remaining steps.  Results are bit-identical to probe_kernel by construction; tests run both.
//
// Building the chains (once per image, on the device):
//   1. every slot a lookup can end at looks up its 20 possible successors (x2..x8 + c: consecutive keys, hence
//      consecutive home buckets) and proposes to the one with the same function index whose avg_from_end is closest to
//      its own minus one; a k-mer proposed to by several keeps the best proposer (atomicMin on (distance, slot));
//   2. accepted proposals form disjoint paths (and, rarely, cycles: ACACACAC <-> CACACACA); pointer jumping over
//      (root, distance) pairs ranks every path; members of cycles become chains of one;
//   3. chain lengths -> exclusive prefix sum -> chain index = start[root] + distance.
#pragma once
#include "ckm_probe.cuh"

namespace ckm {

constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr uint64_t kNoPd = 0xFFFFFFFFFFFFFFFFull;  // pd[] entry of a slot that is not in any chain (empty or unreachable)
constexpr uint64_t kPow20_7 = 1280000000ull;

__device__ __forceinline__ bool packed_match(const uint4 &v, uint64_t key) {
    return v.x == (uint32_t)key && (v.y & 0xFu) == (uint32_t)(key >> 32);  // an empty slot has bit 3 set: never equal
}

// lookup_hash_entry (kguts.cc:585-602) over the packed table: slot index of `key`, or kNoSlot
__device__ __forceinline__ uint32_t packed_find(const TableView &tv, uint64_t key, uint4 &v) {
    uint64_t h = table_home(tv, key);
    for (uint64_t guard = 0; guard < tv.num_sigs; guard++) {
        if (tv.occupied && !((__ldg(tv.occupied + (h >> 5)) >> (h & 31u)) & 1u)) return kNoSlot;
        v = __ldg(reinterpret_cast<const uint4 *>(tv.slots) + h);
        if (packed_match(v, key)) return (uint32_t)h;
        if (v.y & 0x8u) return kNoSlot;
        h = (h + 1 == tv.num_sigs) ? 0 : h + 1;
    }
    return kNoSlot;
}

// step 1: membership (pd[h] = (h, 0) for slots a lookup can end at), best successor, proposal
__global__ void __launch_bounds__(256)
chain_propose_kernel(TableView tv, uint64_t *__restrict__ pd, uint32_t *__restrict__ succ, unsigned long long *__restrict__ claim) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= tv.num_sigs) return;
    const uint4 me = __ldg(reinterpret_cast<const uint4 *>(tv.slots) + h);
    uint64_t mine = kNoPd;
    uint32_t best = kNoSlot;
    if (!(me.y & 0x8u)) {
        const uint64_t key = (uint64_t)me.x | ((uint64_t)(me.y & 0x7u) << 32);
        uint4 t;
        // the reference's builder never checks for duplicates (kguts.cc:202-222): a repeated k-mer sits in a second slot
        // that no lookup reaches.  Only reachable slots are copied.
        if (packed_find(tv, key, t) == (uint32_t)h) {
            mine = h;  // root = itself, distance 0
            const uint32_t fI = me.w & (kPackedFieldLimit - 1);
            const int want = (int)((me.y >> 4) & 0xFFFFu) - 1;
            const uint64_t s0 = (key % kPow20_7) * 20ull;
            uint32_t bestd = 0xFFFFFFFFu;
            for (uint32_t c = 0; c < 20u; c++) {
                if (s0 + c == key) continue;
                const uint32_t s = packed_find(tv, s0 + c, t);
                if (s == kNoSlot || (t.w & (kPackedFieldLimit - 1)) != fI) continue;
                const int a = (int)((t.y >> 4) & 0xFFFFu);
                const uint32_t d = (uint32_t)(a > want ? a - want : want - a);
                if (d < bestd) {
                    bestd = d;
                    best = s;
                }
            }
            if (best != kNoSlot) atomicMin(claim + best, ((unsigned long long)bestd << 32) | (unsigned long long)h);
        }
    }
    pd[h] = mine;
    succ[h] = best;
}

// step 2a: the proposal a successor kept becomes its parent pointer (distance 1)
__global__ void __launch_bounds__(256)
chain_link_kernel(uint64_t num_sigs, const uint32_t *__restrict__ succ, const unsigned long long *__restrict__ claim,
                  uint64_t *__restrict__ pd) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= num_sigs) return;
    const uint32_t s = succ[h];
    if (s != kNoSlot && (uint32_t)claim[s] == (uint32_t)h) pd[s] = h | (1ull << 32);
}

// step 2b: pointer jumping in place.  A 64-bit entry (parent, distance to parent) is read and written whole, so any
// interleaving composes true statements "h is d behind p"; roots are fixed points.
__global__ void __launch_bounds__(256)
chain_jump_kernel(uint64_t num_sigs, volatile uint64_t *pd, unsigned int *__restrict__ changed) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= num_sigs) return;
    const uint64_t x = pd[h];
    const uint32_t p = (uint32_t)x;
    if (x == kNoPd || x == h) return;  // not a member, or a root: (itself, distance 0)
    const uint64_t y = pd[p];
    if (y == (uint64_t)p) return;  // parent is a root
    // (on a cycle whose length is a power of two the pointer comes back to h itself with a non-zero distance: not a
    // root, keeps "changing", and is cut below)
    pd[h] = (uint64_t)(uint32_t)y | (((x >> 32) + (y >> 32)) << 32);
    *changed = 1u;
}

// step 2c (only when jumping did not converge): members whose pointer does not end at a root sit on a cycle
__global__ void __launch_bounds__(256)
chain_cycle_mark_kernel(uint64_t num_sigs, const uint64_t *__restrict__ pd, uint32_t *__restrict__ mark) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= num_sigs) return;
    const uint64_t x = pd[h];
    uint32_t m = 0;
    if (x != kNoPd && x != h) {
        const uint32_t p = (uint32_t)x;
        m = (pd[p] != (uint64_t)p) ? 1u : 0u;
    }
    mark[h] = m;
}
__global__ void __launch_bounds__(256)
chain_cycle_cut_kernel(uint64_t num_sigs, const uint32_t *__restrict__ mark, uint64_t *__restrict__ pd) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= num_sigs) return;
    if (mark[h]) pd[h] = h;
}

// step 3a: chain length at its root
__global__ void __launch_bounds__(256)
chain_length_kernel(uint64_t num_sigs, const uint64_t *__restrict__ pd, uint32_t *__restrict__ len) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= num_sigs) return;
    const uint64_t x = pd[h];
    if (x == kNoPd) return;
    atomicMax(len + (uint32_t)x, (uint32_t)(x >> 32) + 1u);
}

// step 3a': every chain is followed by kChainPad unused indices, so that the residue string of the compact copy (below) can
// carry the last k-mer's seven trailing residues
constexpr uint32_t kChainPad = CKM_KMER_SIZE - 1;
__global__ void __launch_bounds__(256) chain_pad_kernel(uint64_t num_sigs, uint32_t *__restrict__ len) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < num_sigs && len[h]) len[h] += kChainPad;
}

// step 3c: the compact form of the copy that probe_pc_kernel reads.  Consecutive members of a chain overlap by seven
// residues (a successor is one of the twenty k-mers x2..x8+c, step 1), so a chain IS a residue string and its k-mers are the
// string's windows: cres[i] = first residue of the k-mer at index i (bit 7 set: a k-mer starts here), followed after the
// chain's last member by that k-mer's other seven residues; cpay[i] = (function_wt bits, word 3 of the packed slot).
// "The query window equals the k-mer at index i" becomes an 8-byte string comparison against 1 + 8 bytes per window instead of
// a key comparison against 16.
constexpr uint8_t kCresNone = 0x7F;  // no residue (never equals a residue code)
__global__ void __launch_bounds__(256)
chain_compact_kernel(TableView tv, const uint64_t *__restrict__ pd, const uint64_t *__restrict__ start, const uint32_t *__restrict__ len,
                     uint8_t *__restrict__ cres, uint2 *__restrict__ cpay) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= tv.num_sigs) return;
    const uint64_t x = pd[h];
    if (x == kNoPd) return;
    const uint32_t root = (uint32_t)x, d = (uint32_t)(x >> 32);
    const uint64_t pos = start[root] + d;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(tv.slots) + h);
    uint64_t key = (uint64_t)v.x | ((uint64_t)(v.y & 0x7u) << 32);
    uint8_t r[CKM_KMER_SIZE];
#pragma unroll
    for (int t = CKM_KMER_SIZE - 1; t >= 0; t--) {
        r[t] = (uint8_t)(key % 20ull);
        key /= 20ull;
    }
    cres[pos] = r[0] | 0x80u;
    cpay[pos] = make_uint2(v.z, v.w);
    if (d + 1u + kChainPad == len[root]) {
#pragma unroll
        for (int t = 1; t < CKM_KMER_SIZE; t++) cres[pos + t] = r[t];
    }
}

// step 3b: copy every member to chain[start[root] + distance]
__global__ void __launch_bounds__(256)
chain_place_kernel(TableView tv, const uint64_t *__restrict__ pd, const uint64_t *__restrict__ start, uint4 *__restrict__ chain,
                   uint2 *__restrict__ cpos, unsigned long long *__restrict__ n_roots) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool root = false;
    if (h < tv.num_sigs) {
        const uint64_t x = pd[h];
        uint32_t pos = kNoSlot, tag = 0xFFFFFFFFu;
        if (x != kNoPd) {
            pos = (uint32_t)(start[(uint32_t)x] + (x >> 32));
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(tv.slots) + h);
            chain[pos] = v;
            tag = v.x;
            root = x == h;
        }
        cpos[h] = make_uint2(tag, pos);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, root);
    if ((threadIdx.x & 31u) == 0 && b) atomicAdd(n_roots, (unsigned long long)__popc(b));
}

// ---------------------------------------------------------------------------------------------------
// K1 through the neighbour copy (packed slots only)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr uint32_t kChainPasses = 3;  // anchor passes per warp step before everything left is hash-probed

struct HitWords {  // words 1..3 of the packed slot behind a hit (word 0 is the low half of the key)
    uint32_t y, z, w;
};

#ifdef CKM_EXPERIMENTS  // the walking kernel: kept for A/B runs only (slower than plain probing)
template <int MINB>
__global__ void __launch_bounds__(kProbeThreads, MINB)
probe_chain_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n,
                   HitRec *__restrict__ hits, uint64_t *__restrict__ hit_keys, uint16_t *__restrict__ hit_avg,
                   uint32_t *__restrict__ n_hits, unsigned long long *__restrict__ totals) {
    __shared__ uint8_t lut[256];
    fill_aa_lut(lut);
    __syncthreads();

    constexpr uint32_t full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint4 *__restrict__ slots = reinterpret_cast<const uint4 *>(tv.slots);
    const uint32_t nsig = (uint32_t)tv.num_sigs;  // the neighbour copy is only built for tables below 2^32 buckets
    uint32_t my_probes = 0, my_hits = 0, my_chain = 0;

    for (uint32_t i = warp0; i < n; i += n_warps) {
        const uint64_t seq_base = __ldg(offsets + i);
        const uint32_t len = (uint32_t)(__ldg(offsets + i + 1) - seq_base);
        uint32_t count = 0;
        if (len > CKM_KMER_SIZE) {
            uint32_t nwin = len - CKM_KMER_SIZE;  // the last window is never probed (kguts.cc:792, 798)
            const uint8_t *p0 = residues + seq_base;
            const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
            const uint32_t *wb = reinterpret_cast<const uint32_t *>(p0 - s);
            const uint32_t nwords = (len + s + 3u) >> 2;
            const uint32_t sh = 8u * s;
            HitRec *out = hits + seq_base;

            // per protein: chain index minus position of the last chain-resolved hit; whether chains pay for this protein
            bool has_base = false, use_chain = true;
            uint32_t base = 0, via_chain = 0, via_anchor = 0;

            for (uint32_t t0 = 0; t0 < nwin; t0 += kTile) {
                const TileKeys tk = tile_keys(lut, wb, nwords, sh, t0, lane, len, nwin);
                const uint32_t q0 = t0 + 4u * lane;
                const uint32_t act = tk.act;
                uint32_t h[4];
#pragma unroll
                for (int j = 0; j < 4; j++) h[j] = (uint32_t)table_home(tv, tk.key[j]);

                HitWords hv[4];       // the slot behind each hit
                uint32_t hm = 0;      // windows that hit
                int lj = -1;          // this lane's highest window whose chain index is known, and that index minus its position
                uint32_t lb = 0;
                uint32_t t_chain = 0, t_anchor = 0;

                // ---- occupancy bits (L2) and, in the same round trip, the chain entries the previous hits predict ----
                uint32_t need = act;
                uint32_t bw[4];
                const bool predict = has_base && use_chain;
                if (tv.occupied) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (act & (1u << j)) bw[j] = __ldg(tv.occupied + (h[j] >> 5));
                }
                if (predict) {
                    uint4 cv[4];
                    uint32_t ok = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t idx = base + q0 + j;
                        if ((act & (1u << j)) && idx < tv.n_chain) {
                            cv[j] = __ldg(tv.chain + idx);
                            ok |= 1u << j;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((ok & (1u << j)) && packed_match(cv[j], tk.key[j])) {
                            hv[j].y = cv[j].y;
                            hv[j].z = cv[j].z;
                            hv[j].w = cv[j].w;
                            hm |= 1u << j;
                            lj = j;
                            lb = base;
                            t_chain++;
                        }
                    }
                    need &= ~hm;
                }
                if (tv.occupied) {  // a window whose home slot is empty is a miss
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((need & (1u << j)) && !((bw[j] >> (h[j] & 31u)) & 1u)) need &= ~(1u << j);
                }

                // ---- anchors (hash probes) and the windows behind them ----
                uint32_t pass = 0;
                while (__any_sync(full, need != 0u)) {
                    const bool chain_pass = use_chain && pass < kChainPasses;
                    uint32_t anchors = need;
                    if (chain_pass) {  // first window of every run of unresolved windows
                        uint32_t prev = __shfl_up_sync(full, need >> 3, 1) & 1u;
                        if (lane == 0) prev = 0;
                        anchors = need & ~((need << 1) | prev);
                    }
                    // without a prediction to lose (first step of a protein) the anchors' chain indices are fetched together
                    // with their home slots; they are right whenever the match is at the home slot
                    const bool spec = chain_pass && !has_base;
                    uint32_t ab[4], home = 0;
                    if (spec) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (anchors & (1u << j)) ab[j] = __ldg(tv.cpos + h[j]).y;
                    }
                    // linear probing, the lane's windows side by side: h = (h+1) % size_hash until match or empty (kguts.cc:585-602)
                    uint32_t pend = anchors, ah = 0, steps = 0;
                    while (pend) {
                        uint4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (pend & (1u << j)) v[j] = __ldg(slots + h[j]);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (pend & (1u << j)) {
                                if (packed_match(v[j], tk.key[j])) {
                                    ah |= 1u << j;
                                    if (steps == 0) home |= 1u << j;
                                    hv[j].y = v[j].y;
                                    hv[j].z = v[j].z;
                                    hv[j].w = v[j].w;
                                    pend &= ~(1u << j);
                                } else if (v[j].y & 0x8u) {
                                    pend &= ~(1u << j);
                                } else {
                                    const uint32_t hn = (h[j] + 1u == nsig) ? 0u : h[j] + 1u;
                                    if (tv.occupied) {
                                        if ((hn >> 5) != (h[j] >> 5)) bw[j] = __ldg(tv.occupied + (hn >> 5));
                                        if (!((bw[j] >> (hn & 31u)) & 1u)) pend &= ~(1u << j);
                                    }
                                    h[j] = hn;
                                }
                            }
                        }
                        if (++steps >= nsig) pend = 0;  // a table without an empty slot
                    }
                    need &= ~anchors;
                    hm |= ah;
                    if (chain_pass) {
                        t_anchor += __popc(ah);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if ((ah & (1u << j)) && !(spec && (home & (1u << j)))) ab[j] = __ldg(tv.cpos + h[j]).y;
                        uint32_t last = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (ah & (1u << j)) {
                                ab[j] -= q0 + j;
                                last = ab[j];
                                if (j >= lj) {
                                    lj = j;
                                    lb = ab[j];
                                }
                            }
                        }
                        const uint32_t m = __ballot_sync(full, ah != 0u);
                        if (m && __any_sync(full, need != 0u)) {
                            // the nearest anchor hit in front of each unresolved window: in this lane, else the last one of
                            // the nearest lower lane that has any
                            const uint32_t lower = m & ((1u << lane) - 1u);
                            const uint32_t src = lower ? 31u - __clz(lower) : 0u;
                            uint32_t rb = __shfl_sync(full, last, src);
                            bool rv = lower != 0u;
                            uint4 cv[4];
                            uint32_t ok = 0;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (ah & (1u << j)) {
                                    rb = ab[j];
                                    rv = true;
                                } else if ((need & (1u << j)) && rv) {
                                    const uint32_t idx = rb + q0 + j;
                                    if (idx < tv.n_chain) {
                                        cv[j] = __ldg(tv.chain + idx);
                                        ab[j] = rb;
                                        ok |= 1u << j;
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if ((ok & (1u << j)) && packed_match(cv[j], tk.key[j])) {
                                    hv[j].y = cv[j].y;
                                    hv[j].z = cv[j].z;
                                    hv[j].w = cv[j].w;
                                    hm |= 1u << j;
                                    need &= ~(1u << j);
                                    if (j >= lj) {
                                        lj = j;
                                        lb = ab[j];
                                    }
                                    t_chain++;
                                }
                            }
                        }
                    }
                    pass++;
                }
                my_probes += __popc(act);

                // ---- carry the chain offset of the last chain-known hit of this step; does following chains pay? ----
                if (use_chain) {
                    const uint32_t m = __ballot_sync(full, lj >= 0);
                    if (m) {
                        base = __shfl_sync(full, lb, 31u - __clz(m));
                        has_base = true;
                    }
                    via_chain += __reduce_add_sync(full, t_chain);
                    via_anchor += __reduce_add_sync(full, t_anchor);
                    if (via_anchor > 2u * via_chain + 16u) use_chain = false;
                }

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
                        HitRec rec;
                        rec.pos = q0 + j;
                        rec.fI = hv[j].w & (kPackedFieldLimit - 1);
                        rec.wt = __uint_as_float(hv[j].z);
                        rec.oI = (int32_t)(((hv[j].y >> 20) & 0xFFFu) | ((hv[j].w >> 22) << 12)) - 1;
                        out[o] = rec;
                        if (hit_keys) hit_keys[seq_base + o] = tk.key[j];
                        if (hit_avg) hit_avg[seq_base + o] = (uint16_t)((hv[j].y >> 4) & 0xFFFFu);
                        o++;
                    }
                }
                count += tile_hits;
            }
            if (lane == 0) my_chain += via_chain;
        }
        if (lane == 0) n_hits[i] = count;
        if (lane == 0) my_hits += count;
    }

    // batch totals: one atomic per warp (totals[4] = hits answered from the neighbour copy)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_probes += __shfl_down_sync(full, my_probes, d);
        my_hits += __shfl_down_sync(full, my_hits, d);
    }
    if (lane == 0) {
        atomicAdd(totals + 0, (unsigned long long)my_probes);
        atomicAdd(totals + 1, (unsigned long long)my_hits);
        atomicAdd(totals + 4, (unsigned long long)my_chain);
    }
}

#endif  // CKM_EXPERIMENTS

}  // namespace ckm
