// K1 with the ordered scoring scan inside the SM: probing warps (producers) + one scan warp (consumer) per block.
//
// On the calls / best-call path (no hit list, no OTU statistics asked for) nothing but 20-byte calls and the 28-byte
// best-call record has to leave the SM, yet K1 + scan_kernel move every hit through HBM as a
// 16-byte record, written once and read once -- a third of K1's DRAM traffic and all of scan_kernel's 0.75 ms per C2 step.
// Running the transducer inside the probing warp does not pay either: the scan is serial per protein, a warp is one
// protein, so 31 lanes idle through ~600 extra warp instructions per 128-window step and the kernel outgrows the
// instruction cache (6.4 ms against 4.0 + 0.75; profiles/r2/fused_in_warp_*).  The transducer wants one THREAD per
// protein -- so here it gets one:
//
//   producers  P warps per block, each exactly probe_hint_kernel's step (ckm_hint.cuh: chain compare, occupancy words,
//              left-over hash probes), one protein at a time, claimed four at a time from a global counter.  The slot
//              behind every hit of a step is already parked in the warp's shared-memory stage; instead of compacting and
//              storing hit records the warp publishes the step: a 128-bit hit mask, the mask of hits that start a run
//              of one function, weight and function index of every hit compacted in position order, the step's first
//              position and, with the protein's last step, an end flag.  Four steps may be in flight per producer, so the
//              warp is never held up by a scan lane that is a step or two behind.
//   consumer   warp P.  Lane w serves producer w: it takes the published step run by run -- rs_hit (ckm_warp_scan.cuh) for the hits
//              that can change the run state, a plain ordered f32 sum for the rest -- and at the end flag does the final
//              flush and find_best_call, and writes calls / n_calls / best.  P proteins are scanned side
//              by side by one instruction stream -- scan_kernel's thread-per-protein shape without the trip through HBM.
//
// Hand-off is two counters per producer in shared memory (published / consumed, written by one side each) with
// __threadfence_block() on both sides; a producer only ever waits for its own lane of the consumer and that lane only for
// its producer, so there is no cycle to deadlock on.  Spins are bounded all the same: a stuck hand-off raises totals[7] and
// the host call fails with CKM_ECUDA instead of hanging the device.
//
// hints == nullptr (no neighbour copy, or it is suspended): every window takes the hash probe; same kernel.
#pragma once
#include "ckm_hint.cuh"
#include "ckm_warp_scan.cuh"

namespace ckm {

constexpr uint32_t kPcEnd = 1u;     // the protein ends with this record
constexpr uint32_t kPcFin = 2u;     // the producer has no more work
constexpr uint32_t kPcOneRun = 4u;  // all hits of the step share one function index
constexpr int kPcProducers = 31;    // probing warps per block: 31 + 1 scan warp = 1 024 threads x 64 registers fill an SM
constexpr uint32_t kPcChunk = 4u;   // sequences claimed per atomic
constexpr uint32_t kPcDepth = 2u;   // published steps a producer may be ahead of its scan lane
constexpr uint32_t kPcSpinLimit = 1u << 21;  // polls (a few tenths of a second) before a hand-off is declared stuck

// One published step.  `hits` bit e: window t0 + e hit.  `starts` is the subset of `hits` that begin a run: the step's first
// hit and every hit whose function index differs from the hit before it.  The payload of the step -- weight and function index
// of every hit, in position order -- is in the W / FI arrays of the same slot (record number % kPcDepth).
struct __align__(16) PcRecord {
    uint32_t hits[4];
    uint32_t starts[4];
    uint32_t t0, flags, index, n;
    uint64_t seq_base, pad_;
};
struct __align__(16) PcSync {
    volatile uint32_t pub;    // records published (written by the producer)
    volatile uint32_t done;   // records consumed (written by the scan lane)
    volatile uint32_t abort;  // entry 0 only: a hand-off timed out somewhere in the block, nobody waits any more
    uint32_t pad_;
    uint32_t found[4];        // producer-private: left-over windows of the step whose hash probe hit
};
static_assert(sizeof(PcRecord) == 64 && sizeof(PcSync) == 32, "shared-memory layout");

// Dynamic shared memory of a block with P producers:
//   residue LUT | sync[P] | records[P][D] | queue[P][128] (8 B: key, window) | stage[P][128] (8 B: weight, function word) | W[P][D][128] | FI[P][D][128]
//   | land[P][2][20] (residue string of the neighbour copy behind the step)
// 4.2 KB per producer, as much as probe_hint_kernel uses per warp -- on purpose: what shared memory takes, the L1 loses, and
// this kernel lives on its L1 (with the carve-out at its maximum probe_hint_kernel itself takes 7.1 ms per C2 step instead
// of 4.0; profiles/r2/carveout_ab.jsonl).
constexpr uint32_t kPcWSlot = kTile + 4;  // weights of a step + zero padding (the scan lane adds them four at a time)
constexpr uint32_t kPcLandWords = 20;  // residue bytes of the neighbour copy behind one 64-window half of a step: 64 + 7 (+3 of alignment)
constexpr size_t kPcPerProducer = sizeof(PcSync) + kPcDepth * sizeof(PcRecord) + kTile * sizeof(uint2) + kTile * sizeof(uint2) +
                                  kPcDepth * (kPcWSlot + kTile) * 4 + 2 * kPcLandWords * 4;
constexpr size_t pc_smem_bytes(int P) { return 256 + (size_t)P * kPcPerProducer; }

// floor(off / max(1, min_hits)) by a multiply: exact for off < 2^64 / min_hits (call_region_base, ckm_scan.cuh)
static inline uint64_t call_region_magic(int min_hits) { return min_hits > 1 ? ~0ull / (uint64_t)min_hits + 1ull : 0ull; }
__device__ __forceinline__ uint64_t call_region_base_fast(uint64_t off, uint32_t i, uint64_t magic) {
    return (magic ? __umul64hi(off, magic) : off) + i;
}

// stage[] index of a step's window e: 8-byte entries, read four consecutive ones per lane (publish) and written 32 consecutive
// ones per instruction (the asynchronous copy); XOR-ing the low four bits with the next three keeps both conflict-free
__device__ __forceinline__ uint32_t stage8_at(uint32_t e) { return e ^ ((e >> 4) & 15u); }

// 128-bit window masks of a step
struct M128 {
    uint64_t lo, hi;
};
__device__ __forceinline__ bool m_any(const M128 &m) { return (m.lo | m.hi) != 0ull; }
__device__ __forceinline__ uint32_t m_first(const M128 &m) { return m.lo ? __ffsll((long long)m.lo) - 1 : 63 + __ffsll((long long)m.hi); }
__device__ __forceinline__ uint32_t m_last(const M128 &m) { return m.hi ? 127 - __clzll((long long)m.hi) : 63 - __clzll((long long)m.lo); }
__device__ __forceinline__ uint32_t m_count(const M128 &m) { return __popcll(m.lo) + __popcll(m.hi); }
__device__ __forceinline__ M128 m_below(uint32_t n) {  // bits [0, n), n <= 128
    M128 r;
    r.lo = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
    r.hi = n <= 64u ? 0ull : (n >= 128u ? ~0ull : ((1ull << (n - 64u)) - 1ull));
    return r;
}
__device__ __forceinline__ M128 m_and(const M128 &a, const M128 &b) { return M128{a.lo & b.lo, a.hi & b.hi}; }
__device__ __forceinline__ M128 m_andnot(const M128 &a, const M128 &b) { return M128{a.lo & ~b.lo, a.hi & ~b.hi}; }
__device__ __forceinline__ void m_drop_first(M128 &m) {
    if (m.lo) m.lo &= m.lo - 1ull;
    else m.hi &= m.hi - 1ull;
}
__device__ __forceinline__ void m_drop_last(M128 &m) {
    if (m.hi) m.hi &= ~(1ull << (63 - __clzll((long long)m.hi)));
    else m.lo &= ~(1ull << (63 - __clzll((long long)m.lo)));
}

// The scan warp.  Lane w < P serves producer w.  A record is taken run by run: the first hits of a run go through rs_hit one
// at a time until the run's function is the current one (two hits at most: kguts.cc:852-856 flushes on the second and the
// carry-over makes it current); from there on every further hit of the run does nothing but num++, count++, weight added,
// last position -- applied in bulk, the weights added in order.  (With max_gap < 127 a gap can fall between two hits of one
// step, and every hit takes rs_hit.)  find_best_call of a protein with a single call is made here from registers; proteins
// with several calls (a tenth of C2's) are left to best_fixup_kernel so that their long, divergent walk does not hold up
// the other lanes.  What counts here is the length of the longest path through one iteration, not the work per record: the lanes
// run in lockstep and a producer waits for its lane once it is two steps ahead (a separate shorter path for one-run
// records made the iteration longer -- both paths are walked whenever one lane needs the general one -- and K1 8 % slower;
// more scan warps cost probing warps, 1.6 % each: profiles/r2/k1_experiments.md).
// first / count: the producers this scan warp serves
__device__ __forceinline__ void pc_consume(PcSync *syncs, const PcRecord *recs, const float *wts, const uint32_t *fis, uint32_t lane,
                                           uint32_t first, uint32_t count, uint32_t index_base, const FusedArgs &fa,
                                           unsigned long long *totals) {
    constexpr uint32_t full = 0xffffffffu;
    const bool mine = lane < count;
    const uint32_t w = first + (mine ? lane : 0u);
    PcSync *sy = syncs + w;
    const PcRecord *recD = recs + kPcDepth * w;
    const float *wtsD = wts + (size_t)w * kPcDepth * kPcWSlot;
    const uint32_t *fisD = fis + (size_t)w * kPcDepth * kTile;
    const float *W = wtsD;
    const uint32_t *FI = fisD;
    const bool per_hit = fa.prm.max_gap < kTile - 1;
    bool fin = !mine, have = false, at_start = true;
    uint32_t seen = 0, t0 = 0, flags = 0, index = 0, n = 0, k = 0, my_calls = 0, idle = 0;
    M128 H = {0ull, 0ull}, R = {0ull, 0ull};
    ckm_call_t *calls = nullptr;
    RunState S;
    rs_begin(S);
    while (!__all_sync(full, fin)) {
        if (!fin && !have && sy->pub != seen) {
            __threadfence_block();
            const uint32_t slot = seen & (kPcDepth - 1u);
            const PcRecord *r = recD + slot;
            const uint4 mh = *reinterpret_cast<const uint4 *>(r->hits);
            const uint4 ms = *reinterpret_cast<const uint4 *>(r->starts);
            const uint4 q = *reinterpret_cast<const uint4 *>(&r->t0);
            H.lo = (uint64_t)mh.x | ((uint64_t)mh.y << 32);
            H.hi = (uint64_t)mh.z | ((uint64_t)mh.w << 32);
            R.lo = (uint64_t)ms.x | ((uint64_t)ms.y << 32);
            R.hi = (uint64_t)ms.z | ((uint64_t)ms.w << 32);
            t0 = q.x;
            flags = q.y;
            index = q.z;
            n = q.w;
            k = 0;
            W = wtsD + slot * kPcWSlot;
            FI = fisD + slot * kTile;
            if (flags & kPcFin) {
                fin = true;
            } else {
                have = true;
                if (at_start) {
                    calls = fa.calls + call_region_base_fast(r->seq_base, index_base + index, fa.call_magic);
                    at_start = false;
                }
            }
        }
        if (!__any_sync(full, have)) {  // nothing published anywhere: leave the issue slots to the producers
            if (++idle > kPcSpinLimit || syncs[0].abort) {
                if (lane == 0) atomicExch(totals + 7, 1ull);
                syncs[0].abort = 1u;
                break;
            }
            __nanosleep(100);
            continue;
        }
        idle = 0;
        if (have) {
            if (k < n) {  // ---- one run ----
                M128 run = H;
                uint32_t L = n;
                if (!(flags & kPcOneRun)) {
                    const uint32_t s = m_first(H);
                    R = m_andnot(R, m_below(s + 1u));
                    const uint32_t nxt = m_any(R) ? m_first(R) : (uint32_t)kTile;
                    run = m_and(H, m_below(nxt));
                    H = m_andnot(H, run);
                    L = m_count(run);
                }
                const uint32_t G = FI[k];
                uint32_t i = 0;
                do {
                    const uint32_t e = m_first(run);
                    m_drop_first(run);
                    rs_hit(S, fa.prm, t0 + e, G, W[k + i], calls);
                    i++;
                } while (i < L && (per_hit || !(S.num > 0 && S.cur_fI == G)));
                if (i < L) {  // the rest of the run in bulk
                    const uint32_t r = L - i;
                    float ws = S.wsum;
                    uint32_t x = k + i;
                    const uint32_t xe = k + L;
                    for (; (x & 3u) && x < xe; x++) ws += W[x];
                    for (; x + 4u <= xe; x += 4u) {
                        const float4 v = *reinterpret_cast<const float4 *>(W + x);
                        ws += v.x;
                        ws += v.y;
                        ws += v.z;
                        ws += v.w;
                    }
                    for (; x < xe; x++) ws += W[x];
                    S.wsum = ws;
                    S.num += r;
                    S.fI_count += (int32_t)r;
                    const uint32_t pl = t0 + m_last(run);
                    S.last_match_pos = pl;
                    if (r >= 2u) {
                        m_drop_last(run);
                        S.p2_pos = t0 + m_last(run);
                        S.p2_fI = G;
                        S.p2_wt = W[xe - 2u];
                    } else {
                        S.p2_pos = S.p1_pos;
                        S.p2_fI = S.p1_fI;
                        S.p2_wt = S.p1_wt;
                    }
                    S.p1_pos = pl;
                    S.p1_fI = G;
                    S.p1_wt = W[xe - 1u];
                }
                k += L;
            }
            if (k >= n) {
                if (flags & kPcEnd) {
                    if ((int)S.num >= fa.prm.min_hits) rs_flush(S, fa.prm, calls);
                    const uint32_t nc = S.n_calls;
                    fa.n_calls[index] = nc;
                    if (fa.best && nc <= 1u) {  // several calls: best_fixup_kernel
                        ckm_best_t b;
                        b.function_index = -1;
                        b.ambig_a = b.ambig_b = -1;
                        b.flags = 0;
                        b.score = b.weighted_score = b.score_offset = 0.0f;
                        if (nc) {
                            b.flags = CKM_BEST_HAS_CALLS;
                            Top2 top;
                            top.n = 0;
                            const FScore fs = {(int)S.c_fI, S.c_count, S.c_weighted};
                            top.push(fs);
                            best_from_top2(top, b);
                        }
                        fa.best[index] = b;
                    }
                    my_calls += nc;
                    rs_begin(S);
                    at_start = true;
                }
                __threadfence_block();  // the payload reads above come before the release of the slot
                sy->done = ++seen;
                have = false;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_calls += __shfl_down_sync(full, my_calls, d);
    if (lane == 0 && my_calls) atomicAdd(totals + 2, (unsigned long long)my_calls);
}

// find_best_call for the proteins probe_pc_kernel left out: those with more than one call
__global__ void __launch_bounds__(256)
best_fixup_kernel(const uint64_t *__restrict__ offsets, uint32_t n, uint32_t index_base, FusedArgs fa) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t nc = fa.n_calls[i];
    if (nc <= 1u) return;
    const uint64_t base = call_region_base_fast(__ldg(offsets + i), index_base + i, fa.call_magic);
    ckm_best_t b;
    find_best_call_dev(fa.calls + base, nc, fa.calls_work + base, b);
    fa.best[i] = b;
}

// P producer warps + C scan warps per block; dynamic shared memory pc_smem_bytes(P).  `work` is the next unclaimed sequence of
// this launch (zero at launch).  totals: [0] += probes, [1] += hits, [2] += calls, [4] += hits answered from the neighbour
// copy, [7] = hand-off failure.
template <int P, int C>
__global__ void __launch_bounds__((P + C) * 32, 1)
probe_pc_kernel(TableView tv, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n, uint32_t index_base,
                const uint32_t *__restrict__ hints, uint32_t *__restrict__ n_hits, unsigned long long *__restrict__ totals, FusedArgs fa,
                unsigned long long *__restrict__ work) {
    extern __shared__ __align__(16) uint8_t pc_smem[];
    uint8_t *lut = pc_smem;
    PcSync *syncs = reinterpret_cast<PcSync *>(pc_smem + 256);
    PcRecord *recs = reinterpret_cast<PcRecord *>(syncs + P);
    uint2 *queues = reinterpret_cast<uint2 *>(recs + kPcDepth * P);
    uint2 *stages = queues + P * kTile;
    float *wts = reinterpret_cast<float *>(stages + P * kTile);
    uint32_t *fis = reinterpret_cast<uint32_t *>(wts + (size_t)P * kPcDepth * kPcWSlot);
    uint32_t *lands = fis + (size_t)P * kPcDepth * kTile;
    fill_aa_lut(lut);
    if (threadIdx.x < P) {
        syncs[threadIdx.x].pub = 0;
        syncs[threadIdx.x].done = 0;
        syncs[threadIdx.x].abort = 0;
    }
    __syncthreads();

    constexpr uint32_t full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#ifdef CKM_EXPERIMENTS  // ablation (wrong results on purpose): no scan warp, nobody waits -- the probing warps' own pace
    const bool no_scan = tv.tuning & 0x400000u, no_left = tv.tuning & 0x800000u, no_pub = tv.tuning & 0x1000000u;
    if (no_scan && warp >= (uint32_t)P) return;
#else
    constexpr bool no_scan = false, no_left = false, no_pub = false;
#endif
    if (warp >= (uint32_t)P) {  // scan warp c serves producers [c * per, (c + 1) * per)
        constexpr uint32_t per = (P + C - 1) / C;
        const uint32_t first = (warp - P) * per;
        pc_consume(syncs, recs, wts, fis, lane, first, first < (uint32_t)P ? min(per, (uint32_t)P - first) : 0u, index_base, fa, totals);
        return;
    }

    PcSync *sy = syncs + warp;
    PcRecord *recD = recs + kPcDepth * warp;
    uint2 *queue = queues + warp * kTile;  // left-over windows of the step: (key low word, key high bits | window << 8)
    uint2 *stage = stages + warp * kTile;  // (weight bits, word 3 of the packed slot) behind each window of the step that hit
    float *wtsD = wts + (size_t)warp * kPcDepth * kPcWSlot;
    uint32_t *fisD = fis + (size_t)warp * kPcDepth * kTile;
    uint32_t *land = lands + (size_t)warp * 2 * kPcLandWords;
    uint32_t pubc = 0;  // records this warp has published
    const uint32_t lt = (1u << lane) - 1u;
    const uint4 *__restrict__ slots = reinterpret_cast<const uint4 *>(tv.slots);
    const uint32_t nsig = (uint32_t)tv.num_sigs;
    uint32_t my_probes = 0, my_hits = 0, my_chain = 0;
    const uint64_t pol_first = policy_evict_first();
    const uint32_t m35 = tv.m35;

    // hm: this lane's windows that hit (their slots are in `stage`).  Publishes the step -- hit mask, run starts, weights and
    // function indices in position order -- into slot (pubc & (kPcDepth - 1u)) once the scan lane has released it.  Returns the
    // step's hit count.
    auto publish = [&](uint32_t hm, uint32_t t0, uint32_t flags, uint32_t index, uint64_t seq_base) -> uint32_t {
        if (pubc >= kPcDepth && !no_scan) {  // record pubc - kPcDepth consumed?
            uint32_t spins = 0;
            while ((int32_t)(sy->done - (pubc - kPcDepth + 1u)) < 0) {
                if (++spins > kPcSpinLimit || syncs[0].abort) {
                    if (lane == 0) atomicExch(totals + 7, 1ull);
                    syncs[0].abort = 1u;
                    break;
                }
                __nanosleep(200);
            }
            __threadfence_block();
        }
        const uint32_t slot = pubc & (kPcDepth - 1u);
        PcRecord *r = recD + slot;
        float *W = wtsD + slot * kPcWSlot;
        uint32_t *FI = fisD + slot * kTile;
        uint32_t n_step = 0, rs = 0, n_runs = 0;
        const uint32_t anyb = __ballot_sync(full, hm != 0u);
        if (anyb) {
            uint32_t fi[4] = {0u, 0u, 0u, 0u}, wz[4] = {0u, 0u, 0u, 0u}, lastf = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (hm & (1u << j)) {
                    const uint2 zw = stage[stage8_at(4u * lane + j)];
                    wz[j] = zw.x;
                    fi[j] = zw.y & (kPackedFieldLimit - 1);
                    lastf = fi[j];
                }
            // run starts: a hit whose function index differs from the hit before it (or that has none before it in this step)
            const uint32_t below = anyb & lt;
            uint32_t pf = __shfl_sync(full, lastf, below ? 31 - __clz(below) : 0);
            bool hp = below != 0u;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (hm & (1u << j)) {
                    if (!hp || fi[j] != pf) rs |= 1u << j;
                    pf = fi[j];
                    hp = true;
                }
            n_runs = __reduce_add_sync(full, __popc(rs));
            // payload in position order
            const uint32_t cnt = __popc(hm);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(full, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            n_step = __shfl_sync(full, incl, 31);
            uint32_t o = incl - cnt;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (hm & (1u << j)) {
                    W[o] = __uint_as_float(wz[j]);
                    FI[o] = fi[j];
                    o++;
                }
            if (lane < 4u) W[n_step + lane] = 0.0f;  // x + (+0) == x: the scan lane adds the weights four at a time
        }
        const uint32_t sh4 = 4u * (lane & 7u), gm = 0xFFu << (lane & 24u);
        const uint32_t hw = __reduce_or_sync(gm, hm << sh4), sw = __reduce_or_sync(gm, rs << sh4);
        if ((lane & 7u) == 0u) {
            r->hits[lane >> 3] = hw;
            r->starts[lane >> 3] = sw;
        }
        if (lane == 0) {
            *reinterpret_cast<uint4 *>(&r->t0) = make_uint4(t0, flags | (n_runs == 1u ? kPcOneRun : 0u), index, n_step);
            r->seq_base = seq_base;
        }
        __syncwarp();
        pubc++;
        if (lane == 0) {
            __threadfence_block();
            sy->pub = pubc;
        }
        return n_step;
    };
    auto claim = [&]() -> uint32_t {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(work, (unsigned long long)kPcChunk);
        return (uint32_t)(c > 0xFFFFFFFFull ? 0xFFFFFFFFull : c);
    };

    uint32_t nxt = claim();
    for (;;) {
        const uint32_t c0 = __shfl_sync(full, nxt, 0);
        if (c0 >= n) break;
        nxt = claim();  // one chunk ahead: the atomic's round trip hides behind this chunk
        const uint32_t c1 = min(n, c0 + kPcChunk);
        for (uint32_t i = c0; i < c1; i++) {
            const uint64_t seq_base = __ldg(offsets + i);
            const uint32_t len = (uint32_t)(__ldg(offsets + i + 1) - seq_base);
            uint32_t count = 0;
            // this kernel runs the scan that assumes no protein can fill the 39 998-hit window: a caller of
            // ckm_call_batch_device that understated max_len is told so instead of getting indices that silently diverge
            if (len > kHitCap + CKM_KMER_SIZE && lane == 0) atomicExch(totals + 6, 1ull);
            if (len <= CKM_KMER_SIZE) {
                publish(0u, 0u, kPcEnd, i, seq_base);
            } else {
                uint32_t nwin = len - CKM_KMER_SIZE;  // the last window is never probed (kguts.cc:792, 798)
                const uint32_t nseg = (nwin + kHintSeg - 1) >> kHintShift;
                const uint32_t *hp = hints ? hints + hint_region(seq_base, index_base + i) : nullptr;
                const uint8_t *p0 = residues + seq_base;
                const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
                const uint32_t *wb = reinterpret_cast<const uint32_t *>(p0 - s);
                const uint32_t nwords = (len + s + 3u) >> 2;
                const uint32_t sh = 8u * s;
                uint32_t hv = kNoHint;     // hint of segment (32 * round + lane), reloaded every 32 segments
                uint32_t carry = kNoHint;  // last hint seen in front of the current segment (the first one there is, to begin with)
                uint32_t rw, rx;           // residue words of the step, fetched one step ahead
                tile_words(wb, nwords, 0, lane, rw, rx);

                for (uint32_t t0 = 0; t0 < nwin; t0 += kTile) {
                    uint32_t mh = kNoHint;
                    if (hp) {
                        const uint32_t seg0 = t0 >> kHintShift;  // first of the kSegsPerTile segments of this step
                        if ((seg0 & 31u) == 0u) {
                            hv = (seg0 + lane < nseg) ? __ldg(hp + seg0 + lane) : kNoHint;
                            if (carry == kNoHint) {
                                const uint32_t m = __ballot_sync(full, hv != kNoHint);
                                if (m) carry = __shfl_sync(full, hv, __ffs(m) - 1);
                            }
                        }
                        // this lane's hint: its segment's, else the nearest one in front, else the first one of the protein
#pragma unroll
                        for (int k = 0; k < kSegsPerTile; k++) {
                            uint32_t fk = __shfl_sync(full, hv, (seg0 & 31u) + k);
                            if (fk == kNoHint) fk = carry;
                            else carry = fk;
                            if ((int)(lane * kSegsPerTile >> 5) == k) mh = fk;
                        }
                    }
                    // The part of the neighbour copy the hints predict for this step is copied into shared memory asynchronously,
                    // before the keys are even built (the addresses depend on the position only): (weight, function word) of the
                    // 128 indices straight into the stage -- a window that matches is then already where a hit has to be -- and
                    // the residue string behind each 64-window half.  No register holds any of it while it is in flight.
                    static_assert(kSegsPerTile == 2, "one hint per 64-window half of a step");
                    const uint32_t mh0 = __shfl_sync(full, mh, 0), mh1 = __shfl_sync(full, mh, 16);
                    const bool hinted = mh0 != kNoHint || mh1 != kNoHint;
                    if (hinted) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {  // payload of window e = 32 j + lane
                            const uint32_t e = 32u * j + lane, mhe = j < 2 ? mh0 : mh1, idx = mhe + t0 + e;
                            if (mhe != kNoHint && t0 + e < nwin && idx < tv.n_chain) cp_async_8(stage + stage8_at(e), tv.cpay + idx, pol_first);
                        }
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) {  // residues [mh + t0 + 64 hf, + 71) as aligned words; what is not there is poisoned
                            const uint32_t mhe = hf ? mh1 : mh0, wb = ((mhe + t0 + 64u * hf) & ~3u) + 4u * lane;
                            if (lane < kPcLandWords) {
                                if (mhe != kNoHint && wb < tv.n_chain) cp_async_4(land + hf * kPcLandWords + lane, tv.cres + wb);
                                else land[hf * kPcLandWords + lane] = 0x7F7F7F7Fu;
                            }
                        }
                        cp_async_commit();
                        if (t0 + kTile < nwin) {  // the next step's part of the copy into L2
                            const uint32_t nb = mh1 + t0 + kTile;
                            if (mh1 != kNoHint && nb < tv.n_chain) {
                                if (lane < 8u) prefetch_l2(tv.cpay + nb + 16u * lane);
                                else if (lane == 8u) prefetch_l2(tv.cres + nb);
                            }
                        }
                    }

                    const TileKeys tk = tile_keys_from(lut, rw, rx, sh, t0, lane, len, nwin);
                    const bool last = !(t0 + kTile < nwin);
                    if (!last) tile_words(wb, nwords, t0 + kTile, lane, rw, rx);
                    const uint32_t act = tk.act;
                    // pull into L2 what the next step (or, from the first step, the next protein of this chunk) starts with
                    if (lane < 2u && t0 + 2u * kTile + 128u * lane < len + (i + 1 < c1 ? 256u : 0u)) prefetch_l2(p0 + t0 + 2u * kTile + 128u * lane);
                    if (t0 == 0 && hints && i + 1 < c1 && lane == 2u) prefetch_l2(hints + hint_region(seq_base + len, index_base + i + 1));
                    uint32_t h[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) h[j] = m35 ? fast_mod35(tk.key[j], nsig, m35) : (uint32_t)fast_mod(tk.key[j], tv.num_sigs, tv.magic);

                    // ---- occupancy words (L2) ----
                    uint32_t bw[4] = {0u, 0u, 0u, 0u};
                    if (tv.occupied) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (act & (1u << j)) bw[j] = __ldg(tv.occupied + (h[j] >> 5));
                    }
                    uint32_t hm = 0;
                    if (hinted) {
                        cp_async_wait_all();
                        __syncwarp();
                        // the residues of the copy behind this lane's four windows, against the query's: window j matches when a
                        // k-mer starts at its index (bit 7) and the eight residues from there on are equal
                        const uint32_t hf = lane >> 4, l16 = lane & 15u, mhe = hf ? mh1 : mh0;
                        const uint32_t *lw = land + hf * kPcLandWords + l16;
                        const uint32_t s8 = 8u * ((mhe + t0) & 3u);
                        const uint32_t w0 = lw[0], w1 = lw[1], w2 = lw[2], w3 = lw[3];
                        const uint32_t ra = __funnelshift_r(w0, w1, s8), rb = __funnelshift_r(w1, w2, s8), rc = __funnelshift_r(w2, w3, s8);
                        const uint32_t e0 = __vcmpeq4(ra & 0x7F7F7F7Fu, tk.c0), e1 = __vcmpeq4(rb & 0x7F7F7F7Fu, tk.c1),
                                       e2 = __vcmpeq4(rc & 0x7F7F7F7Fu, tk.c2);
                        const uint32_t eq = (e0 & 1u) | ((e0 >> 7) & 2u) | ((e0 >> 14) & 4u) | ((e0 >> 21) & 8u) | ((e1 << 4) & 16u) |
                                            ((e1 >> 3) & 32u) | ((e1 >> 10) & 64u) | ((e1 >> 17) & 128u) | ((e2 << 8) & 256u) |
                                            ((e2 << 1) & 512u) | ((e2 >> 6) & 1024u);
                        const uint32_t sig = ((ra >> 7) & 1u) | ((ra >> 14) & 2u) | ((ra >> 21) & 4u) | ((ra >> 28) & 8u);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (((eq >> j) & 0xFFu) == 0xFFu) hm |= 1u << j;
                        hm &= sig & act;
                    }
                    my_chain += __popc(hm);
                    uint32_t need = act & ~hm;
                    if (tv.occupied) {  // a window whose home slot is empty is a miss
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if ((need & (1u << j)) && !((bw[j] >> (h[j] & 31u)) & 1u)) need &= ~(1u << j);
                    }

                    // ---- what is left: lookup_hash_entry (kguts.cc:585-602).  The windows left are few (a dozen per step when the
                    //      neighbour copy is in use) and scattered over the lanes, so they are queued in shared memory and probed
                    //      one per lane; a queue entry is the key and the window, the home bucket is recomputed and its occupancy
                    //      word re-read (an L1 hit) by the probing lane ----
                    if (!no_left && __any_sync(full, need != 0u)) {
                        uint32_t n_left = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t b = __ballot_sync(full, (need >> j) & 1u);
                            if (need & (1u << j))
                                queue[n_left + __popc(b & lt)] = make_uint2((uint32_t)tk.key[j], (uint32_t)(tk.key[j] >> 32) | ((4u * lane + j) << 8));
                            n_left += __popc(b);
                        }
                        if (lane < 4u) sy->found[lane] = 0u;
                        __syncwarp();
                        for (uint32_t k = lane; k < n_left; k += 32u) {
                            const uint2 it = queue[k];
                            const uint64_t key = (uint64_t)it.x | ((uint64_t)(it.y & 0xFFu) << 32);
                            const uint32_t h0 = m35 ? fast_mod35(key, nsig, m35) : (uint32_t)fast_mod(key, tv.num_sigs, tv.magic);
                            uint32_t hh = h0;
                            uint4 v = ldg_v4_hint(slots + hh, pol_first);
                            const uint32_t ow0 = tv.occupied ? __ldg(tv.occupied + (hh >> 5)) : 0u;
                            // a window gets here because its home slot is taken -- nearly always by another k-mer, so the probe
                            // sequence goes on: when the same occupancy word says the next slot is taken too, fetch it now
                            if ((hh & 31u) != 31u && hh + 1u < nsig && ((ow0 >> ((hh & 31u) + 1u)) & 1u)) {
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
                                if (hh == h0) break;  // a table without an empty slot
                                if (tv.occupied) {
                                    const uint32_t ow = (hh >> 5) == (h0 >> 5) ? ow0 : __ldg(tv.occupied + (hh >> 5));
                                    if (!((ow >> (hh & 31u)) & 1u)) break;
                                }
                                v = __ldg(slots + hh);
                            }
                            if (found) {
                                const uint32_t e = it.y >> 8;
                                stage[stage8_at(e)] = make_uint2(v.z, v.w);
                                atomicOr(&sy->found[e >> 5], 1u << (e & 31u));
                            }
                        }
                        __syncwarp();
                        hm |= (sy->found[lane >> 3] >> (4u * (lane & 7u))) & need;
                    }
                    my_probes += __popc(act);
                    if (no_pub) count += __reduce_add_sync(full, __popc(hm));
                    else count += publish(hm, t0, last ? kPcEnd : 0u, i, seq_base);
                }
            }
            if (lane == 0) n_hits[i] = count;
            if (lane == 0) my_hits += count;
        }
    }
    publish(0u, 0u, kPcFin, 0u, 0ull);

    // batch totals: one atomic per warp (totals[4] = hits answered from the neighbour copy)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_probes += __shfl_down_sync(full, my_probes, d);
        my_chain += __shfl_down_sync(full, my_chain, d);
    }
    if (lane == 0) {
        atomicAdd(totals + 0, (unsigned long long)my_probes);
        atomicAdd(totals + 1, (unsigned long long)my_hits);
        atomicAdd(totals + 4, (unsigned long long)my_chain);
    }
}

}  // namespace ckm
