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

constexpr uint32_t kPcEnd = 1u;    // the protein ends with this record
constexpr uint32_t kPcFin = 2u;    // the producer has no more work
constexpr uint32_t kPcEmit = 4u;   // after the record's weights: the run is over, a call is due if its weighted sum suffices
constexpr uint32_t kPcReset = 8u;  // after the record's weights: the run is over without a call
constexpr int kPcProducers = 31;    // probing warps per block: 31 + 1 scan warp = 1 024 threads x 64 registers fill an SM
constexpr uint32_t kPcChunk = 4u;   // sequences claimed per atomic
constexpr uint32_t kPcDepth = 2u;   // published records a producer may be ahead of its scan lane
constexpr uint32_t kPcSpinLimit = 1u << 21;  // polls (a few tenths of a second) before a hand-off is declared stuck

// One published record: "add these n_w weights to the run's sum, in this order; then, if the flags say so, the run is over".
// The weights are in the W array of the same slot (record number % kPcDepth).  The call fields describe the run that ends
// (kPcEmit): everything of a call but its weighted sum, which only the scan lane knows.
struct __align__(16) PcRecord {
    uint32_t n_w, flags, index, count;
    uint32_t start, end, fI, pad_;
    uint64_t seq_base, pad2_;
};
struct __align__(16) PcSync {
    volatile uint32_t pub;    // records published (written by the producer)
    volatile uint32_t done;   // records consumed (written by the scan lane)
    volatile uint32_t abort;  // entry 0 only: a hand-off timed out somewhere in the block, nobody waits any more
    uint32_t pad_;
    uint32_t found[4];        // producer-private: left-over windows of the step whose hash probe hit
};
// The run state of KmerGuts::gather_hits that decides control flow (kguts.cc:816-856), carried by the probing warp from step to
// step (warp-uniform; parked in shared memory between steps): everything but the weighted sum.
struct __align__(16) PcRun {
    uint32_t cur, num, cnt, first;            // current_fI, stored hits, matching hits, position of the run's first hit
    uint32_t lastm, p1_pos, p1_fI, p1_wt;     // last matching position; the newest stored hit
};
static_assert(sizeof(PcRecord) == 48 && sizeof(PcSync) == 32 && sizeof(PcRun) == 32, "shared-memory layout");

constexpr uint32_t kPcWSlot = kTile + 8;  // weights of a record: a step's hits, one carried over from the step before, zero padding
constexpr uint32_t kPcLandWords = 20;  // residue bytes of the neighbour copy behind one 64-window half of a step: 64 + 7 (+3 of alignment)
// Everything a probing warp and its scan lane share, in one block of shared memory: a single base register addresses all of
// it (the kernel runs at the 64-register limit of a 1 024-thread block).  3.4 KB per producer -- what shared memory takes, the L1
// loses, and this kernel lives on its L1 (with the carve-out at its maximum probe_hint_kernel itself takes 7.1 ms per C2 step
// instead of 4.0; profiles/r2/k1_experiments.md).
struct __align__(16) PcWarp {
    PcSync sync;
    PcRun run;
    PcRecord rec[kPcDepth];
    uint2 queue[kTile];  // left-over windows of the step: (key low word, key high bits | window << 8)
    uint2 stage[kTile];  // (weight bits, word 3 of the packed slot) behind each window of the step that hit
    float W[kPcDepth][kPcWSlot];
    uint32_t land[2][kPcLandWords];  // residue string of the neighbour copy behind each half of the step
};
constexpr size_t pc_smem_bytes(int P) { return 256 + (size_t)P * sizeof(PcWarp); }

// floor(off / max(1, min_hits)) by a multiply: exact for off < 2^64 / min_hits (call_region_base, ckm_scan.cuh)
static inline uint64_t call_region_magic(int min_hits) { return min_hits > 1 ? ~0ull / (uint64_t)min_hits + 1ull : 0ull; }
__device__ __forceinline__ uint64_t call_region_base_fast(uint64_t off, uint32_t i, uint64_t magic) {
    return (magic ? __umul64hi(off, magic) : off) + i;
}

// stage[] index of a step's window e: 8-byte entries, read four consecutive ones per lane (publish) and written 32 consecutive
// ones per instruction (the asynchronous copy); XOR-ing the low four bits with the next three keeps both conflict-free
__device__ __forceinline__ uint32_t stage8_at(uint32_t e) { return e ^ ((e >> 4) & 15u); }

// The scan warp.  Lane w < P serves producer w.  The probing warp has already run the control part of the state machine
// (which hits count for the current run, where runs end: pc_scan_step below); what is left for a thread is what only a thread
// can do bit-exactly -- the f32 sum of a run's matching weights in hit order (kguts.cc:744-751) -- and, where a run ends, the
// test of that sum against min_weighted_hits, the call record, and at the end of the protein find_best_call of a single
// call (proteins with several calls, a tenth of C2's, are left to best_fixup_kernel so that their long, divergent walk does
// not hold up the other lanes).  One record per iteration and lane; the lanes run in lockstep, so the iteration is as long
// as the longest record -- at most a step's hits -- whatever the mix of functions among the hits.
// first / count: the producers this scan warp serves
__device__ __forceinline__ void pc_consume(PcWarp *warps, uint32_t lane, uint32_t first, uint32_t count, uint32_t index_base,
                                           const FusedArgs &fa, unsigned long long *totals) {
    constexpr uint32_t full = 0xffffffffu;
    const bool mine = lane < count;
    PcWarp *me = warps + first + (mine ? lane : 0u);
    PcSync *sy = &me->sync;
    PcSync *sy0 = &warps[0].sync;
    const float min_weighted = (float)fa.prm.min_weighted_hits;
    bool fin = !mine, at_start = true;
    uint32_t seen = 0, my_calls = 0, idle = 0, nc = 0;
    float ws = 0.0f;
    ckm_call_t *calls = nullptr;
    uint32_t c_fI = 0;  // the last call emitted: find_best_call of a single call needs no memory
    int32_t c_count = 0;
    float c_weighted = 0.0f;
    while (!__all_sync(full, fin)) {
        const bool have = !fin && sy->pub != seen;
        if (!__any_sync(full, have)) {  // nothing published anywhere: leave the issue slots to the producers
            if (++idle > kPcSpinLimit || sy0->abort) {
                if (lane == 0) atomicExch(totals + 7, 1ull);
                sy0->abort = 1u;
                break;
            }
            __nanosleep(100);
            continue;
        }
        idle = 0;
        if (have) {
            __threadfence_block();
            const uint32_t slot = seen & (kPcDepth - 1u);
            const PcRecord *r = &me->rec[slot];
            const uint4 q = *reinterpret_cast<const uint4 *>(&r->n_w);  // n_w, flags, index, count
            if (q.y & kPcFin) {
                fin = true;
            } else {
                if (at_start) {
                    calls = fa.calls + call_region_base_fast(r->seq_base, index_base + q.z, fa.call_magic);
                    at_start = false;
                }
                const float *W = me->W[slot];
                float s = ws;
                for (uint32_t x = 0; x < q.x; x += 4u) {  // padded with +0 up to a multiple of four: x + (+0) == x
                    const float4 v = *reinterpret_cast<const float4 *>(W + x);
                    s += v.x;
                    s += v.y;
                    s += v.z;
                    s += v.w;
                }
                ws = s;
                if (q.y & (kPcEmit | kPcReset)) {  // process_set_of_hits, kguts.cc:753-759 (the count test was made by the producer)
                    if ((q.y & kPcEmit) && ws >= min_weighted) {
                        const uint4 cf = *reinterpret_cast<const uint4 *>(&r->start);  // start, end, fI
                        ckm_call_t c;
                        c.start = cf.x;
                        c.end = cf.y;
                        c.count = (int32_t)q.w;
                        c.function_index = cf.z;
                        c.weighted_hits = ws;
                        calls[nc] = c;
                        c_fI = cf.z;
                        c_count = (int32_t)q.w;
                        c_weighted = ws;
                        nc++;
                    }
                    ws = 0.0f;
                }
                if (q.y & kPcEnd) {
                    fa.n_calls[q.z] = nc;
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
                            const FScore fs = {(int)c_fI, c_count, c_weighted};
                            top.push(fs);
                            best_from_top2(top, b);
                        }
                        fa.best[q.z] = b;
                    }
                    my_calls += nc;
                    nc = 0;
                    ws = 0.0f;
                    at_start = true;
                }
            }
            __threadfence_block();  // the payload reads above come before the release of the slot
            sy->done = ++seen;
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
    PcWarp *warps = reinterpret_cast<PcWarp *>(pc_smem + 256);
    fill_aa_lut(lut);
    if (threadIdx.x < P) {
        warps[threadIdx.x].sync.pub = 0;
        warps[threadIdx.x].sync.done = 0;
        warps[threadIdx.x].sync.abort = 0;
        warps[threadIdx.x].run.num = 0;
        warps[threadIdx.x].run.cnt = 0;
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
        pc_consume(warps, lane, first, first < (uint32_t)P ? min(per, (uint32_t)P - first) : 0u, index_base, fa, totals);
        return;
    }

    PcWarp *const me = warps + warp;
    PcSync *const sy = &me->sync;
    PcSync *const sy0 = &warps[0].sync;
    uint2 *const queue = me->queue;
    uint2 *const stage = me->stage;
    PcRun *const run = &me->run;
    uint32_t *const land = &me->land[0][0];
    uint32_t pubc = 0;  // records this warp has published
    const uint32_t lt = (1u << lane) - 1u;
    const uint4 *__restrict__ slots = reinterpret_cast<const uint4 *>(tv.slots);
    const uint32_t nsig = (uint32_t)tv.num_sigs;
    uint32_t my_probes = 0, my_hits = 0, my_chain = 0;
    const uint64_t pol_first = policy_evict_first();

    const uint32_t max_gap = (uint32_t)fa.prm.max_gap;
    const int min_hits = fa.prm.min_hits;

    // waits until the scan lane has released the slot record number pubc goes into
    auto wait_slot = [&]() {
        if (pubc >= kPcDepth && !no_scan) {  // record pubc - kPcDepth consumed?
            uint32_t spins = 0;
            while ((int32_t)(sy->done - (pubc - kPcDepth + 1u)) < 0) {
                if (++spins > kPcSpinLimit || sy0->abort) {
                    if (lane == 0) atomicExch(totals + 7, 1ull);
                    sy0->abort = 1u;
                    break;
                }
                __nanosleep(200);
            }
            __threadfence_block();
        }
    };
    // the entry of `stage` behind window e of the step: (weight bits, function index)
    auto staged = [&](uint32_t e) -> uint2 {
        const uint2 z = stage[stage8_at(e)];
        return make_uint2(z.x, z.y & (kPackedFieldLimit - 1));
    };

    // One step's hits through the CONTROL part of gather_hits (kguts.cc:816-856, 873-876) -- in parallel.  hm: this lane's
    // windows that hit (their (weight, function word) are in `stage`).  The state machine's decisions depend on the hits'
    // function indices and positions only, and they have a closed form.  With "pair" = two consecutive hits of one run with
    // the same function:
    //   * a run begins at the first hit ever and at every hit that follows a gap of more than max_gap (821-836);
    //   * current_fI after a hit = the function of the latest run-begin hit or pair up to it (a pair of another function than
    //     current_fI is what ends a run, 852-856, and the carry-over, 772-777, makes the pair's function the current one);
    //   * a hit counts for its run when its function is current_fI, and the first hit of a run-ending pair counts for the NEW
    //     run (carried over), which it begins.
    // So the warp marks run beginnings and matching hits with two short per-lane walks joined by ballots and shuffles, cuts
    // the step at the run beginnings (none in nearly every step) and publishes, per piece, the matching weights in order --
    // the one thing that must be done serially, by the scan lane: "add these n_w weights to the run's sum; then, if the flags
    // say so, the run is over".  Hits of other functions scattered through a run cost nothing extra.  In nearly every step all
    // hits are of the run's own function (`simple`) and none of the walks is made.  Returns the step's hit count.
    auto scan_step = [&](uint32_t hm, uint32_t t0, bool last, uint32_t index, uint64_t seq_base) -> uint32_t {
        const uint32_t anyb = __ballot_sync(full, hm != 0u);
        if (!anyb && !last) return 0u;  // nothing the scan lane needs to hear of
        const uint4 sa = *reinterpret_cast<const uint4 *>(&run->cur);
        uint32_t cur = sa.x, num = sa.y, cnt = sa.z, first = sa.w, lastm = run->lastm;
        uint32_t n_step = 0, addm = 0, startm = 0, beginm = 0, lo = 0xFu;  // lo: this lane's windows not yet accounted for
        uint32_t e_last = 0;  // the step's last hit (window index)
        bool pre = false;     // the step's first hit completes a run-ending pair with the last hit of the step before
        bool simple = false;
        uint32_t wz[4] = {0u, 0u, 0u, 0u};  // weights of this lane's four windows (whatever is there for windows that did not hit)
        if (anyb) {
            uint32_t fi[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint2 z = staged(4u * lane + j);
                wz[j] = z.x;
                fi[j] = z.y;
            }
            const uint32_t fl = __ffs(anyb) - 1u, ll = 31u - __clz(anyb);  // first / last lane with a hit
            n_step = __reduce_add_sync(full, __popc(hm));
            const uint32_t e0 = 4u * fl + (__ffs(__shfl_sync(full, hm, fl)) - 1u);
            e_last = 4u * ll + (31u - __clz(__shfl_sync(full, hm, ll)));
            // Nearly every step: all hits of one function, which is the run's (or there is no run yet), and no gap in front
            const uint32_t F0 = staged(e0).y;
            const uint32_t neq = (fi[0] != F0 ? 1u : 0u) | (fi[1] != F0 ? 2u : 0u) | (fi[2] != F0 ? 4u : 0u) | (fi[3] != F0 ? 8u : 0u);
            simple = fa.prm.max_gap >= (int)kTile - 1 && !__any_sync(full, (neq & hm) != 0u) &&  // (a negative max_gap wraps: every hit a gap)
                     (num == 0u || (F0 == cur && !((uint32_t)(run->p1_pos + max_gap) < t0 + e0)));
            if (simple) {
                addm = hm;
                if (num == 0u) {  // 833-836
                    cur = F0;
                    first = t0 + e0;
                }
                num += n_step;
                cnt += n_step;
                lastm = t0 + e_last;
            } else {
                const uint32_t wbase = t0 + 4u * lane;
                const uint32_t my_last_pos = wbase + (hm ? 31u - __clz(hm) : 0u);
                uint32_t lastf = 0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (hm & (1u << j)) lastf = fi[j];
                // the hit before this lane's first one: the last hit of the nearest lane below that has one, else the one carried in
                const uint32_t below = anyb & lt;
                const uint32_t src = below ? 31u - __clz(below) : 0u;
                uint32_t qF = __shfl_sync(full, lastf, src), qP = __shfl_sync(full, my_last_pos, src);
                bool qh = true;
                if (!below) {
                    qF = run->p1_fI;
                    qP = run->p1_pos;
                    qh = num > 0u;
                }
                // walk 1: run-begin hits, pairs, and the function of this lane's last "setter" (run-begin hit or pair)
                uint32_t pairm = 0, setF = 0;
                bool has_set = false;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (hm & (1u << j)) {
                        const uint32_t pos = wbase + j;
                        const bool beg = !qh || (uint32_t)(qP + max_gap) < pos;  // 821: unsigned arithmetic, as there
                        const bool pair = !beg && fi[j] == qF;
                        beginm |= beg ? (1u << j) : 0u;
                        pairm |= pair ? (1u << j) : 0u;
                        if (beg || pair) {
                            setF = fi[j];
                            has_set = true;
                        }
                        qF = fi[j];
                        qP = pos;
                        qh = true;
                    }
                // current_fI on entry to this lane
                const uint32_t setb = __ballot_sync(full, has_set);
                const uint32_t sbelow = setb & lt;
                uint32_t c = __shfl_sync(full, setF, sbelow ? 31u - __clz(sbelow) : 0u);
                if (!sbelow) c = cur;
                // walk 2: pairs that end a run, hits that match
                uint32_t trigm = 0, matchm = 0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (hm & (1u << j)) {
                        if (beginm & (1u << j)) {
                            c = fi[j];
                        } else if (pairm & (1u << j)) {
                            trigm |= fi[j] != c ? (1u << j) : 0u;
                            c = fi[j];
                        }
                        matchm |= fi[j] == c ? (1u << j) : 0u;
                    }
                // the hit before a run-ending pair's second hit begins the new run (and counts for it)
                uint32_t predm = 0;
#pragma unroll
                for (int j = 1; j < 4; j++)
                    if (trigm & (1u << j)) {
                        const uint32_t lower = hm & ((1u << j) - 1u);
                        if (lower) predm |= 1u << (31 - __clz(lower));
                    }
                const uint32_t tb = __ballot_sync(full, (trigm & hm & (0u - hm)) != 0u);  // lanes whose FIRST hit ends a run
                const uint32_t above = anyb & ~lt & ~(1u << lane);
                if (hm && above && ((tb >> (__ffs(above) - 1)) & 1u)) predm |= 1u << (31 - __clz(hm));
                addm = matchm | predm;
                startm = beginm | predm;
                pre = (tb >> fl) & 1u;
            }
        }
        // The step in pieces, cut where runs begin: [the carried pair's run end] [run beginnings inside the step]* [the rest].
        // One record at most per piece, all from this one place (the kernel has to stay inside the instruction cache).
        bool prepend = false;
        for (;;) {
            uint32_t seg = lo, flags = 0, sl = 0, sj = 0, eb = 0, rel = 0, kind = 2u;
            bool is_begin = false;
            if (pre) {
                kind = 0u;
                seg = 0u;
            } else if (const uint32_t sbal = simple ? 0u : __ballot_sync(full, (startm & lo) != 0u)) {
                kind = 1u;
                sl = __ffs(sbal) - 1u;
                sj = __ffs(__shfl_sync(full, startm & lo, sl)) - 1u;
                eb = 4u * sl + sj;  // the window the new run begins with
                rel = eb > 4u * lane ? min(eb - 4u * lane, 4u) : 0u;
                seg = lo & ((1u << rel) - 1u);
                is_begin = (__shfl_sync(full, beginm, sl) >> sj) & 1u;
            }
            const uint32_t a4 = addm & seg;
            const uint32_t ab = __ballot_sync(full, a4 != 0u);
            if (!simple) {  // account for the windows in `seg`: stored hits, matching hits, last matching position
                const uint32_t both = __reduce_add_sync(full, (__popc(a4) << 16) | __popc(hm & seg));
                num += both & 0xFFFFu;
                cnt += both >> 16;
                if (ab) {
                    const uint32_t hl = 31u - __clz(ab);
                    lastm = t0 + 4u * hl + (31u - __clz(__shfl_sync(full, a4, hl)));
                }
            }
            // the run that ends here: a run-ending pair flushes unconditionally (852-856), a gap only a run of min_hits stored
            // hits (821-831), the end of the protein likewise (873-876); a call needs min_hits matching hits (753) and, by
            // the scan lane's judgement, its weighted sum
            if (kind == 2u) {
                if (last) flags = kPcEnd | ((int)num < min_hits ? 0u : (int)cnt >= min_hits ? kPcEmit : kPcReset);
            } else if (kind == 0u || !is_begin) {
                flags = (int)cnt >= min_hits ? kPcEmit : kPcReset;
            } else if (num > 0u) {
                flags = ((int)num >= min_hits && (int)cnt >= min_hits) ? kPcEmit : kPcReset;
            }
            if (flags || prepend || ab) {
                // the record: this lane's weights of a4, compacted in position order (behind the carried hit's if `prepend`)
                wait_slot();
                const uint32_t slot = pubc & (kPcDepth - 1u);
                float *W = me->W[slot];
                uint32_t n_w = prepend ? 1u : 0u;
                if (ab) {
                    const uint32_t c4 = __popc(a4);
                    uint32_t incl = c4;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(full, incl, d);
                        if (lane >= (uint32_t)d) incl += t;
                    }
                    const uint32_t o = n_w + incl - c4;
                    n_w += __shfl_sync(full, incl, 31);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (a4 & (1u << j)) W[o + __popc(a4 & ((1u << j) - 1u))] = __uint_as_float(wz[j]);
                }
                if (lane < 4u) W[n_w + lane] = 0.0f;  // x + (+0) == x: the scan lane adds the weights four at a time
                if (lane == 0) {
                    if (prepend) W[0] = __uint_as_float(run->p1_wt);
                    PcRecord *r = &me->rec[slot];
                    *reinterpret_cast<uint4 *>(&r->n_w) = make_uint4(n_w, flags, index, cnt);
                    *reinterpret_cast<uint4 *>(&r->start) = make_uint4(first, lastm + (CKM_KMER_SIZE - 1), cur, 0u);
                    r->seq_base = seq_base;
                }
                __syncwarp();
                pubc++;
                if (lane == 0) {
                    __threadfence_block();
                    sy->pub = pubc;
                }
            }
            prepend = false;
            if (kind == 2u) break;
            if (kind == 0u) {  // the carried hit begins the new run; its weight goes in front of this step's
                cur = run->p1_fI;
                first = run->p1_pos;
                lastm = first;
                num = 1;
                cnt = 1;
                prepend = true;
                pre = false;
            } else {
                cur = staged(eb).y;
                first = t0 + eb;
                num = 0;
                cnt = 0;
                lo &= ~((1u << rel) - 1u);
                if (lane == sl) startm &= ~(1u << sj);
            }
        }
        if (last) {
            num = 0;
            cnt = 0;
        }
        if (lane == 0) {
            *reinterpret_cast<uint4 *>(&run->cur) = make_uint4(cur, num, cnt, first);
            run->lastm = lastm;
            if (anyb) {  // the newest stored hit
                const uint2 z = staged(e_last);
                run->p1_pos = t0 + e_last;
                run->p1_fI = z.y;
                run->p1_wt = z.x;
            }
        }
        __syncwarp();
        return n_step;
    };
    // a record without weights: an empty protein's end, or this warp's last word
    auto emit_bare = [&](uint32_t flags, uint32_t index, uint64_t seq_base) {
        wait_slot();
        PcRecord *r = &me->rec[pubc & (kPcDepth - 1u)];
        pubc++;
        if (lane == 0) {
            *reinterpret_cast<uint4 *>(&r->n_w) = make_uint4(0u, flags, index, 0u);
            r->seq_base = seq_base;
            __threadfence_block();
            sy->pub = pubc;
        }
        __syncwarp();
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
                emit_bare(kPcEnd, i, seq_base);
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
                    for (int j = 0; j < 4; j++) h[j] = (uint32_t)table_home(tv, tk.key[j]);

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
                            const uint32_t h0 = (uint32_t)table_home(tv, key);
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
                    else count += scan_step(hm, t0, last, i, seq_base);
                }
            }
            if (lane == 0) n_hits[i] = count;
            if (lane == 0) my_hits += count;
        }
    }
    emit_bare(kPcFin, 0u, 0ull);

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
