// K2: ordered scoring scan over each protein's compacted hit list, fused with find_best_call.
//
// Replaces the run logic of KmerGuts::gather_hits (kguts.cc:816-856, 873-876), process_set_of_hits
// (734-781) and find_best_call (1008-1199).  The scan is a serial transducer per protein (state: the
// stored-hit window, current_fI, the last two stored hits), so it runs one THREAD per protein -- 32
// proteins per warp with every lane busy -- over the position-ordered hit list K1 left in HBM.  All
// f32 sums are accumulated in the reference's order (hit order within a run, call order within
// find_best_call), so scores are bit-identical, not merely close.
//
// Run statistics are kept incrementally: current_fI is fixed for the lifetime of a stored-hit window
// (set when the window is empty, kguts.cc:833-836, or at carry-over, 772-777), so counting matching
// hits and adding their weights as they are stored performs exactly the additions that
// process_set_of_hits performs when it re-walks the window (744-751).
#pragma once
#include "ckm_common.cuh"

namespace ckm {

constexpr int kScanThreads = 128;
constexpr int kScanBatch = 8;  // hit records in flight per thread
constexpr uint32_t kHitCap = CKM_MAX_HITS_PER_SEQ - 2;  // kguts.cc:850

// Where protein i's calls live before compaction.  Every emitted call counts >= max(1,min_hits) hits
// and no hit is counted by two calls (carried-over hits did not match the run that was just flushed),
// so n_calls_i <= (len_i - 8) / mh; floor(off/mh) + i leaves at least floor(len_i/mh) + 1 slots.
__host__ __device__ __forceinline__ uint64_t call_region_base(uint64_t residue_offset, uint32_t i, int min_hits) {
    const uint64_t mh = min_hits > 1 ? (uint64_t)min_hits : 1ull;
    return residue_offset / mh + i;
}

struct FScore {  // FuncScore + key, kguts.cc:984-1000
    int fI;
    int count;
    float weighted;
};

// std::partial_sort(vec.begin(), vec.begin()+2, vec.end(), weighted-desc) as libstdc++ executes it
// (__heap_select + __sort_heap over a 2-element heap; bits/stl_algo.h, bits/stl_heap.h), consumed as a
// stream: vec is produced in ascending function index (std::map order) and only vec[0..2] are read
// afterwards (kguts.cc:1134-1196).  The heap top v0 is the SMALLER weight of the two kept.
struct Top2 {
    FScore v0, v1, v2;
    uint32_t n;
    __device__ __forceinline__ void adjust(const FScore &value) {  // __adjust_heap(first, 0, 2, value)
        v0 = v1;
        if (v0.weighted > value.weighted) {
            v1 = v0;
            v0 = value;
        } else {
            v1 = value;
        }
    }
    __device__ __forceinline__ void push(const FScore &e) {
        if (n == 0) {
            v0 = e;
        } else if (n == 1) {
            v1 = e;
            const FScore t = v0;
            adjust(t);  // __make_heap
        } else {
            if (n == 2) v2 = e;
            if (e.weighted > v0.weighted) {  // __pop_heap(first, middle, i)
                if (n == 2) v2 = v0;
                adjust(e);
            }
        }
        n++;
    }
    __device__ __forceinline__ void finish() {  // __sort_heap on two elements swaps them
        if (n > 1) {
            const FScore t = v0;
            v0 = v1;
            v1 = t;
        }
    }
};

__device__ __forceinline__ void best_from_top2(const Top2 &t, ckm_best_t &out) {
    float score_offset = t.n == 1 ? (float)t.v0.count : (float)(t.v0.count - t.v1.count);  // 1149-1152
    out.score_offset = score_offset;
    if (score_offset >= 5.0f) {
        out.function_index = t.v0.fI;
        out.score = (float)t.v0.count;
        out.weighted_score = t.v0.weighted;
    } else if (t.n >= 2) {
        bool ambig = false;
        if (t.n == 2) {
            ambig = true;
        } else {
            const float pair_offset = (float)(t.v1.count - t.v2.count);
            if (pair_offset > 5.0f) {
                ambig = true;
                out.score_offset = pair_offset;
                out.weighted_score = t.v0.weighted;
            }
        }
        if (ambig) {
            out.flags |= CKM_BEST_AMBIG;
            out.ambig_a = t.v0.fI;
            out.ambig_b = t.v1.fI;
            out.score = (float)t.v0.count;
        }
    }
}

// find_best_call over calls[0..n) (kguts.cc:1008-1199); `work` is n call slots of scratch
// (no __restrict__: the caller may have written calls[] itself just before, and a non-coherent load must not be used)
__device__ void find_best_call_dev(const ckm_call_t *calls, uint32_t n, ckm_call_t *work, ckm_best_t &out) {
    out.function_index = -1;
    out.ambig_a = out.ambig_b = -1;
    out.flags = 0;
    out.score = out.weighted_score = out.score_offset = 0.0f;
    if (n == 0) return;
    out.flags = CKM_BEST_HAS_CALLS;
    Top2 top;
    top.n = 0;
    if (n == 1) {  // by far the common case: one call, one function
        const ckm_call_t c = calls[0];
        FScore e = {(int)c.function_index, c.count, c.weighted_hits};
        top.push(e);
        best_from_top2(top, out);
        return;
    }
    // 1023-1040: collapse adjacent calls with the same function
    uint32_t nc = 0;
    for (uint32_t i = 0; i < n;) {
        ckm_call_t cur = calls[i++];
        while (i < n && cur.function_index == calls[i].function_index) {
            cur.end = calls[i].end;
            cur.count += calls[i].count;
            cur.weighted_hits += calls[i].weighted_hits;
            i++;
        }
        work[nc++] = cur;
    }
    // 1063-1086: F1 | F2 (count < 5) | F1 with combined count >= 10 -> one F1; in place (nm <= i)
    uint32_t nm = 0;
    for (uint32_t i = 0; i < nc;) {
        ckm_call_t cur = work[i++];
        while (i + 1 < nc && cur.function_index == work[i + 1].function_index && work[i].count < 5 &&
               cur.count + work[i + 1].count >= 10) {
            cur.end = work[i + 1].end;
            cur.count += work[i + 1].count;
            cur.weighted_hits += work[i + 1].weighted_hits;
            i += 2;
        }
        work[nm++] = cur;
    }
    // 1108-1128: per-function totals in ascending (int) function index; sums in merged-call order
    long long last = -(1ll << 40);
    for (;;) {
        long long next = 1ll << 40;
        for (uint32_t i = 0; i < nm; i++) {
            const long long f = (int)work[i].function_index;
            if (f > last && f < next) next = f;
        }
        if (next == (1ll << 40)) break;
        FScore e = {(int)next, 0, 0.0f};
        bool first = true;
        for (uint32_t i = 0; i < nm; i++) {
            if ((int)work[i].function_index == (int)next) {
                if (first) {
                    e.count = work[i].count;
                    e.weighted = work[i].weighted_hits;
                    first = false;
                } else {
                    e.count += work[i].count;
                    e.weighted += work[i].weighted_hits;
                }
            }
        }
        top.push(e);
        last = next;
    }
    top.finish();
    best_from_top2(top, out);
}

struct ScanArgs {
    const uint64_t *offsets;   // n+1 residue offsets (hit regions are indexed by them)
    const HitRec *hits;
    const uint16_t *hit_avg;   // non-null iff order_constraint
    const uint32_t *n_hits;
    uint32_t *stored_idx;      // GENERAL only: index (within the protein's hit list) of every stored hit
    ckm_call_t *calls;         // call regions (call_region_base)
    ckm_call_t *calls_work;    // same geometry, scratch for find_best_call
    uint32_t *n_calls;
    ckm_otu_t *otus;           // regions indexed by residue offset; null unless OTU stats wanted
    uint32_t *n_otus;
    ckm_best_t *best;          // null unless wanted
    unsigned long long *totals;  // [2] += calls
    uint32_t n;
    uint32_t index_base;       // batch index of sequence 0 of this launch (chunked launches share the call regions)
    Params prm;
};

// GENERAL = order_constraint != 0 or a protein long enough to saturate the 39998-hit window: then the
// stored hits are a strict subsequence of the hit list and their indices are kept in stored_idx.
template <bool GENERAL>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(ScanArgs a) {
    // A thread owns a protein and walks its hit list serially, so a warp lasts as long as its longest list.  The block's
    // proteins are therefore dealt to its threads in order of hit count (counting sort on n_hits / 4 in shared memory): the
    // 32 lanes of a warp then finish together.  Which thread takes which protein changes no result.
    __shared__ uint32_t s_bins[256];
    __shared__ uint16_t s_order[kScanThreads];
    {
        const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t nh0 = i0 < a.n ? a.n_hits[i0] : 0u;
        const uint32_t bin = min(nh0 >> 2, 255u);
        for (uint32_t t = threadIdx.x; t < 256u; t += blockDim.x) s_bins[t] = 0;
        __syncthreads();
        atomicAdd(&s_bins[bin], 1u);
        __syncthreads();
        if (threadIdx.x < 32u) {  // exclusive scan of the 256 bins by one warp, 8 bins per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                v[k] = s_bins[threadIdx.x * 8 + k];
                sum += v[k];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (threadIdx.x >= (uint32_t)d) incl += t;
            }
            uint32_t ex = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                s_bins[threadIdx.x * 8 + k] = ex;
                ex += v[k];
            }
        }
        __syncthreads();
        s_order[atomicAdd(&s_bins[bin], 1u)] = (uint16_t)threadIdx.x;
        __syncthreads();
    }
    const uint32_t i = blockIdx.x * blockDim.x + s_order[threadIdx.x];
    if (i >= a.n) return;
    const uint64_t base = a.offsets[i];
    const HitRec *H = a.hits + base;
    const uint16_t *A = a.hit_avg ? a.hit_avg + base : nullptr;
    uint32_t *S = GENERAL ? a.stored_idx + base : nullptr;
    const uint32_t nh = a.n_hits[i];
    if (!GENERAL && nh > kHitCap) atomicExch(a.totals + 6, 1ull);  // max_len was understated (ckm_call_batch_device): reported, not ignored
    ckm_call_t *calls = a.calls + call_region_base(base, a.index_base + i, a.prm.min_hits);
    ckm_otu_t *otus = a.otus ? a.otus + base : nullptr;
    const int min_hits = a.prm.min_hits;
    const float min_weighted = (float)a.prm.min_weighted_hits;
    const uint32_t max_gap = (uint32_t)a.prm.max_gap;

    uint32_t n_calls = 0, n_otus = 0;
    // stored-hit window = stored[bstart, bstart+num)
    uint32_t num = 0, bstart = 0, n_stored = 0;
    uint32_t cur_fI = 0, first_pos = 0, last_match_pos = 0;
    int fI_count = 0;
    float wsum = 0.0f;
    // last two stored hits: p1 = newest, p2 = the one before
    uint32_t p1_pos = 0, p1_fI = 0, p1_avg = 0, p2_pos = 0, p2_fI = 0;
    float p1_wt = 0.0f, p2_wt = 0.0f;

    auto flush = [&]() {  // process_set_of_hits, kguts.cc:734-781
        if (fI_count >= min_hits && wsum >= min_weighted) {
            ckm_call_t c;
            c.start = first_pos;
            c.end = last_match_pos + (CKM_KMER_SIZE - 1);
            c.count = fI_count;
            c.function_index = cur_fI;
            c.weighted_hits = wsum;
            calls[n_calls++] = c;
            if (otus) {  // 760-769: otu_map[oI]++ for every matching hit of the window
                for (uint32_t k = 0; k < num; k++) {
                    const uint32_t idx = GENERAL ? S[bstart + k] : bstart + k;
                    const HitRec h = H[idx];
                    if (h.fI != cur_fI) continue;
                    uint32_t e = 0;
                    while (e < n_otus && otus[e].otu_index != h.oI) e++;
                    if (e == n_otus) {
                        otus[e].otu_index = h.oI;
                        otus[e].count = 0;
                        n_otus++;
                    }
                    otus[e].count++;
                }
            }
        }
        // 772-780; with fewer than two stored hits the reference reads before hits[] (undefined
        // behaviour, reachable only with min_hits < 2): defined here as "no carry"
        if (num >= 2 && p2_fI != cur_fI && p2_fI == p1_fI) {
            cur_fI = p1_fI;
            bstart += num - 2;
            num = 2;
            first_pos = p2_pos;
            fI_count = 2;
            wsum = 0.0f;
            wsum += p2_wt;
            wsum += p1_wt;
            last_match_pos = p1_pos;
        } else {
            bstart += num;
            num = 0;
            fI_count = 0;
            wsum = 0.0f;
        }
    };

    // One hit of the ordered list through the state machine.  The hit records are fetched kScanBatch at a time as
    // independent 16-byte loads: the walk is otherwise one exposed HBM/L2 latency per hit (a thread owns a whole
    // protein), which is what bounded this kernel before (1 M proteins x ~190 hits: 0.98 ms, DRAM 3 TB/s idle).
    auto step = [&](const HitRec &h, uint32_t k, uint32_t avg) {
        // gap rule, 821-831 (unsigned int arithmetic)
        if (num > 0 && (uint32_t)(p1_pos + max_gap) < h.pos) {
            if ((int)num >= min_hits) {
                flush();
            } else {
                bstart += num;
                num = 0;
                fI_count = 0;
                wsum = 0.0f;
            }
        }
        if (num == 0) cur_fI = h.fI;  // 833-836
        bool ok = true;
        if (GENERAL && A) {
            if (a.prm.order_constraint && num != 0) {  // 838-842: unsigned difference, labs() of it <= 20
                const uint32_t d = (h.pos - p1_pos) - (uint32_t)((int)p1_avg - (int)avg);
                ok = (h.fI == p1_fI) && d <= 20u;
            }
        }
        if (ok) {
            if (num < kHitCap) {  // 850-851: beyond the cap the hit lands in hits[num] but is never counted
                if (GENERAL) S[n_stored] = k;
                n_stored++;
                if (num == 0) first_pos = h.pos;
                num++;
                if (h.fI == cur_fI) {
                    fI_count++;
                    wsum += h.wt;
                    last_match_pos = h.pos;
                }
                p2_pos = p1_pos; p2_fI = p1_fI; p2_wt = p1_wt;
                p1_pos = h.pos; p1_fI = h.fI; p1_wt = h.wt; p1_avg = avg;
            }
            // 852-856: two stored hits in a row of another function end the run
            if (num > 1 && cur_fI != h.fI && p2_fI == p1_fI) flush();
        }
    };
    for (uint32_t k0 = 0; k0 < nh; k0 += kScanBatch) {
        HitRec hb[kScanBatch];
        uint32_t ab[kScanBatch];
#pragma unroll
        for (int u = 0; u < kScanBatch; u++) {
            if (k0 + u < nh) {
                hb[u] = H[k0 + u];
                ab[u] = (GENERAL && A) ? (uint32_t)A[k0 + u] : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < kScanBatch; u++)
            if (k0 + u < nh) step(hb[u], k0 + u, ab[u]);
    }
    if ((int)num >= min_hits) flush();  // 873-876

    a.n_calls[i] = n_calls;
    if (a.n_otus) {
        // std::map iteration order: ascending otu_index (insertion sort, the list is tiny)
        for (uint32_t x = 1; x < n_otus; x++) {
            const ckm_otu_t v = otus[x];
            uint32_t y = x;
            while (y > 0 && otus[y - 1].otu_index > v.otu_index) {
                otus[y] = otus[y - 1];
                y--;
            }
            otus[y] = v;
        }
        a.n_otus[i] = n_otus;
    }
    if (a.best) {
        ckm_best_t b;
        find_best_call_dev(calls, n_calls, a.calls_work + call_region_base(base, a.index_base + i, a.prm.min_hits), b);
        a.best[i] = b;
    }
    {  // batch total: one atomic per (converged part of a) warp
        const unsigned m = __activemask();
        const uint32_t s = __reduce_add_sync(m, n_calls);
        if ((threadIdx.x & 31u) == (uint32_t)(__ffs(m) - 1) && s) atomicAdd(a.totals + 2, (unsigned long long)s);
    }
}

}  // namespace ckm
