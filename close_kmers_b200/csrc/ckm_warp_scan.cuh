// The ordered scoring scan as a register-resident transducer, one hit at a time: S1/S2 (the run logic of
// KmerGuts::gather_hits, kguts.cc:816-856, 873-876, and process_set_of_hits, 734-781) followed by B1 (find_best_call,
// 1008-1199) at the end of the protein (ckm_pc.cuh).  It is the non-GENERAL case of scan_kernel (ckm_scan.cuh: no order_constraint, the
// protein cannot saturate the 39 998-hit window) without the OTU bookkeeping, written so that the scan lanes of
// probe_pc_kernel (ckm_pc.cuh) can run it on hits that never leave the SM.  All f32 sums are made in hit order
// (744-751), so calls and best-call records are bit-identical to scan_kernel's and to the reference's.
#pragma once
#include "ckm_scan.cuh"

namespace ckm {

struct FusedArgs {
    ckm_call_t *calls;       // call regions (call_region_base)
    ckm_call_t *calls_work;  // same geometry, scratch of find_best_call; null when best is
    uint32_t *n_calls;       // per sequence of this launch
    ckm_best_t *best;        // per sequence of this launch, or null
    uint64_t call_magic;     // call_region_magic(prm.min_hits)
    Params prm;
};

struct RunState {
    uint32_t num, cur_fI;
    uint32_t first_pos, last_match_pos;
    int32_t fI_count;
    float wsum;
    uint32_t p1_pos, p1_fI;  // newest stored hit
    float p1_wt;
    uint32_t p2_pos, p2_fI;  // the stored hit before it
    float p2_wt;
    uint32_t n_calls;
    uint32_t c_fI;  // the last call emitted: find_best_call of a single call needs no memory
    int32_t c_count;
    float c_weighted;
};

__device__ __forceinline__ void rs_begin(RunState &S) {
    S.num = 0;
    S.cur_fI = 0;
    S.first_pos = S.last_match_pos = 0;
    S.fI_count = 0;
    S.wsum = 0.0f;
    S.p1_pos = S.p1_fI = S.p2_pos = S.p2_fI = 0;
    S.p1_wt = S.p2_wt = 0.0f;
    S.n_calls = 0;
    S.c_fI = 0;
    S.c_count = 0;
    S.c_weighted = 0.0f;
}

__device__ __forceinline__ void rs_reset_run(RunState &S) {
    S.num = 0;
    S.fI_count = 0;
    S.wsum = 0.0f;
}

// process_set_of_hits, kguts.cc:734-781 (the incremental form, as scan_kernel's flush)
__device__ __forceinline__ void rs_flush(RunState &S, const Params &prm, ckm_call_t *calls) {
    if (S.fI_count >= prm.min_hits && S.wsum >= (float)prm.min_weighted_hits) {
        ckm_call_t c;
        c.start = S.first_pos;
        c.end = S.last_match_pos + (CKM_KMER_SIZE - 1);
        c.count = S.fI_count;
        c.function_index = S.cur_fI;
        c.weighted_hits = S.wsum;
        calls[S.n_calls] = c;
        S.c_fI = S.cur_fI;
        S.c_count = S.fI_count;
        S.c_weighted = S.wsum;
        S.n_calls++;
    }
    if (S.num >= 2 && S.p2_fI != S.cur_fI && S.p2_fI == S.p1_fI) {  // 772-777: the last two hits seed the next run
        S.cur_fI = S.p1_fI;
        S.num = 2;
        S.first_pos = S.p2_pos;
        S.fI_count = 2;
        float w = 0.0f;
        w += S.p2_wt;
        w += S.p1_wt;
        S.wsum = w;
        S.last_match_pos = S.p1_pos;
    } else {
        rs_reset_run(S);
    }
}

// one hit through the state machine (kguts.cc:816-856)
__device__ __forceinline__ void rs_hit(RunState &S, const Params &prm, uint32_t pos, uint32_t fI, float wt, ckm_call_t *calls) {
    if (S.num > 0 && (uint32_t)(S.p1_pos + (uint32_t)prm.max_gap) < pos) {  // gap rule, 821-831 (unsigned int arithmetic)
        if ((int)S.num >= prm.min_hits) rs_flush(S, prm, calls);
        else rs_reset_run(S);
    }
    if (S.num == 0) {  // 833-836
        S.cur_fI = fI;
        S.first_pos = pos;
    }
    S.num++;
    if (fI == S.cur_fI) {
        S.fI_count++;
        S.wsum += wt;
        S.last_match_pos = pos;
    }
    S.p2_pos = S.p1_pos;
    S.p2_fI = S.p1_fI;
    S.p2_wt = S.p1_wt;
    S.p1_pos = pos;
    S.p1_fI = fI;
    S.p1_wt = wt;
    if (S.num > 1 && S.cur_fI != fI && S.p2_fI == S.p1_fI) rs_flush(S, prm, calls);  // 852-856
}

}  // namespace ckm
