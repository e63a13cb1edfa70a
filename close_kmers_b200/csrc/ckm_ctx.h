// Internal: the context behind ckm_ctx* (device table, parameters, grow-only work buffers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "ckm_common.cuh"

int ckm_fail(int code, const char *fmt, ...);

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);  // grow-only; contents are NOT preserved
    void release();
};
struct PinBuf {  // page-locked host memory
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

struct ckm_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: second lane of the pipelined host path
    cudaStream_t stream_h2d = nullptr;                 // the pipelined path's host-to-device copies: never queued behind a kernel
    static constexpr int kCopyEvents = 64;
    cudaEvent_t ev_copy[kCopyEvents] = {};             // chunk k's copy has landed (ring)
    cudaEvent_t ev_ready = nullptr, ev_done2 = nullptr;
    uint64_t pipeline_chunk_bytes = 48ull << 20, pipeline_min_bytes = 32ull << 20;  // 48 MB: profiles/r1/tune_e2e_chunk_sizes_r1g.jsonl
    uint32_t pipeline_ramp_div = 6, pipeline_tail_div = 4;    // first chunk = chunk/ramp_div (then doubling), last = chunk/tail_div (profiles/r2/tune_e2e_r2c.jsonl)
    bool force_raw = false;
    uint32_t tuning = 0;  // TableView::tuning
    uint32_t probe_group_override = 0;  // CKM_PROBE_GROUP (read once at ctx creation)
    bool staged_upload = true;          // CKM_NO_STAGED_UPLOAD=1 (read once) switches the threaded upload of pageable input off
    // automatic fall-back of K1 from the neighbour copy to plain hash probing (see ckm_api.cu: adapt_probe_path)
    bool last_fused = false;  // the last batch went through probe_pc_kernel
    uint64_t pc_seq = 0;  // launches of probe_pc_kernel so far (picks the work counter)
    bool copy_suspended = false, last_used_copy = false;
    uint32_t copy_retry_in = 0, copy_backoff = 16, copy_suspensions = 0;
    int l2_fetch = 0;  // cudaLimitMaxL2FetchGranularity in effect

    // signature table in HBM
    bool shares_tables = false;  // a ckm_clone: table / occupied / family tables belong to the parent
    DevBuf table, occupied;  // occupied: 1 bit per slot, only for tables larger than L2
    DevBuf cres, cpay;       // compact form of the copy: residue string + (weight, function word) per index (probe_pc_kernel)
    DevBuf chain, cpos;      // neighbour-ordered copy of the occupied slots + slot -> index in it (ckm_chain.cuh)
    uint32_t n_chain = 0;
    uint64_t n_chains = 0, chain_cycles = 0;
    double chain_build_ms = 0.0;
    int l2_bytes = 0;
    bool has_l2_window = false;
    cudaAccessPolicyWindow l2_window;  // persisting window over `occupied`
    uint64_t num_sigs = 0, magic = 0;  // buckets of the table in HBM (the image's, or the library's own power of two)
    uint64_t image_buckets = 0;        // num_sigs of the image header (what ckm_num_sigs reports)
    uint32_t hbits = 0;                // TableView::hbits
    bool reference_hash = false;       // CKM_REFERENCE_HASH=1 (read once): keep the image's slot order and key % num_sigs
    int slot_bytes = 0;

    std::vector<std::string> functions, otu_names;  // function.index / otu.index
    ckm::Params prm;
    uint64_t launches = 0;

    // optional per-kernel timing (ckm_profile_*): events bracketing K1 and K2 of every batch
    bool profiling = false;
    struct ProfEv { cudaEvent_t e0, e1, e2; bool has_scan; };
    std::vector<ProfEv> prof;

    // current batch
    uint32_t cur_n = 0, cur_flags = 0;
    uint64_t cur_total = 0;
    const uint64_t *cur_off = nullptr;

    // device work buffers
    DevBuf in_packed, in_woff;  // packed input of ckm_call_batch_packed (seven residues per word) and its word offsets
    DevBuf in_res, in_off, totals, hits, hit_keys, hit_avg, n_hits, stored_idx, calls, calls_work, n_calls;
    DevBuf work;   // work counters of probe_pc_kernel
    DevBuf hints;  // per 32-window segment: where the protein sits in chain[] (ckm_hint.cuh)
    DevBuf otus, n_otus, best, ps_blocks;
    DevBuf hit_off, call_off, otu_off, hits_out, calls_out, otus_out;
    // family voting (ckm_family.cuh)
    struct Family {
        bool loaded = false;
        DevBuf table, ids, fam_func, fam_pgf, func_sid;  // device images of the side tables
        uint64_t mask = 0;
        uint32_t n_fams = 0, n_functions = 0, hypo_sid = 0;
        std::vector<std::string> pgf_names, plf;         // host strings for the response text
        DevBuf hit_fam, E, gcap, gofs, gscratch, matches;  // per-batch work buffers
        DevBuf class_seen, overflow;                       // fam_vote_kernel: classes present; proteins SMALL hands to LARGE
        DevBuf sofs, snd, sentries, sout_off, sout;        // ckm_family_scores: per-protein (family, count, weight) lists
        PinBuf h_scores, h_score_off;
    } fam;
    PinBuf h_fam;

    // /add postings + /matrix (ckm_matrix.cuh)
    struct Post {
        DevBuf keys, eids;  // the appended (k-mer, peg id) pairs
        uint64_t n = 0, mask = 0, n_keys = 0;  // n_keys: distinct k-mers, known after indexing
        bool dirty = true;
        DevBuf tkeys, tcnt, tcur, toff, slots, ids, occ;                     // index built lazily
        DevBuf d_eids, d_first, rcap, rofs, nd, out_off, entries, out;      // per-request work buffers
        PinBuf h_out, h_off;
        void release_all() {
            DevBuf *b[] = {&keys, &eids, &tkeys, &tcnt, &tcur, &toff, &slots, &ids, &occ, &d_eids, &d_first, &rcap, &rofs, &nd, &out_off,
                           &entries, &out};
            for (auto x : b) x->release();
            h_out.release();
            h_off.release();
        }
    } post, famnr;  // famnr: (k-mer, family id) pairs being collected by ckm_family_nr_add
    // postings of the mappings that are not selected (ckm_postings_select): one KmerPegMapping per "/mapping/<key>"
    std::map<uint32_t, Post> post_store;
    uint32_t post_key = 0;
    bool matrix_tile_on_device = false;  // ckm_matrix_rows_device: the COO tile stays in HBM, rows ordered by partner id
    uint64_t matrix_walked = 0;
    uint64_t famnr_compact_at = 1ull << 28, famnr_last_unique = 0;  // dedupe the collected pairs once this many are held

    // fastq path (ckm_fq.cuh)
    struct Fq {
        DevBuf nfrag, naa, frag_base, res_base, frag_off, frag_res;
        DevBuf best_frame, best_score, best_n, best_first, match_off, matches;
        PinBuf h_frag_base, h_frag_off, h_frag_res, h_best_frame, h_best_score, h_match_off, h_matches;
    } fq;

    // staged upload of pageable caller memory (upload_batch): per worker thread two pinned buffers, a stream and two events
    static constexpr int kStageThreads = 4;
    static constexpr size_t kStageChunk = 8u << 20;
    PinBuf stage_buf[kStageThreads][2];
    cudaStream_t stage_stream[kStageThreads] = {};
    cudaEvent_t stage_ev[kStageThreads][2] = {};

    // pinned host buffers handed out through ckm_batch_out_t
    PinBuf h_off, h_totals, h_hit_off, h_hits, h_call_off, h_calls, h_otu_off, h_otus, h_best;

    void free_all() {
        if (shares_tables) {  // drop the borrowed handles before the common release below
            table = DevBuf();
            occupied = DevBuf();
            chain = DevBuf();
            cpos = DevBuf();
            cres = DevBuf();
            cpay = DevBuf();
            DevBuf *borrowed[] = {&fam.table, &fam.ids, &fam.fam_func, &fam.fam_pgf, &fam.func_sid};
            for (auto b : borrowed) *b = DevBuf();
        }
        DevBuf *d[] = {&table, &occupied, &chain, &cpos, &cres, &cpay, &hints, &work, &in_packed, &in_woff, &in_res, &in_off, &totals, &hits, &hit_keys, &hit_avg, &n_hits, &stored_idx, &calls, &calls_work,
                       &n_calls, &otus, &n_otus, &best, &ps_blocks, &hit_off, &call_off, &otu_off, &hits_out, &calls_out,
                       &otus_out};
        for (auto b : d) b->release();
        DevBuf *f[] = {&fam.table, &fam.ids, &fam.fam_func, &fam.fam_pgf, &fam.func_sid, &fam.hit_fam, &fam.E, &fam.gcap,
                       &fam.gofs, &fam.gscratch, &fam.matches, &fam.sofs, &fam.snd, &fam.sentries, &fam.sout_off, &fam.sout, &fam.class_seen, &fam.overflow};
        for (auto b : f) b->release();
        fam.h_scores.release();
        fam.h_score_off.release();
        post.release_all();
        famnr.release_all();
        for (auto &e : post_store) e.second.release_all();
        post_store.clear();
        DevBuf *q[] = {&fq.nfrag, &fq.naa, &fq.frag_base, &fq.res_base, &fq.frag_off, &fq.frag_res, &fq.best_frame,
                       &fq.best_score, &fq.best_n, &fq.best_first, &fq.match_off, &fq.matches};
        for (auto b : q) b->release();
        PinBuf *qh[] = {&fq.h_frag_base, &fq.h_frag_off, &fq.h_frag_res, &fq.h_best_frame, &fq.h_best_score, &fq.h_match_off,
                        &fq.h_matches};
        for (auto b : qh) b->release();
        for (int t = 0; t < kStageThreads; t++) {
            for (int k = 0; k < 2; k++) {
                stage_buf[t][k].release();
                if (stage_ev[t][k]) cudaEventDestroy(stage_ev[t][k]);
                stage_ev[t][k] = nullptr;
            }
            if (stage_stream[t]) cudaStreamDestroy(stage_stream[t]);
            stage_stream[t] = nullptr;
        }
        PinBuf *h[] = {&h_off, &h_totals, &h_hit_off, &h_hits, &h_call_off, &h_calls, &h_otu_off, &h_otus, &h_best, &h_fam};
        for (auto b : h) b->release();
    }
};
