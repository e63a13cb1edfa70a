// Internal: the context behind ckm_ctx* (device table, parameters, grow-only work buffers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
    cudaStream_t stream = nullptr;
    bool force_raw = false;
    int l2_fetch = 0;  // cudaLimitMaxL2FetchGranularity in effect

    // signature table in HBM
    DevBuf table;
    uint64_t num_sigs = 0, magic = 0;
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
    DevBuf in_res, in_off, totals, hits, hit_keys, hit_avg, n_hits, stored_idx, calls, calls_work, n_calls;
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
    } fam;
    PinBuf h_fam;

    // pinned host buffers handed out through ckm_batch_out_t
    PinBuf h_off, h_totals, h_hit_off, h_hits, h_call_off, h_calls, h_otu_off, h_otus, h_best;

    void free_all() {
        DevBuf *d[] = {&table, &in_res, &in_off, &totals, &hits, &hit_keys, &hit_avg, &n_hits, &stored_idx, &calls, &calls_work,
                       &n_calls, &otus, &n_otus, &best, &ps_blocks, &hit_off, &call_off, &otu_off, &hits_out, &calls_out,
                       &otus_out};
        for (auto b : d) b->release();
        DevBuf *f[] = {&fam.table, &fam.ids, &fam.fam_func, &fam.fam_pgf, &fam.func_sid, &fam.hit_fam, &fam.E, &fam.gcap,
                       &fam.gofs, &fam.gscratch, &fam.matches};
        for (auto b : f) b->release();
        PinBuf *h[] = {&h_off, &h_totals, &h_hit_off, &h_hits, &h_call_off, &h_calls, &h_otu_off, &h_otus, &h_best, &h_fam};
        for (auto b : h) b->release();
    }
};
