// Family voting on the GPU: FamilyMapper::on_hit + find_best_family_match (family_mapper.cc:46-205, 287-330).
// Included at the end of ckm_api.cu (same translation unit: shares the ctx, CU/RC macros and run_device).
//
// Data model in HBM (read-only snapshot of KmerPegMapping, kmer.h:118-127):
//   * kmer_to_family_id_  -> open-addressed table of 16-byte slots {key+1, list offset, list length} keyed by
//     the 35-bit k-mer, at most half full, plus the concatenated family-id lists;
//   * family_data_        -> fam_func_sid[f] (interned function string), fam_pgf[f] (interned PGF string);
//   * function.index      -> func_sid[fI] (same interning), so the reference's STRING comparison
//     family.function == called function (family_mapper.cc:150-169) is an integer comparison here.
//
// Two kernels per batch, both one warp per protein, after K1 (hits with keys) and K2 (best call):
//   fam_lookup_kernel : lanes = hits; each probes the k-mer -> list table and records (offset, length); the
//                       warp sum of lengths E_i bounds the number of distinct families the protein touches.
//   fam_vote_kernel   : hits are consumed IN ORDER (each family's weighted_total is an f32 sum in hit order),
//                       lanes = entries of one hit's family list, accumulating into a per-warp open-addressed
//                       map in shared memory (E_i <= kFamSmemE) or in a global scratch region sized 2*E_i.
#pragma once

namespace ckm {

constexpr uint32_t kFamSmemCap = 1024;  // slots of the per-warp shared-memory maps
constexpr uint32_t kFamSmemE = 512;     // use them when the protein has at most this many list entries (cap = 2E)
constexpr int kFamStage = 8;            // list entries prefetched per hit (matrix_row_kernel)
constexpr uint32_t kFamEntries = 256;   // (family, weight) entries of 32 hits staged per step by fam_vote_kernel
// Two instantiations of fam_vote_kernel share one batch: SMALL (maps of <= kFamSmallCap slots, 8 warps per block, 4
// blocks per SM) takes the proteins whose map fits -- every fastq fragment does -- and LARGE (1024-slot maps, 4 warps per
// block, 2 blocks per SM) takes the rest, including those whose maps live in global scratch.
constexpr uint32_t kFamSmallCap = 256;
template <uint32_t CAP>
struct FamVoteCfg {
    static constexpr int kWarps = CAP <= kFamSmallCap ? 8 : 4;
    static constexpr uint32_t kWarpWords = 5 * CAP + 2 * kFamEntries;
    static constexpr size_t kSmem = (size_t)kWarps * kWarpWords * 4;
};

struct FamSlot {
    uint64_t key1;  // k-mer + 1, 0 = empty
    uint32_t off, cnt;
};

__device__ __forceinline__ uint64_t fam_hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void __launch_bounds__(256)
fam_build_kernel(const uint64_t *__restrict__ kmers, const uint64_t *__restrict__ fam_off, uint64_t n, FamSlot *table,
                 uint64_t mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key1 = kmers[i] + 1;
    uint64_t s = fam_hash(kmers[i]) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS((unsigned long long *)&table[s].key1, 0ull, (unsigned long long)key1);
        if (prev == 0ull || prev == key1) {  // a repeated k-mer keeps its last list
            table[s].off = (uint32_t)fam_off[i];
            table[s].cnt = (uint32_t)(fam_off[i + 1] - fam_off[i]);
            return;
        }
        s = (s + 1) & mask;
    }
}

struct FamTables {
    const FamSlot *table;
    uint64_t mask;
    const uint32_t *fam_ids;
    const uint32_t *fam_func_sid, *fam_pgf, *func_sid;
    uint32_t n_fams, n_functions, hypo_sid;
};

__global__ void __launch_bounds__(256)
fam_lookup_kernel(FamTables ft, const uint64_t *__restrict__ offsets, const uint64_t *__restrict__ hit_keys,
                  const uint32_t *__restrict__ n_hits, uint32_t n, uint2 *__restrict__ hit_fam, uint32_t *__restrict__ E,
                  uint32_t *__restrict__ gcap, uint32_t *__restrict__ class_seen = nullptr) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const uint64_t base = offsets[w];
    const uint32_t nh = n_hits[w];
    uint32_t e = 0;
    for (uint32_t k = lane; k < nh; k += 32) {
        const uint64_t key = hit_keys[base + k];
        uint64_t s = fam_hash(key) & ft.mask;
        uint2 r = make_uint2(0u, 0u);
        for (;;) {
            const FamSlot sl = ft.table[s];
            if (sl.key1 == key + 1) { r = make_uint2(sl.off, sl.cnt); break; }
            if (sl.key1 == 0) break;  // kmer_to_family_id_.find(...) == end (family_mapper.cc:296-297)
            s = (s + 1) & ft.mask;
        }
        hit_fam[base + k] = r;
        e += r.y;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
    if (lane == 0) {
        E[w] = e;
        uint32_t cap = 0;
        if (e > kFamSmemE) {  // power of two >= 2*E
            cap = 1u;
            while (cap < 2u * e && cap < 0x80000000u) cap <<= 1;
        }
        gcap[w] = cap;
        if (class_seen) {  // which fam_vote_kernel instantiations have work: [0] shared-memory maps, [1] global-scratch maps
            const uint32_t cls = cap == 0 ? 0u : 1u;
            if (*((volatile uint32_t *)class_seen + cls) == 0u) class_seen[cls] = 1u;
        }
    }
}

// optional second output of fam_vote_kernel: every (family, hit_count, weighted_total) of protein i at entries[ofs[i]...]
struct FamScoreOut {
    ckm_score_t *entries;
    const uint64_t *ofs;  // exclusive prefix sum of E
    uint32_t *n_distinct;
};

// compacts the per-protein score regions into CSR order, each protein's entries by ascending id (rank sort: ids are
// distinct within a protein)
__global__ void __launch_bounds__(256)
score_export_kernel(const ckm_score_t *__restrict__ entries, const uint64_t *__restrict__ ofs, const uint32_t *__restrict__ n_distinct,
                    const uint64_t *__restrict__ out_off, uint32_t n, ckm_score_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const ckm_score_t *src = entries + ofs[i];
    const uint32_t nd = n_distinct[i];
    ckm_score_t *dst = out + out_off[i];
    for (uint32_t k = lane; k < nd; k += 32) {
        const ckm_score_t e = src[k];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < nd; j++) rank += src[j].id < e.id;
        dst[rank] = e;
    }
}

// open-addressed insert-or-find of `id1` (id+1) in keys[0..mask]; returns the slot; *fresh = newly claimed
__device__ __forceinline__ uint32_t map_slot(uint32_t *keys, uint32_t mask, uint32_t id1, bool *fresh) {
    uint32_t s = (id1 * 2654435761u) & mask;
    for (;;) {
        const uint32_t prev = atomicCAS(&keys[s], 0u, id1);
        if (prev == 0u) { *fresh = true; return s; }
        if (prev == id1) { *fresh = false; return s; }
        s = (s + 1) & mask;
    }
}

template <uint32_t CAP>
__global__ void __launch_bounds__(FamVoteCfg<CAP>::kWarps * 32)
fam_vote_kernel(FamTables ft, const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ n_hits,
                const uint2 *__restrict__ hit_fam, const uint32_t *__restrict__ E, const uint32_t *__restrict__ gcap,
                const uint64_t *__restrict__ gofs, uint32_t *__restrict__ gscratch, const ckm_best_t *__restrict__ best,
                uint32_t n, ckm_family_match_t *__restrict__ out, FamScoreOut so, uint32_t *__restrict__ overflow,
                uint32_t *__restrict__ any_overflow) {
    extern __shared__ __align__(16) uint32_t fam_smem[];  // FamVoteCfg<CAP>::kSmem bytes, carved per warp below
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    uint32_t *const my_smem = fam_smem + (size_t)wib * FamVoteCfg<CAP>::kWarpWords;
    uint32_t *const s_keys = my_smem, *const s_cnt = my_smem + CAP;
    float *const s_w = reinterpret_cast<float *>(my_smem + 2 * CAP);
    uint32_t *const s_pkeys = my_smem + 3 * CAP;
    float *const s_pw = reinterpret_cast<float *>(my_smem + 4 * CAP);
    uint32_t *const s_fam = my_smem + 5 * CAP;
    float *const s_wt = reinterpret_cast<float *>(my_smem + 5 * CAP + kFamEntries);
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += n_warps) {
        uint32_t *keys, *cnt, *pkeys, cap;
        float *wsum, *pw;
        // shared-memory maps sized to the protein: at most E distinct families, kept at most half full.  E counts list
        // ENTRIES; a protein usually touches far fewer distinct families, so the SMALL instantiation takes every protein
        // whose map may live in shared memory with at most kFamSmallCap slots and hands the few that really hold more
        // than kFamSmallCap/2 families over to the LARGE one (overflow[i]), which runs after it.
        cap = 32u;
        while (cap < 2u * E[i]) cap <<= 1;
        constexpr bool kSmall = CAP <= kFamSmallCap;
        if (kSmall) {
            if (gcap[i] != 0) continue;
            cap = min(cap, CAP);
        } else if (gcap[i] == 0 && overflow[i] == 0u) {
            continue;  // done by the SMALL instantiation
        }
        const bool may_overflow = kSmall && 2u * E[i] > cap;
        uint32_t n_fresh = 0;
        bool bailed = false;
        const uint64_t base = offsets[i];
        const uint32_t nh = n_hits[i];
        if (gcap[i] == 0) {
            keys = s_keys; cnt = s_cnt; wsum = s_w; pkeys = s_pkeys; pw = s_pw;
            for (uint32_t s = lane; s < cap; s += 32) { keys[s] = 0u; pkeys[s] = 0u; }
        } else {  // global scratch: 5 arrays of `cap` words, keys pre-zeroed by the host-side memset
            cap = gcap[i];
            uint32_t *g = gscratch + gofs[i] * 5ull;
            keys = g; cnt = g + cap; wsum = (float *)(g + 2ull * cap); pkeys = g + 3ull * cap; pw = (float *)(g + 4ull * cap);
        }
        const uint32_t mask = cap - 1;
        __syncwarp();

        // ---- F1: on_hit for every hit, in position order (family_mapper.cc:287-330) ----
        // A family's weighted_total is an f32 sum in hit order, so additions to ONE family are a serial chain -- but the
        // map look-ups are not.  Per step of 32 hits: every hit lane writes its (family, 1/|list|) entries, in hit order,
        // into a shared-memory list; then lanes = entries: each finds or claims its family's slot (one atomicCAS latency
        // for 32 entries instead of one per hit), entries that share a slot are ranked by lane (= hit order) with
        // match.any, and the additions are applied rank by rank.
        if (E[i] != 0) {
            for (uint32_t k0 = 0; k0 < nh; k0 += 32) {
                const uint2 my = (k0 + lane < nh) ? hit_fam[base + k0 + lane] : make_uint2(0u, 0u);
                uint32_t incl = my.y;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += t;
                }
                const uint32_t n_ent = __shfl_sync(0xffffffffu, incl, 31);
                if (n_ent == 0) continue;
                if (may_overflow && (n_ent > kFamEntries || n_fresh > cap / 2u)) {  // warp-uniform
                    bailed = true;
                    break;
                }
                if (n_ent <= kFamEntries) {
                    const float weight = 1.0f / (float)my.y;  // 1.0f / counts.size(), family_mapper.cc:300
                    uint32_t e = incl - my.y;
                    for (uint32_t t = 0; t < my.y; t++, e++) {
                        s_fam[e] = ft.fam_ids[my.x + t];
                        s_wt[e] = weight;
                    }
                    __syncwarp();
                    for (uint32_t e0 = 0; e0 < n_ent; e0 += 32) {
                        // SMALL: at most kFamSmallCap/2 families before a tile and 32 more in it -- the map cannot fill up
                        if (may_overflow && n_fresh > cap / 2u) break;
                        const uint32_t me = e0 + lane;
                        const uint32_t fam = me < n_ent ? s_fam[me] : 0xffffffffu;
                        const bool ok = fam < ft.n_fams;  // no family_data_ entry: never reported (146-148)
                        uint32_t slot = 0x80000000u | lane;  // distinct for idle lanes
                        float w = 0.0f;
                        bool fresh = false;
                        if (ok) {
                            slot = map_slot(keys, mask, fam + 1, &fresh);
                            if (fresh) { cnt[slot] = 0u; wsum[slot] = 0.0f; }
                            w = s_wt[me];
                        }
                        __syncwarp();
                        if (may_overflow) n_fresh += __popc(__ballot_sync(0xffffffffu, fresh));
                        const uint32_t peers = __match_any_sync(0xffffffffu, slot);
                        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                        const uint32_t last = __reduce_max_sync(0xffffffffu, ok ? rank : 0u);
                        for (uint32_t r = 0; r <= last; r++) {
                            if (ok && rank == r) { cnt[slot] += 1u; wsum[slot] += w; }
                            __syncwarp();
                        }
                    }
                } else {  // a step with very long lists: hit by hit, lanes = entries of one hit
                    const uint32_t lim = min(32u, nh - k0);
                    for (uint32_t j = 0; j < lim; j++) {
                        const uint32_t off_j = __shfl_sync(0xffffffffu, my.x, j);
                        const uint32_t cnt_j = __shfl_sync(0xffffffffu, my.y, j);
                        if (cnt_j == 0) continue;
                        const float weight = 1.0f / (float)cnt_j;
                        for (uint32_t t = lane; t < cnt_j; t += 32) {
                            const uint32_t fam = ft.fam_ids[off_j + t];
                            if (fam >= ft.n_fams) continue;
                            bool fresh;
                            const uint32_t sl = map_slot(keys, mask, fam + 1, &fresh);
                            if (fresh) { cnt[sl] = 1u; wsum[sl] = 0.0f + weight; }
                            else { cnt[sl] += 1u; wsum[sl] += weight; }
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
        }

        if (kSmall) {
            if (may_overflow && n_fresh > cap / 2u) bailed = true;
            if (lane == 0) {
                overflow[i] = bailed ? 1u : 0u;
                if (bailed && *((volatile uint32_t *)any_overflow) == 0u) *any_overflow = 1u;
            }
            if (bailed) {
                __syncwarp();
                continue;
            }
        }

        // ---- seq_score_ as it stands after the hits (LookupRequest::on_hit, lookup_request.cc:441-464), for callers
        // that report every family: (id, hit_count, weighted_total) per distinct family, in map-slot order ----
        if (so.entries != nullptr) {
            ckm_score_t *dst = so.entries + so.ofs[i];
            uint32_t total = 0;
            if (E[i] != 0) {
                for (uint32_t s0 = 0; s0 < cap; s0 += 32) {
                    const uint32_t s = s0 + lane;
                    const uint32_t key = keys[s];
                    const bool q = key != 0u;
                    const uint32_t m = __ballot_sync(0xffffffffu, q);
                    if (q) {
                        ckm_score_t e;
                        e.id = key - 1;
                        e.hit_count = cnt[s];
                        e.weighted_total = wsum[s];
                        dst[total + __popc(m & ((1u << lane) - 1u))] = e;
                    }
                    total += __popc(m);
                }
            }
            if (lane == 0) so.n_distinct[i] = total;
        }

        // ---- F2: best call -> matching families -> best PLF / rolled-up PGF (family_mapper.cc:98-204) ----
        const ckm_best_t b = best[i];
        int32_t fidx = -1;
        uint32_t sid = ft.hypo_sid;
        if (b.function_index >= 0 && (uint32_t)b.function_index < ft.n_functions) {
            fidx = b.function_index;
            sid = ft.func_sid[fidx];
        }
        float lbest = 0.0f;
        uint32_t lfam = 0xffffffffu;
        if (E[i] != 0) {
            for (uint32_t s0 = 0; s0 < cap; s0 += 32) {
                const uint32_t s = s0 + lane;
                const uint32_t key = keys[s];
                bool q = false;
                float w = 0.0f;
                uint32_t pg = 0;
                if (key != 0u && cnt[s] >= 3u && ft.fam_func_sid[key - 1] == sid) {  // kmer_hit_threshold_ = 3
                    q = true;
                    w = wsum[s];
                    pg = ft.fam_pgf[key - 1];
                    if (w > lbest || (w == lbest && w > 0.0f && key - 1 < lfam)) { lbest = w; lfam = key - 1; }
                }
                // pgf_rollup[pgf] += weighted_total, one qualifying family at a time in slot order
                uint32_t m = __ballot_sync(0xffffffffu, q);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t pg_b = __shfl_sync(0xffffffffu, pg, src);
                    const float w_b = __shfl_sync(0xffffffffu, w, src);
                    if (lane == 0) {
                        bool fresh;
                        const uint32_t ps = map_slot(pkeys, mask, pg_b + 1, &fresh);
                        if (fresh) pw[ps] = 0.0f + w_b;
                        else pw[ps] += w_b;
                    }
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {  // max weighted_total; exact ties -> smallest family id
            const float ow = __shfl_xor_sync(0xffffffffu, lbest, d);
            const uint32_t of = __shfl_xor_sync(0xffffffffu, lfam, d);
            if (ow > lbest || (ow == lbest && of < lfam)) { lbest = ow; lfam = of; }
        }
        float gbest = 0.0f;
        uint32_t gfam = 0xffffffffu;
        if (lfam != 0xffffffffu) {
            for (uint32_t s = lane; s < cap; s += 32) {
                const uint32_t key = pkeys[s];
                if (key != 0u) {
                    const float w = pw[s];
                    if (w > gbest || (w == gbest && w > 0.0f && key - 1 < gfam)) { gbest = w; gfam = key - 1; }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float ow = __shfl_xor_sync(0xffffffffu, gbest, d);
                const uint32_t of = __shfl_xor_sync(0xffffffffu, gfam, d);
                if (ow > gbest || (ow == gbest && of < gfam)) { gbest = ow; gfam = of; }
            }
        }
        if (lane == 0) {
            ckm_family_match_t r;
            r.gfam = gfam == 0xffffffffu ? -1 : (int32_t)gfam;
            r.lfam = lfam == 0xffffffffu ? -1 : (int32_t)lfam;
            r.gfam_score = gbest;
            r.lfam_score = lbest;
            r.score = b.score;
            r.function_index = fidx;
            out[i] = r;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) fam_validate_kernel(const uint64_t *__restrict__ fam_off, const uint32_t *__restrict__ fam_ids,
                                                           uint64_t n_kmers, unsigned int *__restrict__ dup_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_kmers) return;
    const uint64_t a = fam_off[i], b = fam_off[i + 1];
    for (uint64_t x = a; x < b; x++)
        for (uint64_t y = x + 1; y < b; y++)
            if (fam_ids[x] == fam_ids[y]) atomicOr(dup_flag, 1u);
}

}  // namespace ckm

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#include <unordered_map>

// family_data_ -> interned ids on the device + host strings (shared by ckm_family_load and ckm_family_nr_finish)
static int family_install_metadata(ckm_ctx *c, uint32_t n_families, const char *const *pgf, const char *const *plf,
                                   const char *const *function) {
    ckm_ctx::Family &F = c->fam;
    // intern function and PGF strings (family_mapper.cc:150-169 compares strings)
    std::unordered_map<std::string, uint32_t> sids, pgfs;
    auto intern = [](std::unordered_map<std::string, uint32_t> &m, const std::string &s) {
        auto it = m.find(s);
        if (it != m.end()) return it->second;
        const uint32_t id = (uint32_t)m.size();
        m.emplace(s, id);
        return id;
    };
    const uint32_t hypo = intern(sids, "hypothetical protein");
    std::vector<uint32_t> func_sid(c->functions.size()), fam_func(n_families), fam_pgf(n_families);
    for (size_t i = 0; i < c->functions.size(); i++) {
        const std::string &s = c->functions[i];
        // an empty or "A ?? B" name is "hypothetical protein" to find_best_family_match (103-123)
        func_sid[i] = (s.empty() || s.find(" ?? ") != std::string::npos) ? hypo : intern(sids, s);
    }
    F.pgf_names.clear();
    F.plf.assign(n_families, std::string());
    for (uint32_t f = 0; f < n_families; f++) {
        fam_func[f] = intern(sids, function[f]);
        const size_t before = pgfs.size();
        fam_pgf[f] = intern(pgfs, pgf[f]);
        if (pgfs.size() != before) F.pgf_names.emplace_back(pgf[f]);
        F.plf[f] = plf[f];
    }
    F.n_fams = n_families;
    F.n_functions = (uint32_t)c->functions.size();
    F.hypo_sid = hypo;
    RC(F.fam_func.ensure(((size_t)n_families + 1) * 4));
    RC(F.fam_pgf.ensure(((size_t)n_families + 1) * 4));
    RC(F.func_sid.ensure((func_sid.size() + 1) * 4));
    if (n_families) {
        CU(cudaMemcpyAsync(F.fam_func.p, fam_func.data(), (size_t)n_families * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(F.fam_pgf.p, fam_pgf.data(), (size_t)n_families * 4, cudaMemcpyHostToDevice, c->stream));
    }
    if (!func_sid.empty())
        CU(cudaMemcpyAsync(F.func_sid.p, func_sid.data(), func_sid.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
    return 0;
}

extern "C" int ckm_family_load(ckm_ctx *c, uint64_t n_kmers, const uint64_t *kmers, const uint64_t *fam_offsets,
                               const uint32_t *fam_ids, uint32_t n_families, const char *const *pgf, const char *const *plf,
                               const char *const *function) {
    if (!c) return ckm_fail(CKM_EINVAL, "ctx is NULL");
    if (n_kmers && (!kmers || !fam_offsets || !fam_ids)) return ckm_fail(CKM_EINVAL, "NULL family table");
    if (c->shares_tables) return ckm_fail(CKM_ESTATE, "a clone reads its parent's family tables and cannot load its own");
    CU(cudaSetDevice(c->device));
    const uint64_t n_entries = n_kmers ? fam_offsets[n_kmers] : 0;
    if (n_entries >= (1ull << 32)) return ckm_fail(CKM_EINVAL, "more than 2^32 k-mer->family entries");
    ckm_ctx::Family &F = c->fam;
    RC(family_install_metadata(c, n_families, pgf, plf, function));
    uint64_t cap = 16;
    while (cap < 2 * n_kmers) cap <<= 1;
    F.mask = cap - 1;
    RC(F.table.ensure(cap * sizeof(FamSlot)));
    RC(F.ids.ensure((n_entries + 1) * 4));
    DevBuf d_k, d_o, d_flag;
    RC(d_k.ensure((n_kmers + 1) * 8));
    RC(d_o.ensure((n_kmers + 2) * 8));
    RC(d_flag.ensure(64));
    CU(cudaMemsetAsync(F.table.p, 0, cap * sizeof(FamSlot), c->stream));
    CU(cudaMemsetAsync(d_flag.p, 0, 4, c->stream));
    if (n_kmers) {
        CU(cudaMemcpyAsync(d_k.p, kmers, n_kmers * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_o.p, fam_offsets, (n_kmers + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        if (n_entries) CU(cudaMemcpyAsync(F.ids.p, fam_ids, n_entries * 4, cudaMemcpyHostToDevice, c->stream));
    }
    unsigned int dup = 0;
    if (n_kmers) {
        const unsigned blocks = (unsigned)((n_kmers + 255) / 256);
        fam_validate_kernel<<<blocks, 256, 0, c->stream>>>((const uint64_t *)d_o.p, (const uint32_t *)F.ids.p, n_kmers,
                                                          (unsigned int *)d_flag.p);
        fam_build_kernel<<<blocks, 256, 0, c->stream>>>((const uint64_t *)d_k.p, (const uint64_t *)d_o.p, n_kmers,
                                                       (FamSlot *)F.table.p, F.mask);
        c->launches += 2;
    }
    CU(cudaMemcpyAsync(&dup, d_flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    d_k.release();
    d_o.release();
    d_flag.release();
    if (dup) return ckm_fail(CKM_EINVAL, "a k-mer's family list contains a repeated family id (kmer.cc:216-230 dedupes them)");
    F.loaded = true;
    return 0;
}

extern "C" const char *ckm_family_pgf_name(const ckm_ctx *c, int32_t g) {
    return (g >= 0 && (size_t)g < c->fam.pgf_names.size()) ? c->fam.pgf_names[g].c_str() : "";
}
extern "C" const char *ckm_family_plf_name(const ckm_ctx *c, int32_t l) {
    return (l >= 0 && (size_t)l < c->fam.plf.size()) ? c->fam.plf[l].c_str() : "";
}
extern "C" const char *ckm_family_function_name(const ckm_ctx *c, const ckm_family_match_t *m) {
    if (m->function_index >= 0 && (size_t)m->function_index < c->functions.size()) {
        const std::string &s = c->functions[m->function_index];
        if (!s.empty() && s.find(" ?? ") == std::string::npos) return s.c_str();
    }
    return "hypothetical protein";
}

// K1 (hits + keys) and K2 (calls + best) must already have run on (d_res, d_off); leaves matches in c->fam.matches
static int family_device(ckm_ctx *c, const uint64_t *d_off, uint32_t n, uint64_t total, bool want_scores = false) {
    ckm_ctx::Family &F = c->fam;
    RC(F.hit_fam.ensure((total + 1) * sizeof(uint2)));
    RC(F.E.ensure(((size_t)n + 1) * 4));
    RC(F.gcap.ensure(((size_t)n + 1) * 4));
    RC(F.gofs.ensure(((size_t)n + 2) * 8));
    RC(F.matches.ensure(((size_t)n + 1) * sizeof(ckm_family_match_t)));
    if (n == 0) return 0;
    FamTables ft;
    ft.table = (const FamSlot *)F.table.p;
    ft.mask = F.mask;
    ft.fam_ids = (const uint32_t *)F.ids.p;
    ft.fam_func_sid = (const uint32_t *)F.fam_func.p;
    ft.fam_pgf = (const uint32_t *)F.fam_pgf.p;
    ft.func_sid = (const uint32_t *)F.func_sid.p;
    ft.n_fams = F.n_fams;
    ft.n_functions = F.n_functions;
    ft.hypo_sid = F.hypo_sid;
    const unsigned lb = (unsigned)(((uint64_t)n * 32 + 255) / 256);
    RC(F.class_seen.ensure(64));
    RC(F.overflow.ensure(((size_t)n + 1) * 4));
    CU(cudaMemsetAsync(F.class_seen.p, 0, 16, c->stream));
    fam_lookup_kernel<<<lb, 256, 0, c->stream>>>(ft, d_off, (const uint64_t *)c->hit_keys.p, (const uint32_t *)c->n_hits.p, n,
                                                 (uint2 *)F.hit_fam.p, (uint32_t *)F.E.p, (uint32_t *)F.gcap.p,
                                                 (uint32_t *)F.class_seen.p);
    c->launches++;
    RC(prefix_sum(c, (const uint32_t *)F.gcap.p, n, (uint64_t *)F.gofs.p));
    uint64_t gtotal = 0;
    uint32_t class_seen[2] = {1u, 1u};
    CU(cudaMemcpyAsync(&gtotal, (const uint64_t *)F.gofs.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(class_seen, F.class_seen.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (gtotal) {
        RC(F.gscratch.ensure(gtotal * 5 * 4));
        CU(cudaMemsetAsync(F.gscratch.p, 0, gtotal * 5 * 4, c->stream));
    }
    FamScoreOut so = {nullptr, nullptr, nullptr};
    if (want_scores) {  // regions of E_i entries: a protein touches at most that many distinct families
        RC(F.sofs.ensure(((size_t)n + 2) * 8));
        RC(F.snd.ensure(((size_t)n + 1) * 4));
        RC(F.sout_off.ensure(((size_t)n + 2) * 8));
        RC(prefix_sum(c, (const uint32_t *)F.E.p, n, (uint64_t *)F.sofs.p));
        uint64_t etotal = 0;
        CU(cudaMemcpyAsync(&etotal, (const uint64_t *)F.sofs.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        RC(F.sentries.ensure((etotal + 1) * sizeof(ckm_score_t)));
        so.entries = (ckm_score_t *)F.sentries.p;
        so.ofs = (const uint64_t *)F.sofs.p;
        so.n_distinct = (uint32_t *)F.snd.p;
    }
    {
        typedef FamVoteCfg<kFamSmallCap> S;
        typedef FamVoteCfg<kFamSmemCap> L;
        CU(cudaFuncSetAttribute(fam_vote_kernel<kFamSmemCap>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kSmem));
        CU(cudaFuncSetAttribute(fam_vote_kernel<kFamSmallCap>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kSmem));
        const unsigned sb = (unsigned)std::min<uint64_t>(((uint64_t)n + S::kWarps - 1) / S::kWarps, (uint64_t)c->sm_count * 32);
        const unsigned lb2 = (unsigned)std::min<uint64_t>(((uint64_t)n + L::kWarps - 1) / L::kWarps, (uint64_t)c->sm_count * 8);
        // SMALL takes every protein whose map fits shared memory and flags the few that hold more than kFamSmallCap/2
        // distinct families; LARGE takes those and the proteins with global-scratch maps.  Each instantiation walks the
        // whole batch, so it is launched only if it has something to do (every fastq fragment is "small").
        uint32_t *ovf = (uint32_t *)F.overflow.p, *any_ovf = (uint32_t *)F.class_seen.p + 2;
        if (class_seen[0]) {
            fam_vote_kernel<kFamSmallCap><<<sb, S::kWarps * 32, S::kSmem, c->stream>>>(
                ft, d_off, (const uint32_t *)c->n_hits.p, (const uint2 *)F.hit_fam.p, (const uint32_t *)F.E.p,
                (const uint32_t *)F.gcap.p, (const uint64_t *)F.gofs.p, (uint32_t *)F.gscratch.p, (const ckm_best_t *)c->best.p, n,
                (ckm_family_match_t *)F.matches.p, so, ovf, any_ovf);
            c->launches++;
        }
        bool need_large = class_seen[1] != 0;
        if (!need_large && class_seen[0]) {
            uint32_t any = 0;
            CU(cudaMemcpyAsync(&any, any_ovf, 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            need_large = any != 0;
        }
        if (need_large) {
            fam_vote_kernel<kFamSmemCap><<<lb2, L::kWarps * 32, L::kSmem, c->stream>>>(
                ft, d_off, (const uint32_t *)c->n_hits.p, (const uint2 *)F.hit_fam.p, (const uint32_t *)F.E.p,
                (const uint32_t *)F.gcap.p, (const uint64_t *)F.gofs.p, (uint32_t *)F.gscratch.p, (const ckm_best_t *)c->best.p, n,
                (ckm_family_match_t *)F.matches.p, so, ovf, any_ovf);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    return 0;
}

// LookupRequest's per-sequence accumulation in family mode (lookup_request.cc:138-166, 441-464): every family a
// sequence's hits touch, with hit_count (= hit_total) and weighted_total, plus find_best_call and the FamilyMapper
// match of ckm_family_batch, in one pass
extern "C" int ckm_family_scores(ckm_ctx *c, const char *residues, const uint64_t *offsets, uint32_t n, ckm_family_scores_t *out) {
    if (!c || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    if (!c->fam.loaded) return ckm_fail(CKM_ESTATE, "ckm_family_scores before ckm_family_load");
    memset(out, 0, sizeof *out);
    ckm_ctx::Family &F = c->fam;
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, residues, offsets, n, &total, &max_len));
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(max_len, 1u),
                  CKM_WANT_HITS | CKM_WANT_CALLS | CKM_WANT_BEST));
    RC(family_device(c, (const uint64_t *)c->in_off.p, n, total, true));
    uint64_t ns = 0;
    if (n) {
        RC(prefix_sum(c, (const uint32_t *)F.snd.p, n, (uint64_t *)F.sout_off.p));
        CU(cudaMemcpyAsync(&ns, (const uint64_t *)F.sout_off.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    RC(F.sout.ensure((ns + 1) * sizeof(ckm_score_t)));
    RC(F.h_scores.ensure((ns + 1) * sizeof(ckm_score_t)));
    RC(F.h_score_off.ensure(((size_t)n + 2) * 8));
    RC(c->h_fam.ensure(((size_t)n + 1) * sizeof(ckm_family_match_t)));
    RC(c->h_best.ensure(((size_t)n + 1) * sizeof(ckm_best_t)));
    if (n) {
        score_export_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, c->stream>>>(
            (const ckm_score_t *)F.sentries.p, (const uint64_t *)F.sofs.p, (const uint32_t *)F.snd.p, (const uint64_t *)F.sout_off.p, n,
            (ckm_score_t *)F.sout.p);
        c->launches++;
        if (ns) CU(cudaMemcpyAsync(F.h_scores.p, F.sout.p, ns * sizeof(ckm_score_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(F.h_score_off.p, F.sout_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(c->h_fam.p, F.matches.p, (size_t)n * sizeof(ckm_family_match_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(c->h_best.p, c->best.p, (size_t)n * sizeof(ckm_best_t), cudaMemcpyDeviceToHost, c->stream));
    } else {
        *(uint64_t *)F.h_score_off.p = 0;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    out->n = n;
    out->scores = (const ckm_score_t *)F.h_scores.p;
    out->score_offsets = (const uint64_t *)F.h_score_off.p;
    out->best = (const ckm_best_t *)c->h_best.p;
    out->matches = (const ckm_family_match_t *)c->h_fam.p;
    return 0;
}

extern "C" int ckm_family_batch(ckm_ctx *c, const char *residues, const uint64_t *offsets, uint32_t n,
                                const ckm_family_match_t **matches) {
    if (!c || !matches) return ckm_fail(CKM_EINVAL, "NULL argument");
    if (!c->fam.loaded) return ckm_fail(CKM_ESTATE, "ckm_family_batch before ckm_family_load");
    *matches = nullptr;
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, residues, offsets, n, &total, &max_len));
    // ingest_protein: process_aa_seq(id, seq, calls, on_hit, 0) then find_best_call (family_mapper.cc:46-63, 98)
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(max_len, 1u),
                  CKM_WANT_HITS | CKM_WANT_CALLS | CKM_WANT_BEST));
    RC(family_device(c, (const uint64_t *)c->in_off.p, n, total));
    RC(c->h_fam.ensure(((size_t)n + 1) * sizeof(ckm_family_match_t)));
    if (n)
        CU(cudaMemcpyAsync(c->h_fam.p, c->fam.matches.p, (size_t)n * sizeof(ckm_family_match_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    *matches = (const ckm_family_match_t *)c->h_fam.p;
    return 0;
}
