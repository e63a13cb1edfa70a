// Host side of the request handlers, written only against the C ABI (include/ckm.h): one batch call per
// body chunk, then the reference's response text rebuilt from the flat results.  Floats are printed
// exactly as std::ostream prints them in the reference (default precision 6; see text.h), so the text is byte-identical.
#include "../../include/ckm_handlers.h"
#include "text.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <utility>
#include <map>
#include <unordered_map>
#include <vector>

namespace {

char *dup_text(const std::string &s) {
    char *p = (char *)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

// KmerGuts::format_call, kguts.cc:939-947
void put_call(ckm_text::Text &os, const ckm_ctx *ctx, const ckm_call_t &c) {
    os << "CALL\t" << c.start << "\t" << c.end << "\t" << c.count;
    os << "\t" << c.function_index << "\t" << ckm_function_at_index(ctx, (int32_t)c.function_index);
    os << "\t" << c.weighted_hits << "\n";
}

// KmerGuts::format_hit, kguts.cc:949-959
void put_hit(ckm_text::Text &os, const ckm_ctx *ctx, const ckm_hit_t &h) {
    char dc[CKM_KMER_SIZE + 1];
    ckm_decoded_kmer(h.which_kmer, dc);
    os << "HIT\t" << h.offset << "\t" << dc << "\t" << h.avg_from_end << "\t" << ckm_function_at_index(ctx, h.function_index)
       << "\t" << h.function_wt << "\t" << h.otu_index << "\n";
}

// KmerOtuStats::finalize (kguts.h:214-218) + format_otu_stats (kguts.cc:961-973)
void put_otu_stats(ckm_text::Text &os, const std::string &id, uint64_t size, const ckm_otu_t *otus, uint64_t n) {
    std::vector<std::pair<int, int>> by_count;
    by_count.reserve(n);
    for (uint64_t i = 0; i < n; i++) by_count.emplace_back(otus[i].otu_index, otus[i].count);
    std::sort(by_count.begin(), by_count.end(),
              [](const std::pair<int, int> &lhs, const std::pair<int, int> &rhs) { return rhs.second < lhs.second; });
    os << "OTU-COUNTS\t" << id << "[" << size << "]";
    const size_t top = std::min(by_count.size(), (size_t)5);
    for (size_t i = 0; i < top; i++) os << "\t" << by_count[i].second << "-" << by_count[i].first;
    os << "\n";
}

// find_best_call's `function` out-parameter, kguts.cc:1160 and 1176-1196
std::string best_function(const ckm_ctx *ctx, const ckm_best_t &b) {
    if (b.flags & CKM_BEST_AMBIG) {
        std::string f1 = ckm_function_at_index(ctx, b.ambig_a);
        std::string f2 = ckm_function_at_index(ctx, b.ambig_b);
        if (f2 > f1) std::swap(f1, f2);
        return f1 + " ?? " + f2;
    }
    if (b.function_index >= 0) return ckm_function_at_index(ctx, b.function_index);
    return "";
}

}  // namespace

extern "C" {

int ckm_query_text(ckm_ctx *ctx, const char *const *ids, const char *residues, const uint64_t *offsets, uint32_t n,
                   int details, int find_best_call, char **text) {
    if (!text) return CKM_EINVAL;
    *text = nullptr;
    // the handler always builds calls + OTU stats; hits only with details=1 (query_request.cc:105-119)
    uint32_t flags = CKM_WANT_CALLS | CKM_WANT_OTU;
    if (details) flags |= CKM_WANT_HITS;
    if (find_best_call) flags |= CKM_WANT_BEST;
    ckm_batch_out_t o;
    int rc = ckm_call_batch(ctx, residues, offsets, n, flags, &o);
    if (rc) return rc;
    ckm_text::Text os;
    for (uint32_t i = 0; i < n; i++) {
        const std::string id = ids[i];
        const uint64_t len = offsets[i + 1] - offsets[i];
        if (find_best_call) {  // query_request.cc:124-135
            const std::string fn = best_function(ctx, o.best[i]);
            if (!fn.empty()) os << id << "\t" << fn << "\t" << o.best[i].score << "\t" << o.best[i].weighted_score << "\n";
        } else {  // 136-151
            os << "PROTEIN-ID\t" << id << "\t" << len << "\n";
            for (uint64_t k = o.call_offsets[i]; k < o.call_offsets[i + 1]; k++) put_call(os, ctx, o.calls[k]);
            if (details)
                for (uint64_t k = o.hit_offsets[i]; k < o.hit_offsets[i + 1]; k++) put_hit(os, ctx, o.hits[k]);
            put_otu_stats(os, id, len, o.otus + o.otu_offsets[i], o.otu_offsets[i + 1] - o.otu_offsets[i]);
        }
    }
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

// operator<<(ostream&, best_match_t), family_mapper.h:70-75
static void put_match(ckm_text::Text &os, const ckm_ctx *ctx, const ckm_family_match_t &m) {
    os << ckm_family_pgf_name(ctx, m.gfam) << "\t" << m.gfam_score << "\t" << ckm_family_plf_name(ctx, m.lfam) << "\t"
       << m.lfam_score << "\t" << ckm_family_function_name(ctx, &m) << "\t" << m.score;
}

int ckm_family_text(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n, char **text) {
    if (!text) return CKM_EINVAL;
    *text = nullptr;
    const ckm_family_match_t *m = nullptr;
    int rc = ckm_family_batch(ctx, residues, offsets, n, &m);
    if (rc) return rc;
    ckm_text::Text os;
    for (uint32_t i = 0; i < n; i++) {
        put_match(os, ctx, m[i]);
        os << "\n";
    }
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

int ckm_fq_text(ckm_ctx *ctx, const char *const *ids, const char *bases, const uint64_t *offsets, uint32_t n, char **text) {
    if (!text) return CKM_EINVAL;
    *text = nullptr;
    ckm_fq_out_t o;
    int rc = ckm_fq_batch(ctx, bases, offsets, n, &o);
    if (rc) return rc;
    ckm_text::Text os;
    for (uint32_t r = 0; r < n; r++) {
        const std::string id = ids[r];
        if (id.empty()) continue;                // fq_process_request.cc:301-302
        if (!(o.best_score[r] > 0.0)) continue;  // 349
        os << id << "\t" << o.best_frame[r] << "\t" << o.best_score[r] << "\t";
        for (uint64_t k = o.match_offsets[r]; k < o.match_offsets[r + 1]; k++) {
            if (k != o.match_offsets[r]) os << "\t";
            os << o.matches[k].length << "\t";
            put_match(os, ctx, o.matches[k].m);
        }
        os << "\n";
    }
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

}  // extern "C"

struct ckm_mapping {
    std::unordered_map<std::string, uint32_t> peg_to_id;
    std::vector<std::string> id_to_peg;
};

extern "C" {

ckm_mapping *ckm_mapping_new(void) { return new ckm_mapping(); }
void ckm_mapping_free(ckm_mapping *m) { delete m; }
uint32_t ckm_mapping_encode_id(ckm_mapping *m, const char *peg) {  // kmer.cc:272-286
    auto it = m->peg_to_id.find(peg);
    if (it != m->peg_to_id.end()) return it->second;
    const uint32_t id = (uint32_t)m->id_to_peg.size();
    m->peg_to_id.emplace(peg, id);
    m->id_to_peg.emplace_back(peg);
    return id;
}
uint32_t ckm_mapping_assign_new_id(ckm_mapping *m, const char *peg) {  // kmer.h:109-116: a fresh id even for a known peg
    const uint32_t id = (uint32_t)m->id_to_peg.size();
    m->peg_to_id[peg] = id;
    m->id_to_peg.emplace_back(peg);
    return id;
}
const char *ckm_mapping_decode_id(const ckm_mapping *m, uint32_t id) {  // kmer.cc:288-295
    return id < m->id_to_peg.size() ? m->id_to_peg[id].c_str() : "";
}

int ckm_add_text(ckm_ctx *ctx, ckm_mapping *m, const char *const *ids, const char *residues, const uint64_t *offsets,
                 uint32_t n, int silent, char **text) {
    if (!text || !m) return CKM_EINVAL;
    *text = nullptr;
    // process_aa_seq_hits(id, seq, calls, hits, stats) + find_best_call (add_request.cc:133-146)
    ckm_batch_out_t o;
    int rc = ckm_call_batch(ctx, residues, offsets, n, CKM_WANT_CALLS | CKM_WANT_HITS | CKM_WANT_OTU | CKM_WANT_BEST, &o);
    if (rc) return rc;
    ckm_text::Text os;
    std::vector<uint32_t> eids(n);
    for (uint32_t i = 0; i < n; i++) {
        const std::string id = ids[i];
        const uint64_t len = offsets[i + 1] - offsets[i];
        if (!silent) {
            os << "PROTEIN-ID\t" << id << "\t" << len << "\n";
            for (uint64_t k = o.call_offsets[i]; k < o.call_offsets[i + 1]; k++) put_call(os, ctx, o.calls[k]);
            put_otu_stats(os, id, len, o.otus + o.otu_offsets[i], o.otu_offsets[i + 1] - o.otu_offsets[i]);
            std::string fn = best_function(ctx, o.best[i]);
            if (fn.empty() || fn.find(" ?? ") != std::string::npos) fn = "hypothetical protein";  // 147-158
            os << "BEST-CALL\t" << id << "\t" << fn << "\t" << o.best[i].score << "\t" << o.best[i].weighted_score << "\t"
               << o.best[i].score_offset << "\n";
        }
        eids[i] = ckm_mapping_encode_id(m, ids[i]);  // add_request.cc:164
    }
    rc = ckm_postings_append_last(ctx, eids.data(), n);  // 165-170
    if (rc) return rc;
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

uint64_t ckm_matrix_merge_pairs(ckm_pair_t *pairs, uint64_t n_pairs) {
    std::sort(pairs, pairs + n_pairs, [](const ckm_pair_t &a, const ckm_pair_t &b) {
        return a.eid_i != b.eid_i ? a.eid_i < b.eid_i : a.eid_j < b.eid_j;
    });
    uint64_t w = 0;
    for (uint64_t r = 0; r < n_pairs; r++) {
        if (w && pairs[w - 1].eid_i == pairs[r].eid_i && pairs[w - 1].eid_j == pairs[r].eid_j)
            pairs[w - 1].count += pairs[r].count;  // a repeated id in the request shares one map entry
        else
            pairs[w++] = pairs[r];
    }
    return w;
}

int ckm_matrix_text(ckm_ctx *ctx, ckm_mapping *m, const char *const *ids, const char *residues, const uint64_t *offsets,
                    uint32_t n, char **text) {
    if (!text || !m) return CKM_EINVAL;
    *text = nullptr;
    std::vector<uint32_t> eids(n);
    std::map<uint32_t, size_t> matrix_proteins;  // matrix_request.cc:88-90 (a later duplicate overwrites the size)
    for (uint32_t i = 0; i < n; i++) {
        eids[i] = ckm_mapping_encode_id(m, ids[i]);
        matrix_proteins[eids[i]] = (size_t)(offsets[i + 1] - offsets[i]);
    }
    const ckm_pair_t *pairs = nullptr;
    uint64_t np = 0;
    int rc = ckm_matrix_rows(ctx, eids.data(), residues, offsets, n, 0, n, &pairs, &np);
    if (rc) return rc;
    std::vector<ckm_pair_t> v(pairs, pairs + np);
    v.resize(ckm_matrix_merge_pairs(v.data(), v.size()));
    ckm_text::Text os;
    for (const ckm_pair_t &p : v) {  // process_results, matrix_request.cc:171-184
        const size_t l1 = matrix_proteins[p.eid_i], l2 = matrix_proteins[p.eid_j];
        const float score = (float)p.count / ((float)(l1 + l2));
        os << ckm_mapping_decode_id(m, p.eid_i) << "\t" << ckm_mapping_decode_id(m, p.eid_j) << "\t" << p.count << "\t" << score
           << "\n";
    }
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

char *ckm_format_call(const ckm_ctx *ctx, const ckm_call_t *call) {
    ckm_text::Text os;
    put_call(os, ctx, *call);
    return os.dup();
}
char *ckm_format_hit(const ckm_ctx *ctx, const ckm_hit_t *hit) {
    ckm_text::Text os;
    put_hit(os, ctx, *hit);
    return os.dup();
}
char *ckm_format_otu_stats(const char *id, uint64_t seq_len, const ckm_otu_t *otus, uint64_t n_otus) {
    ckm_text::Text os;
    put_otu_stats(os, id, seq_len, otus, n_otus);
    return os.dup();
}
char *ckm_best_function(const ckm_ctx *ctx, const ckm_best_t *best) { return dup_text(best_function(ctx, *best)); }

void ckm_free_text(char *text) { free(text); }

}  // extern "C"
