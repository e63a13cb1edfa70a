// ckm.hpp -- header-only C++ mirror of the reference's engine classes over the C ABI (ckm.h / ckm_handlers.h).
//
// Same names, argument meaning and results as the reference, with ONE difference: the unit of work is the list of
// sequences parsed from a body chunk (the `work_list_t` the handlers already hold, query_request.h:24) instead of one
// sequence, because that is what a GPU call amortises.
//
//   reference                                                   here
//   KmerGuts(kmer_dir, image)                 kguts.h:312       ckm::KmerGuts(kmer_dir, device)
//   set_parameters(map<string,string>)        kguts.cc:244      KmerGuts::set_parameters (same reset-then-apply rule)
//   process_aa_seq[_hits](id, seq, calls, hits, otu)  :879-908  KmerGuts::process_aa_seq(work, want, results)
//   find_best_call(calls, fI, function, score, weighted, offset) :1008   SeqResult::{best_*} (WANT_BEST)
//   format_call / format_hit / format_otu_stats  :939-973       same names
//   function_at_index, encoded_aa_kmer, decoded_kmer            same names
//   FamilyMapper(kguts, mapping).find_best_family_match(id,seq) family_mapper.cc:65   FamilyMapper::find_best_family_match(work)
//   DNASequence(id,seq).get_possible_proteins(trans_table)      dna_seq.cc:9          KmerGuts::get_possible_proteins(reads)
//
// Errors: the reference prints and exit()s; these classes throw std::runtime_error carrying ckm_last_error().
#ifndef CKM_HPP
#define CKM_HPP

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ckm.h"
#include "ckm_handlers.h"

namespace ckm {

typedef ckm_call_t KmerCall;                // kguts.h:166-183
typedef ckm_hit_t hit_in_sequence_t;        // kguts.h:228-233 (slot copy + offset, flattened)
typedef std::pair<std::string, std::string> ProteinSequence;  // (id, seq), prot_seq.h
typedef std::vector<ProteinSequence> work_list_t;

struct KmerOtuStats {  // kguts.h:185-219
    std::string contig_id;
    int contig_len = 0;
    std::vector<std::pair<int, int>> otu_map;  // (otu index, count) in ascending otu index, like std::map iteration
};

struct SeqResult {
    std::vector<KmerCall> calls;
    std::vector<hit_in_sequence_t> hits;
    KmerOtuStats otu_stats;
    // find_best_call outputs (kguts.cc:1008-1199)
    int best_function_index = -1;
    std::string best_function;
    float best_score = 0, best_weighted_score = 0, best_score_offset = 0;
};

inline void check(int rc) {
    if (rc != CKM_OK) throw std::runtime_error(std::string("libckm: ") + ckm_last_error());
}

// concatenated residues + CSR offsets of a work list
struct Flat {
    std::string residues;
    std::vector<uint64_t> offsets;
    std::vector<const char *> ids;
    explicit Flat(const work_list_t &work) {
        offsets.reserve(work.size() + 1);
        offsets.push_back(0);
        for (const auto &w : work) {
            ids.push_back(w.first.c_str());
            residues += w.second;
            offsets.push_back(residues.size());
        }
    }
};

class KmerGuts {
public:
    enum { WANT_CALLS = CKM_WANT_CALLS, WANT_HITS = CKM_WANT_HITS, WANT_OTU = CKM_WANT_OTU, WANT_BEST = CKM_WANT_BEST };

    explicit KmerGuts(const std::string &kmer_dir, int device = 0) { check(ckm_open(kmer_dir.c_str(), device, &ctx_)); }
    ~KmerGuts() { ckm_close(ctx_); }
    KmerGuts(const KmerGuts &) = delete;
    KmerGuts &operator=(const KmerGuts &) = delete;
    ckm_ctx *ctx() const { return ctx_; }

    void set_default_parameters() { ckm_set_default_params(ctx_); }

    // kguts.cc:244-268: reset to defaults, then apply the integer-valued engine parameters; bad integers are ignored
    void set_parameters(const std::map<std::string, std::string> &params) {
        int v[4] = {0, 5, 0, 200};
        static const char *names[4] = {"order_constraint", "min_hits", "min_weighted_hits", "max_gap"};
        for (int k = 0; k < 4; k++) {
            auto it = params.find(names[k]);
            if (it == params.end()) continue;
            try {
                v[k] = std::stoi(it->second);
            } catch (const std::invalid_argument &) {
            } catch (const std::out_of_range &) {
            }
        }
        check(ckm_set_params(ctx_, v[0], v[1], v[2], v[3]));
    }

    // process_aa_seq / process_aa_seq_hits (+ find_best_call) for every (id, seq) of the chunk; `want` says which of
    // calls / hits / otu_stats / best the caller would have passed as non-null
    std::vector<SeqResult> process_aa_seq(const work_list_t &work, unsigned want) {
        Flat f(work);
        ckm_batch_out_t o;
        check(ckm_call_batch(ctx_, f.residues.data(), f.offsets.data(), (uint32_t)work.size(), want, &o));
        std::vector<SeqResult> res(work.size());
        for (size_t i = 0; i < work.size(); i++) {
            SeqResult &r = res[i];
            if (want & WANT_CALLS) r.calls.assign(o.calls + o.call_offsets[i], o.calls + o.call_offsets[i + 1]);
            if (want & WANT_HITS) r.hits.assign(o.hits + o.hit_offsets[i], o.hits + o.hit_offsets[i + 1]);
            if (want & WANT_OTU) {
                r.otu_stats.contig_id = work[i].first;
                r.otu_stats.contig_len = (int)work[i].second.size();
                for (uint64_t k = o.otu_offsets[i]; k < o.otu_offsets[i + 1]; k++)
                    r.otu_stats.otu_map.emplace_back(o.otus[k].otu_index, o.otus[k].count);
            }
            if (want & WANT_BEST) {
                const ckm_best_t &b = o.best[i];
                r.best_function_index = b.function_index;
                char *fn = ckm_best_function(ctx_, &b);
                r.best_function = fn;
                ckm_free_text(fn);
                r.best_score = b.score;
                r.best_weighted_score = b.weighted_score;
                r.best_score_offset = b.score_offset;
            }
        }
        return res;
    }

    const char *function_at_index(int i) const { return ckm_function_at_index(ctx_, i); }  // kguts.h:361-366
    static unsigned long long encoded_aa_kmer(const char *p) { return ckm_encoded_aa_kmer(p); }
    static void decoded_kmer(unsigned long long k, char *decoded) { ckm_decoded_kmer(k, decoded); }

    std::string format_call(const KmerCall &c) const { return take(ckm_format_call(ctx_, &c)); }
    std::string format_hit(const hit_in_sequence_t &h) const { return take(ckm_format_hit(ctx_, &h)); }
    std::string format_otu_stats(const std::string &id, size_t size, const KmerOtuStats &s) const {
        std::vector<ckm_otu_t> v;
        for (const auto &e : s.otu_map) v.push_back({e.first, e.second});
        return take(ckm_format_otu_stats(id.c_str(), size, v.data(), v.size()));
    }

    // DNASequence::get_possible_proteins for every read: per read, six (frame, fragments) entries in the order
    // {1,2,3,-1,-2,-3}; only fragments longer than min_len are materialised (the fq path uses 10)
    typedef std::vector<std::pair<int, std::vector<std::string>>> frames_t;
    std::vector<frames_t> get_possible_proteins(const work_list_t &reads, unsigned min_len = 0) {
        Flat f(reads);
        ckm_fq_fragments_t o;
        check(ckm_fq_translate(ctx_, f.residues.data(), f.offsets.data(), (uint32_t)reads.size(), min_len, &o));
        static const int frames[6] = {1, 2, 3, -1, -2, -3};
        std::vector<frames_t> out(reads.size());
        for (size_t r = 0; r < reads.size(); r++)
            for (int s = 0; s < 6; s++) {
                std::vector<std::string> frags;
                for (uint64_t k = o.frag_frame_offsets[6 * r + s]; k < o.frag_frame_offsets[6 * r + s + 1]; k++)
                    frags.emplace_back(o.residues + o.frag_offsets[k], o.residues + o.frag_offsets[k + 1]);
                out[r].emplace_back(frames[s], std::move(frags));
            }
        return out;
    }

    // response text of the handlers (query_request.cc:103-152, fq_process_request.cc:298-365)
    std::string query(const work_list_t &work, bool details, bool find_best_call) {
        Flat f(work);
        char *t = nullptr;
        check(ckm_query_text(ctx_, f.ids.data(), f.residues.data(), f.offsets.data(), (uint32_t)work.size(), details, find_best_call, &t));
        return take(t);
    }
    std::string fq_lookup(const work_list_t &reads) {
        Flat f(reads);
        char *t = nullptr;
        check(ckm_fq_text(ctx_, f.ids.data(), f.residues.data(), f.offsets.data(), (uint32_t)reads.size(), &t));
        return take(t);
    }

private:
    static std::string take(char *t) {
        std::string s = t ? t : "";
        ckm_free_text(t);
        return s;
    }
    ckm_ctx *ctx_ = nullptr;
};

// FamilyMapper (family_mapper.h:14-68) over the family side tables of a KmerPegMapping
class FamilyMapper {
public:
    struct best_match_t {  // family_mapper.h:20-28
        std::string gfam_id;
        float gfam_score;
        std::string lfam_id;
        float lfam_score;
        std::string function;
        float score;
    };

    // kmer_to_family_id_ as CSR + family_data_ as three string arrays (see ckm_family_load)
    FamilyMapper(KmerGuts *kguts, const std::vector<uint64_t> &kmers, const std::vector<uint64_t> &fam_offsets,
                 const std::vector<uint32_t> &fam_ids, const std::vector<std::string> &pgf, const std::vector<std::string> &plf,
                 const std::vector<std::string> &function)
        : kguts_(kguts) {
        std::vector<const char *> a, b, c;
        for (size_t i = 0; i < pgf.size(); i++) {
            a.push_back(pgf[i].c_str());
            b.push_back(plf[i].c_str());
            c.push_back(function[i].c_str());
        }
        check(ckm_family_load(kguts->ctx(), kmers.size(), kmers.data(), fam_offsets.data(), fam_ids.data(), (uint32_t)pgf.size(),
                              a.data(), b.data(), c.data()));
    }

    std::vector<best_match_t> find_best_family_match(const work_list_t &work) {
        Flat f(work);
        const ckm_family_match_t *m = nullptr;
        check(ckm_family_batch(kguts_->ctx(), f.residues.data(), f.offsets.data(), (uint32_t)work.size(), &m));
        std::vector<best_match_t> out;
        for (size_t i = 0; i < work.size(); i++)
            out.push_back({ckm_family_pgf_name(kguts_->ctx(), m[i].gfam), m[i].gfam_score, ckm_family_plf_name(kguts_->ctx(), m[i].lfam),
                           m[i].lfam_score, ckm_family_function_name(kguts_->ctx(), &m[i]), m[i].score});
        return out;
    }

    // FamilyMapper::find_all_matches(os, id, seq) (family_mapper.cc:207-285) over a work list; `families` is the reference's
    // family_data_ (kmer.h:118-127) as the C ABI takes it
    void find_all_matches(std::ostream &os, const work_list_t &work, const std::vector<ckm_family_data_t> &families) {
        Flat f(work);
        std::vector<const char *> ids;
        for (const auto &w : work) ids.push_back(w.first.c_str());
        char *text = nullptr;
        check(ckm_family_all_matches_text(kguts_->ctx(), families.data(), (uint32_t)families.size(), ids.data(), f.residues.data(),
                                          f.offsets.data(), (uint32_t)work.size(), &text));
        os << text;
        ckm_free_text(text);
    }

private:
    KmerGuts *kguts_;
};

}  // namespace ckm
#endif
