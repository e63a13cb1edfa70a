// LookupRequest's three output modes, from the per-sequence score lists the GPU produced.
//
// The reference walks a std::unordered_map (seq_score_) and, for the listings, std::sort()s its entries by weighted_total
// only; where it leaves the order of equal elements to the container / sort implementation, the order here is ascending
// id.  Sums over that walk (the PGF roll-up) are f32 sums in ascending family id.
#include "lookup.h"
#include "text.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace ckm_lookup {

ckm_lookup_options_t options_from(const ckm_http::Request &r, const FamilyInfo &fams, bool family_mode) {
    ckm_lookup_options_t o;
    memset(&o, 0, sizeof o);
    o.family_mode = family_mode;
    auto take = [&r](const char *name, int &v) {
        try {
            v = std::stoi(r.param(name));
        } catch (const std::invalid_argument &) {
        } catch (const std::out_of_range &) {
        }
    };
    int v = 3;
    take("kmer_hit_threhsold", v);  // sic, lookup_request.cc:51
    o.kmer_hit_threshold = (unsigned int)v;
    v = 0;
    take("find_best_match", v);
    o.find_best_match = v != 0;
    v = 0;
    take("find_reps", v);
    o.find_reps = v != 0;
    v = 0;
    take("allow_ambiguous_functions", v);
    o.allow_ambiguous_functions = v != 0;
    auto g = fams.genus_map.find(r.param("target_genus"));  // lookup_genus: genus_map_[target_genus_]
    if (g != fams.genus_map.end() && !g->second.empty()) {
        try {
            o.target_genus_id = std::stoul(g->second);
        } catch (const std::exception &) {
        }
    }
    return o;
}

// find_best_match && family_mode, lookup_request.cc:201-327
static void best_match_line(ckm_text::Text &os, ckm_ctx *ctx, const ckm_family_data_t *fams, uint32_t n_fams, const ckm_lookup_options_t &o,
                            const char *id, const ckm_score_t *sc, uint64_t n, const ckm_best_t &best) {
    char *fn = ckm_best_function(ctx, &best);
    std::string best_call_function = fn ? fn : "";
    ckm_free_text(fn);
    std::string ambig_function;
    bool do_ambig_test = false;
    if (best_call_function.empty()) {
        best_call_function = "hypothetical protein";
    } else {
        const size_t where = best_call_function.find(" ?? ");
        if (where != std::string::npos) {
            if (o.allow_ambiguous_functions) {
                ambig_function = best_call_function.substr(where + 4);
                best_call_function = best_call_function.substr(0, where);
                do_ambig_test = true;
            } else {
                best_call_function = "hypothetical protein";
            }
        }
    }
    struct top_score {
        float score;
        std::string fam, function;
    };
    top_score best_lf{0.0f, "", ""}, best_gf{0.0f, "", ""};
    std::map<std::string, float> pgf_rollup, pgf_rollup_ambig;
    for (uint64_t k = 0; k < n; k++) {
        if (sc[k].hit_count < o.kmer_hit_threshold) continue;  // hit_total == hit_count
        if (sc[k].id >= n_fams) continue;
        const ckm_family_data_t &fd = fams[sc[k].id];
        if (best_call_function == fd.function) pgf_rollup[fd.pgf] += sc[k].weighted_total;
        else if (do_ambig_test && ambig_function == fd.function) pgf_rollup_ambig[fd.pgf] += sc[k].weighted_total;
        else continue;
        if (sc[k].weighted_total > best_lf.score && fd.genus_id == o.target_genus_id) {
            best_lf.score = sc[k].weighted_total;
            best_lf.fam = fd.plf;
            best_lf.function = fd.function;
        }
    }
    const std::map<std::string, float> *matching = &pgf_rollup;
    if (do_ambig_test && best_lf.function == ambig_function) matching = &pgf_rollup_ambig;
    for (const auto &e : *matching)
        if (e.second > best_gf.score) {
            best_gf.score = e.second;
            best_gf.fam = e.first;
        }
    os << id << "\t" << best_gf.fam << "\t" << best_gf.score << "\t" << best_lf.fam << "\t" << best_lf.score << "\t"
       << (do_ambig_test ? best_lf.function : best_call_function) << "\t" << best.score << "\t" << best.weighted_score << "\n";
}

}  // namespace ckm_lookup

extern "C" int ckm_lookup_text(ckm_ctx *ctx, ckm_mapping *pegs, const ckm_family_data_t *fams, uint32_t n_fams,
                               const ckm_lookup_options_t *opt, const char *const *ids, const char *residues, const uint64_t *offsets,
                               uint32_t n, char **text) {
    using namespace ckm_lookup;
    if (!text || !opt) return CKM_EINVAL;
    *text = nullptr;
    const ckm_lookup_options_t &o = *opt;
    ckm_text::Text os;
    if (o.family_mode) {
        ckm_family_scores_t fs;
        int rc = ckm_family_scores(ctx, residues, offsets, n, &fs);
        if (rc) return rc;
        std::vector<ckm_score_t> vec;
        for (uint32_t i = 0; i < n; i++) {
            const ckm_score_t *sc = fs.scores + fs.score_offsets[i];
            const uint64_t ns = fs.score_offsets[i + 1] - fs.score_offsets[i];
            if (o.find_best_match) {
                best_match_line(os, ctx, fams, n_fams, o, ids[i], sc, ns, fs.best[i]);
                continue;
            }
            // every family, best weighted_total first, up to the first one under the hit threshold (329-377)
            vec.assign(sc, sc + ns);
            std::stable_sort(vec.begin(), vec.end(), [](const ckm_score_t &l, const ckm_score_t &r) { return l.weighted_total > r.weighted_total; });
            os << ids[i] << "\n";
            for (const auto &e : vec) {
                if (e.hit_count < o.kmer_hit_threshold) break;
                static const ckm_family_data_t none = {"", "", "", 0, 0, 0};
                const ckm_family_data_t &fd = e.id < n_fams ? fams[e.id] : none;
                const float scaled = (float)e.hit_count / (float)fd.total_size;
                os << e.hit_count << "\t" << e.hit_count << "\t" << e.weighted_total << "\t" << fd.pgf << "\t" << fd.plf << "\t" << fd.total_size
                   << "\t" << fd.count << "\t" << scaled << "\t" << fd.function << "\n";
                if (o.find_reps) os << "///\n";  // family representatives (--family-reps) are not loaded: no rows
            }
            os << "//\n";
        }
    } else {
        // seq_score_[eid].hit_count++ only: hit_total stays 0, so nothing is listed unless the threshold is 0, and
        // weighted_total is 0 for everybody (the sort key): pegs come out by ascending id here
        const ckm_pair_t *pairs = nullptr;
        const uint64_t *poff = nullptr;
        if (o.kmer_hit_threshold == 0) {
            if (!pegs) return CKM_EINVAL;
            int rc = ckm_postings_scores(ctx, residues, offsets, n, &pairs, &poff);
            if (rc) return rc;
        }
        for (uint32_t i = 0; i < n; i++) {
            os << ids[i] << "\n";
            if (pairs)
                for (uint64_t k = poff[i]; k < poff[i + 1]; k++)
                    os << ckm_mapping_decode_id(pegs, pairs[k].eid_j) << "\t" << pairs[k].count << "\n";
            os << "//\n";
        }
    }
    *text = os.dup();
    return *text ? 0 : CKM_ENOMEM;
}

// FamilyMapper::find_all_matches (family_mapper.cc:207-285): family mode, no best match, kmer_hit_threshold_ = 3 (family_mapper.cc:7)
extern "C" int ckm_family_all_matches_text(ckm_ctx *ctx, const ckm_family_data_t *fams, uint32_t n_fams, const char *const *ids,
                                           const char *residues, const uint64_t *offsets, uint32_t n, char **text) {
    ckm_lookup_options_t o;
    memset(&o, 0, sizeof o);
    o.family_mode = 1;
    o.kmer_hit_threshold = 3;
    return ckm_lookup_text(ctx, nullptr, fams, n_fams, &o, ids, residues, offsets, n, text);
}
