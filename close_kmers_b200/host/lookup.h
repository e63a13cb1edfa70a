// POST /lookup: LookupRequest (lookup_request.cc) above ckm_family_scores / ckm_postings_scores.
#ifndef CKM_HOST_LOOKUP_H
#define CKM_HOST_LOOKUP_H
#include <map>
#include <string>
#include <vector>

#include "../../include/ckm.h"
#include "../../include/ckm_handlers.h"
#include "../../include/ckm_server.h"
#include "http.h"

namespace ckm_lookup {

struct FamilyData {  // KmerPegMapping::family_data_t, kmer.h:58-68
    std::string pgf, plf, function;
    unsigned long genus_id;
    unsigned long total_size;
    unsigned short count;
};

struct FamilyInfo {
    std::vector<FamilyData> data;                // family_data_, by encoded family id
    std::map<std::string, std::string> genus_map;  // genus_map_
};

struct Options {  // LookupRequest's members, lookup_request.cc:36-80
    bool family_mode = false;
    unsigned int kmer_hit_threshold = 3;
    bool find_best_match = false, find_reps = false, allow_ambiguous_functions = false;
    unsigned long target_genus_id = 0;
};

Options options_from(const ckm_http::Request &r, const FamilyInfo &fams, bool family_mode);

// the response text of one batch (lookup_request.cc:155-400), appended to `out`
int lookup_text(ckm_ctx *ctx, ckm_mapping *pegs, const FamilyInfo &fams, const Options &o, const ckm_seq_batch_t &b, std::string &out);

}  // namespace ckm_lookup
#endif
