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
    std::vector<ckm_family_data_t> flat;         // `data` as the C ABI takes it (pointers into `data`)
    void flatten() {
        flat.clear();
        for (const auto &f : data) flat.push_back({f.pgf.c_str(), f.plf.c_str(), f.function.c_str(), f.genus_id, f.total_size, f.count});
    }
};

// LookupRequest's constructor, lookup_request.cc:36-80
ckm_lookup_options_t options_from(const ckm_http::Request &r, const FamilyInfo &fams, bool family_mode);

}  // namespace ckm_lookup
#endif
