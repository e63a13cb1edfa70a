// The worker-lambda body of QueryRequest (query_request.cc:103-152) written against the C++ mirror classes of
// include/ckm.hpp -- the loop a maintainer would have after switching engines.  Reads "<id>\t<seq>" lines, prints the
// /query response text.   usage: wrapper_check <kmer_dir> <details 0|1> <find_best_call 0|1> < work.tsv
#include <iostream>
#include <sstream>
#include <string>

#include "ckm.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int details = std::stoi(argv[2]), find_best_call = std::stoi(argv[3]);
    ckm::work_list_t cur;
    std::string line;
    while (std::getline(std::cin, line)) {
        const size_t tab = line.find('\t');
        if (tab == std::string::npos) continue;
        cur.emplace_back(line.substr(0, tab), line.substr(tab + 1));
    }
    try {
        ckm::KmerGuts kguts(argv[1], 0);
        kguts.set_parameters({});  // owner_->parameters(): none given -> defaults (kguts.cc:244-247)
        std::ostringstream os;
        unsigned want = ckm::KmerGuts::WANT_CALLS | ckm::KmerGuts::WANT_OTU;
        if (details) want |= ckm::KmerGuts::WANT_HITS;
        if (find_best_call) want |= ckm::KmerGuts::WANT_BEST;
        auto results = kguts.process_aa_seq(cur, want);
        for (size_t i = 0; i < cur.size(); i++) {
            const std::string &id = cur[i].first, &seq = cur[i].second;
            ckm::SeqResult &r = results[i];
            if (find_best_call) {
                if (!r.best_function.empty())
                    os << id << "\t" << r.best_function << "\t" << r.best_score << "\t" << r.best_weighted_score << "\n";
            } else {
                os << "PROTEIN-ID\t" << id << "\t" << seq.size() << "\n";
                for (auto c : r.calls) os << kguts.format_call(c);
                if (details)
                    for (auto h : r.hits) os << kguts.format_hit(h);
                os << kguts.format_otu_stats(id, seq.size(), r.otu_stats);
            }
        }
        // the one-call form must give the same bytes
        if (os.str() != kguts.query(cur, details, find_best_call)) {
            std::cerr << "wrapper loop and ckm_query_text disagree\n";
            return 3;
        }
        std::cout << os.str();
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
