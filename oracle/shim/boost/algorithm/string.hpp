// Test-infrastructure shim: the one Boost.StringAlgo call on the path, dna_seq.cc:17
//   boost::split(list, str, boost::is_any_of("*"), boost::token_compress_on)
// Semantics restated from Boost's documented split_iterator behaviour: the input is cut at every
// maximal run of separator characters (token_compress_on merges adjacent separators); a separator
// run at the very start or end still yields one empty leading / trailing token; an input with no
// separator yields the whole string; an empty input yields one empty token.
#pragma once
#include <string>
namespace boost {
namespace algorithm { enum token_compress_mode_type { token_compress_on, token_compress_off }; }
using algorithm::token_compress_on;
using algorithm::token_compress_off;
struct is_any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string& s) { return is_any_of_pred{s}; }
template <class Seq, class Pred>
Seq& split(Seq& out, const std::string& in, Pred pred, algorithm::token_compress_mode_type mode = token_compress_off) {
    out.clear();
    size_t i = 0, n = in.size();
    std::string cur;
    for (;;) {
        size_t j = i;
        while (j < n && !pred(in[j])) ++j;
        out.push_back(in.substr(i, j - i));
        if (j >= n) break;
        ++j;                                   // consume one separator
        if (mode == token_compress_on) while (j < n && pred(in[j])) ++j;
        i = j;
        if (i >= n) { out.push_back(std::string()); break; }
    }
    return out;
}
}
