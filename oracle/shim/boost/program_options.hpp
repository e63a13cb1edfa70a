// Test-infrastructure shim (NOT product code): the minimum of boost::program_options
// that /root/reference/kmer_image.cc:70-72 touches when the reference TUs are compiled
// unchanged for oracle/_ref.  count() -> 0 means "option absent" (no MAP_POPULATE).
#pragma once
#include <cstddef>
#include <string>
namespace boost { namespace program_options {
struct variable_value { template <class T> T as() const { return T(); } };
class variables_map {
public:
    size_t count(const std::string&) const { return 0; }
    variable_value operator[](const std::string&) const { return variable_value(); }
};
} }
