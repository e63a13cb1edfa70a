// Test-infrastructure shim: kmer.h:16-23 (USE_TBB branch) container types mapped onto the std
// containers; the oracle driver is single-threaded per KmerPegMapping, so no concurrency is lost.
#pragma once
#include <unordered_map>
#include <atomic>
#include <utility>
#include <string>
namespace tbb {
template <class K, class V, class H = std::hash<K>> using concurrent_unordered_map = std::unordered_map<K, V, H>;
template <class T> struct atomic {
    std::atomic<T> v;
    atomic() : v(T()) {}
    atomic(T x) : v(x) {}
    T operator++(int) { return v.fetch_add(1); }
    T operator++() { return v.fetch_add(1) + 1; }
    operator T() const { return v.load(); }
    atomic& operator=(T x) { v.store(x); return *this; }
};
}
namespace std {
template <> struct hash<std::pair<std::string, std::string>> {
    size_t operator()(const std::pair<std::string, std::string>& p) const {
        return std::hash<std::string>()(p.first) * 1000003u ^ std::hash<std::string>()(p.second);
    }
};
}
