// Test-infrastructure shim: kmer.h:103 holds a boost::mutex member that the oracle driver never locks.
#pragma once
#include <mutex>
namespace boost { using mutex = std::mutex; template <class M> using lock_guard = std::lock_guard<M>; }
