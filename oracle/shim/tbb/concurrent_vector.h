// Test-infrastructure shim: see concurrent_unordered_map.h
#pragma once
#include <vector>
namespace tbb { template <class T> using concurrent_vector = std::vector<T>; }
