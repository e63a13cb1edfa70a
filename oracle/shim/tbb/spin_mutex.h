// Test-infrastructure shim: see concurrent_unordered_map.h
#pragma once
#include <mutex>
namespace tbb { using spin_mutex = std::mutex; }
