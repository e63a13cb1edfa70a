// Test-infrastructure shim: global.h:14 declares a boost::timer::cpu_timer that the hot path never reads.
#pragma once
namespace boost { namespace timer { struct cpu_timer { void start() {} }; } }
