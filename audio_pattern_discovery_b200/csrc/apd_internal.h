// apd_internal.h -- glue between the translation units of libapd_b200 (not part of the ABI).
#pragma once
#include <string>

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace apd {
// APD_DEBUG=1: host-side phase timings on stderr ("[apd] <what>: x.xx ms").
inline bool debug_enabled()
{
    static const bool on = std::getenv("APD_DEBUG") != nullptr;
    return on;
}
struct PhaseTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what)
    {
        if (!debug_enabled()) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[apd] %s: %.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Message returned by apd_last_error(NULL) on the calling thread (entry points without a context).
void set_thread_error(const std::string& msg);
}  // namespace apd
