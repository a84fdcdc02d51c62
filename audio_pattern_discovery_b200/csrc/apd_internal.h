// apd_internal.h -- glue between the translation units of libapd_b200 (not part of the ABI).
#pragma once
#include <string>

namespace apd {
// Message returned by apd_last_error(NULL) on the calling thread (entry points without a context).
void set_thread_error(const std::string& msg);
}  // namespace apd
