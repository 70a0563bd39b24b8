// Error policy of the header mirror.  The reference asserts (and vanishes under NDEBUG); the mirror
// turns every non-OK status of the C ABI into std::runtime_error so that a missing GPU or a failed
// kernel is never silent -- there is no CPU fallback behind these headers.
#ifndef PBSO_CHECK_H
#define PBSO_CHECK_H
#include <stdexcept>
#include <string>
#include "../pbso_b200.h"
namespace pbso_mirror {
inline void check(int rc, const char* where) {
    if (rc == PBSO_OK) return;
    std::string msg = std::string(where) + ": " + pbso_last_error();
    if (rc == PBSO_ERR_RANGE) throw std::out_of_range(msg);
    throw std::runtime_error(msg);
}
}  // namespace pbso_mirror
#endif
