// np2_error.h — the one exception type that crosses the library's internal layers; the C ABI maps it to
// an NP2_ERR_* code + np2_last_error().
#pragma once
#include <stdexcept>
#include <string>

namespace np2 {
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
}  // namespace np2
