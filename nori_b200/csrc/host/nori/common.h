// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// common.h -- host-side basics of the Nori mirror (namespace, exception, constants, string helpers).
// Mirrors the public surface of ref: include/nori/common.h (NoriException 135-140, Epsilon 38, constants 41-49,
// helpers 142-253) without Eigen / tinyformat, neither of which is available.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#define NORI_NAMESPACE_BEGIN namespace nori {
#define NORI_NAMESPACE_END }

#define Epsilon 1e-4f            /* ref: include/nori/common.h:38 */
#undef M_PI
#define M_PI 3.14159265358979323846f
#define INV_PI 0.31830988618379067154f
#define INV_TWOPI 0.15915494309189533577f

NORI_NAMESPACE_BEGIN

using std::cout;
using std::cerr;
using std::endl;

namespace detail {
inline const char *fmt_arg(const std::string &s) { return s.c_str(); }
template <typename T> inline T fmt_arg(const T &v) { return v; }
}

/// printf-style formatting into a std::string (std::string arguments allowed for %s)
inline std::string format(const char *fmt) { return std::string(fmt); }
template <typename... Args> std::string format(const char *fmt, const Args &...args) {
    int n = std::snprintf(nullptr, 0, fmt, detail::fmt_arg(args)...);
    std::string out((size_t) (n > 0 ? n : 0), '\0');
    if (n > 0) std::snprintf(&out[0], (size_t) n + 1, fmt, detail::fmt_arg(args)...);
    return out;
}

/// Simple exception class, which stores a human-readable error description (ref: include/nori/common.h:135-140)
class NoriException : public std::runtime_error {
public:
    template <typename... Args> NoriException(const char *fmt, const Args &...args)
        : std::runtime_error(format(fmt, args...)) { }
};

inline float degToRad(float value) { return value * (M_PI / 180.0f); }   // ref: include/nori/common.h:215

/// Indent a string by the specified number of spaces (ref: src/common.cpp:19-28)
std::string indent(const std::string &string, int amount = 2);
std::string toLower(const std::string &value);
bool toBool(const std::string &str);
int toInt(const std::string &str);
unsigned int toUInt(const std::string &str);
float toFloat(const std::string &str);
/// Tokenize a string into a list by splitting at 'delim' (ref: src/common.cpp:102-117)
std::vector<std::string> tokenize(const std::string &s, const std::string &delim = ", ", bool includeEmpty = false);
bool endsWith(const std::string &value, const std::string &ending);
std::string timeString(double time, bool precise = false);
std::string memString(size_t size, bool precise = false);

NORI_NAMESPACE_END
