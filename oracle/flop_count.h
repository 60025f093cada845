/* oracle/flop_count.h — TEST / MEASUREMENT INFRASTRUCTURE: instruments the CPU restatement for SURVEY.md §8(d).
 *
 * `g++ -include flop_count.h -DHBO_COUNT_FLOPS hb_oracle.cpp` (make -C oracle count) replaces every `double` of
 * hb_oracle.cpp by a same-size wrapper whose arithmetic counts itself: additions/subtractions, multiplications,
 * divisions, comparisons (incl. min/max/fabs tests), sqrt-class calls (sqrt, pow(x, 0.5f)), pow, exp/sin. The
 * arithmetic is unchanged (results stay bit-identical); hbo_flop_counts() returns the totals. tools/count_flops.py
 * turns them into FP64 operations per HRU*timestep per solver and structure: the algorithmic figure of the roofline.
 */
#ifndef HBO_FLOP_COUNT_H
#define HBO_FLOP_COUNT_H
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

typedef double hbo_raw;
struct hbo_counters {
    unsigned long long add, mul, div, cmp, abs, pow, exp;
};
extern hbo_counters g_hbo_counters;

struct hbo_counted {
    hbo_raw v;
    hbo_counted() = default;
    hbo_counted(hbo_raw x) : v(x) {}
    hbo_counted(float x) : v(x) {}
    hbo_counted(int x) : v(x) {}
    hbo_counted(long long x) : v((hbo_raw)x) {}
    hbo_counted(unsigned x) : v(x) {}
    hbo_counted(size_t x) : v((hbo_raw)x) {}
    explicit operator float() const { return (float)v; }
    explicit operator int() const { return (int)v; }
    explicit operator long long() const { return (long long)v; }
    explicit operator bool() const { return v != 0; }
    hbo_counted& operator+=(hbo_counted o) { ++g_hbo_counters.add; v += o.v; return *this; }
    hbo_counted& operator-=(hbo_counted o) { ++g_hbo_counters.add; v -= o.v; return *this; }
    hbo_counted& operator*=(hbo_counted o) { ++g_hbo_counters.mul; v *= o.v; return *this; }
    hbo_counted& operator/=(hbo_counted o) { ++g_hbo_counters.div; v /= o.v; return *this; }
    hbo_counted operator-() const { return hbo_counted(-v); }
};
static_assert(sizeof(hbo_counted) == sizeof(hbo_raw), "layout");
#define HBO_BIN(op, ctr)                                                                                              \
    inline hbo_counted operator op(hbo_counted a, hbo_counted b) { ++g_hbo_counters.ctr; return hbo_counted(a.v op b.v); } \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                                         \
    inline hbo_counted operator op(hbo_counted a, T b) { ++g_hbo_counters.ctr; return hbo_counted(a.v op (hbo_raw)b); }  \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                                         \
    inline hbo_counted operator op(T a, hbo_counted b) { ++g_hbo_counters.ctr; return hbo_counted((hbo_raw)a op b.v); }
HBO_BIN(+, add)
HBO_BIN(-, add)
HBO_BIN(*, mul)
HBO_BIN(/, div)
#undef HBO_BIN
#define HBO_CMP(op)                                                                                     \
    inline bool operator op(hbo_counted a, hbo_counted b) { ++g_hbo_counters.cmp; return a.v op b.v; } \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                           \
    inline bool operator op(hbo_counted a, T b) { ++g_hbo_counters.cmp; return a.v op (hbo_raw)b; }    \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                           \
    inline bool operator op(T a, hbo_counted b) { ++g_hbo_counters.cmp; return (hbo_raw)a op b.v; }
HBO_CMP(<)
HBO_CMP(>)
HBO_CMP(<=)
HBO_CMP(>=)
HBO_CMP(==)
HBO_CMP(!=)
#undef HBO_CMP
namespace std {
inline hbo_counted fabs(hbo_counted a) { ++g_hbo_counters.abs; return hbo_counted(::fabs(a.v)); }
inline hbo_counted abs(hbo_counted a) { ++g_hbo_counters.abs; return hbo_counted(::fabs(a.v)); }
inline hbo_counted sqrt(hbo_counted a) { ++g_hbo_counters.pow; return hbo_counted(::sqrt(a.v)); }
inline hbo_counted exp(hbo_counted a) { ++g_hbo_counters.exp; return hbo_counted(::exp(a.v)); }
inline hbo_counted sin(hbo_counted a) { ++g_hbo_counters.exp; return hbo_counted(::sin(a.v)); }
inline hbo_counted floor(hbo_counted a) { return hbo_counted(::floor(a.v)); }
template <class T>
inline hbo_counted pow(hbo_counted a, T b) { ++g_hbo_counters.pow; return hbo_counted(::pow(a.v, (hbo_raw)b)); }
inline hbo_counted pow(hbo_counted a, hbo_counted b) { ++g_hbo_counters.pow; return hbo_counted(::pow(a.v, b.v)); }
inline bool isnan(hbo_counted a) { return std::isnan(a.v); }
inline bool isinf(hbo_counted a) { return std::isinf(a.v); }
inline bool isfinite(hbo_counted a) { return std::isfinite(a.v); }
inline hbo_counted max(hbo_counted a, hbo_raw b) { return std::max(a, hbo_counted(b)); }
inline hbo_counted max(hbo_raw a, hbo_counted b) { return std::max(hbo_counted(a), b); }
inline hbo_counted min(hbo_counted a, hbo_raw b) { return std::min(a, hbo_counted(b)); }
inline hbo_counted min(hbo_raw a, hbo_counted b) { return std::min(hbo_counted(a), b); }
}  // namespace std
using std::pow;
using std::sqrt;
using std::fabs;
#define double hbo_counted
#endif
