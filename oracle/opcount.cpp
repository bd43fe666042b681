// opcount.cpp — OPERATION-COUNTING build of the CPU oracle (test / measurement infrastructure, not product code).
//
// SURVEY.md 8(d) prices the reference's literal sequence by convention (add / sub / mul / compare = 1, div = sqrt = 10,
// sin = cos = 40, tan = 70, atan = 60, atan2 = 80, asin = acos = 70, Python-mod = 10) from an ESTIMATED operation count
// and asks for the estimate to be replaced by a measured one.  This file compiles oracle/ukf_oracle.c — the reference's
// operation order, unchanged — with `double` replaced by a counting wrapper, so that running a step counts every
// arithmetic operation and library call the reference's sequence performs.  Same numerical results as liboracle.so.
//   g++ -O1 -std=c++17 -fpermissive -w -shared -fPIC -o oracle/liboracle_opcount.so oracle/opcount.cpp
//   python tools/opcount.py            (drives it on the C2 inputs and prints / stores the weighted count per unit)
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

enum { C_ADD, C_MUL, C_DIV, C_CMP, C_SQRT, C_SIN, C_COS, C_TAN, C_ATAN, C_ATAN2, C_ASIN, C_ACOS, C_MOD, C_HYP, C_EXPLOG, C_ABS, C_N };
static uint64_t g_cnt[C_N];

struct cnt {
  double v;
  cnt() = default;
  cnt(double x) : v(x) {}
  cnt(int x) : v((double)x) {}
  cnt(long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0.0; }
  cnt& operator+=(cnt o) { ++g_cnt[C_ADD]; v += o.v; return *this; }
  cnt& operator-=(cnt o) { ++g_cnt[C_ADD]; v -= o.v; return *this; }
  cnt& operator*=(cnt o) { ++g_cnt[C_MUL]; v *= o.v; return *this; }
  cnt& operator/=(cnt o) { ++g_cnt[C_DIV]; v /= o.v; return *this; }
};
static_assert(sizeof(cnt) == sizeof(double) && alignof(cnt) == alignof(double), "layout-compatible with double");
#define CNT_BIN(op, slot)                                                                      \
  static inline cnt operator op(cnt a, cnt b) { ++g_cnt[slot]; return cnt(a.v op b.v); }       \
  static inline cnt operator op(cnt a, double b) { ++g_cnt[slot]; return cnt(a.v op b); }      \
  static inline cnt operator op(double a, cnt b) { ++g_cnt[slot]; return cnt(a op b.v); }      \
  static inline cnt operator op(cnt a, int b) { ++g_cnt[slot]; return cnt(a.v op b); }         \
  static inline cnt operator op(int a, cnt b) { ++g_cnt[slot]; return cnt(a op b.v); }
CNT_BIN(+, C_ADD) CNT_BIN(-, C_ADD) CNT_BIN(*, C_MUL) CNT_BIN(/, C_DIV)
#define CNT_CMP(op)                                                                    \
  static inline bool operator op(cnt a, cnt b) { ++g_cnt[C_CMP]; return a.v op b.v; }  \
  static inline bool operator op(cnt a, double b) { ++g_cnt[C_CMP]; return a.v op b; } \
  static inline bool operator op(double a, cnt b) { ++g_cnt[C_CMP]; return a op b.v; } \
  static inline bool operator op(cnt a, int b) { ++g_cnt[C_CMP]; return a.v op b; }    \
  static inline bool operator op(int a, cnt b) { ++g_cnt[C_CMP]; return a op b.v; }
CNT_CMP(<) CNT_CMP(>) CNT_CMP(<=) CNT_CMP(>=) CNT_CMP(==) CNT_CMP(!=)
static inline cnt operator-(cnt a) { return cnt(-a.v); }  // sign flip: not an arithmetic operation
static inline cnt operator+(cnt a) { return a; }
#define CNT_F1(name, slot) static inline cnt name(cnt a) { ++g_cnt[slot]; return cnt(::name(a.v)); }
CNT_F1(sqrt, C_SQRT) CNT_F1(sin, C_SIN) CNT_F1(cos, C_COS) CNT_F1(tan, C_TAN) CNT_F1(atan, C_ATAN) CNT_F1(asin, C_ASIN) CNT_F1(acos, C_ACOS)
CNT_F1(sinh, C_HYP) CNT_F1(cosh, C_HYP) CNT_F1(tanh, C_HYP) CNT_F1(asinh, C_HYP) CNT_F1(acosh, C_HYP) CNT_F1(atanh, C_HYP)
CNT_F1(exp, C_EXPLOG) CNT_F1(log, C_EXPLOG) CNT_F1(cbrt, C_EXPLOG)
CNT_F1(fabs, C_ABS) CNT_F1(floor, C_ABS) CNT_F1(trunc, C_ABS)
static inline cnt atan2(cnt a, cnt b) { ++g_cnt[C_ATAN2]; return cnt(::atan2(a.v, b.v)); }
static inline cnt fmod(cnt a, cnt b) { ++g_cnt[C_MOD]; return cnt(::fmod(a.v, b.v)); }
static inline cnt fmod(cnt a, double b) { ++g_cnt[C_MOD]; return cnt(::fmod(a.v, b)); }
static inline cnt pow(cnt a, cnt b) { ++g_cnt[C_EXPLOG]; return cnt(::pow(a.v, b.v)); }
static inline cnt pow(cnt a, double b) { ++g_cnt[C_EXPLOG]; return cnt(::pow(a.v, b)); }
static inline cnt pow(double a, cnt b) { ++g_cnt[C_EXPLOG]; return cnt(::pow(a, b.v)); }
static inline cnt pow(double a, int b) { return cnt(::pow(a, (double)b)); }  // constant table entries (10^i)
static inline int isnan_cnt(cnt a) { return a.v != a.v; }
static inline int isfinite_cnt(cnt a) { return isfinite(a.v); }
#undef isnan
#undef isfinite
#undef isinf
#define isnan(x) isnan_cnt(x)
#define isfinite(x) isfinite_cnt(x)

extern "C" {
#define double cnt
#include "ukf_oracle.c"
#undef double
void opcount_reset(void) { memset(g_cnt, 0, sizeof(g_cnt)); }
void opcount_get(uint64_t* out) { memcpy(out, g_cnt, sizeof(g_cnt)); }
int opcount_slots(void) { return C_N; }
}
