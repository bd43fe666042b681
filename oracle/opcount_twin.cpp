// opcount_twin.cpp — OPERATION-COUNTING build of the host twin (tests/twin/twin.cpp: the product's own
// __host__ __device__ arithmetic headers), measurement infrastructure like opcount.cpp.  `double` is replaced by a
// counting wrapper, so a step counts the primitive operations the IMPLEMENTED algorithm executes: add / sub, mul, fma,
// div, sqrt, compare (the elementary functions of csrc/ssa_math.h are built from these and are counted through them).
//   g++ -O1 -std=c++17 -fpermissive -w -ffp-contract=off -mfma -shared -fPIC -o oracle/libtwin_opcount.so oracle/opcount_twin.cpp
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_ADD, T_MUL, T_FMA, T_DIV, T_SQRT, T_CMP, T_MISC, T_N };
static uint64_t g_cnt[T_N];

struct cnt {
  double v;
  cnt() = default;
  constexpr cnt(double x) : v(x) {}
  constexpr cnt(int x) : v((double)x) {}
  constexpr cnt(long x) : v((double)x) {}
  constexpr cnt(unsigned long x) : v((double)x) {}
  constexpr cnt(unsigned x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0.0; }
  cnt& operator+=(cnt o) { ++g_cnt[T_ADD]; v += o.v; return *this; }
  cnt& operator-=(cnt o) { ++g_cnt[T_ADD]; v -= o.v; return *this; }
  cnt& operator*=(cnt o) { ++g_cnt[T_MUL]; v *= o.v; return *this; }
  cnt& operator/=(cnt o) { ++g_cnt[T_DIV]; v /= o.v; return *this; }
};
static_assert(sizeof(cnt) == sizeof(double) && alignof(cnt) == alignof(double), "layout-compatible with double");
#define CNT_BIN(op, slot)                                                                      \
  static inline cnt operator op(cnt a, cnt b) { ++g_cnt[slot]; return cnt(a.v op b.v); }       \
  static inline cnt operator op(cnt a, double b) { ++g_cnt[slot]; return cnt(a.v op b); }      \
  static inline cnt operator op(double a, cnt b) { ++g_cnt[slot]; return cnt(a op b.v); }      \
  static inline cnt operator op(cnt a, int b) { ++g_cnt[slot]; return cnt(a.v op b); }         \
  static inline cnt operator op(int a, cnt b) { ++g_cnt[slot]; return cnt(a op b.v); }
CNT_BIN(+, T_ADD) CNT_BIN(-, T_ADD) CNT_BIN(*, T_MUL) CNT_BIN(/, T_DIV)
#define CNT_CMP(op)                                                                    \
  static inline bool operator op(cnt a, cnt b) { ++g_cnt[T_CMP]; return a.v op b.v; }  \
  static inline bool operator op(cnt a, double b) { ++g_cnt[T_CMP]; return a.v op b; } \
  static inline bool operator op(double a, cnt b) { ++g_cnt[T_CMP]; return a op b.v; } \
  static inline bool operator op(cnt a, int b) { ++g_cnt[T_CMP]; return a.v op b; }    \
  static inline bool operator op(int a, cnt b) { ++g_cnt[T_CMP]; return a op b.v; }
CNT_CMP(<) CNT_CMP(>) CNT_CMP(<=) CNT_CMP(>=) CNT_CMP(==) CNT_CMP(!=)
static inline cnt operator-(cnt a) { return cnt(-a.v); }
static inline cnt operator+(cnt a) { return a; }
static inline cnt cnt_fma(cnt a, cnt b, cnt c) { ++g_cnt[T_FMA]; return cnt(__builtin_fma(a.v, b.v, c.v)); }
static inline cnt cnt_sqrt(cnt a) { ++g_cnt[T_SQRT]; return cnt(__builtin_sqrt(a.v)); }
static inline cnt fabs(cnt a) { return cnt(::fabs(a.v)); }
static inline cnt trunc(cnt a) { ++g_cnt[T_MISC]; return cnt(::trunc(a.v)); }
static inline cnt floor(cnt a) { ++g_cnt[T_MISC]; return cnt(::floor(a.v)); }
static inline cnt rint(cnt a) { ++g_cnt[T_MISC]; return cnt(::rint(a.v)); }
static inline cnt fmod(cnt a, cnt b) { ++g_cnt[T_MISC]; return cnt(::fmod(a.v, b.v)); }
static inline cnt sqrt(cnt a) { return cnt_sqrt(a); }
#define __builtin_fma(a, b, c) cnt_fma(a, b, c)
#define __builtin_sqrt(a) cnt_sqrt(a)

#define double cnt
#include "../tests/twin/twin.cpp"
#undef double
#undef __builtin_fma
#undef __builtin_sqrt
extern "C" {
void opcount_reset(void) { memset(g_cnt, 0, sizeof(g_cnt)); }
void opcount_get(uint64_t* out) { memcpy(out, g_cnt, sizeof(g_cnt)); }
int opcount_slots(void) { return T_N; }
}
