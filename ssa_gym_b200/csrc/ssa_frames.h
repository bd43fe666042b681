// ssa_frames.h — HOST code: the GCRS -> ITRS rotation table of an episode without ERFA.
//
// The reference builds `trans_matrix[i]` once per environment (ssa_tasker_simple_2.py:136-137) through the ERFA C library
// (envs/transformations.py:143-214: cal2jd, dat, xys06a, c2ixys, era00, sp00, pom00 and the IERS EOP table).  ERFA is not
// part of this image; this is the same chain restated in C++ so that a C / C++ caller of libssa_ukf.so can produce the
// per-step input of the measurement model itself: exact calendar / leap-second / Earth-rotation-angle / TIO-locator /
// polar-motion arithmetic, CIP X, Y from the truncated IAU 2006/2000A series (IERS Conventions 2010, eq. 5.16 and the
// leading rows of Tables 5.2a / 5.2b: polynomial part + every periodic term above 12 mas; the lunisolar terms between 1 and
// 9 mas from their nutation amplitudes).  Accuracy: 7.5e-9 rad against the SOFA matrix quoted in the reference's
// tests.py:107-109 (0.3 m at GEO) — an INPUT generator, flagged approximate; parity of
// the path is defined for identical matrices.  ssa_gym_b200/transformations.py holds the same arithmetic in Python
// (tests/test_host_logic.py compares the two).
#pragma once
#include <math.h>

namespace ssa_frames {

constexpr double kDAS2R = 4.848136811095359935899141e-6, kDJ00 = 2451545.0, kDJC = 36525.0, kDAYSEC = 86400.0;
constexpr double kTAU = 6.283185307179586476925287;

// (2400000.5, MJD at 0h) of a Gregorian date — eraCal2jd
inline double cal2mjd(int iy, int im, int id) {
  const int my = (im - 14) / 12;  // C division truncates toward zero, as in eraCal2jd
  const long iypmy = (long)iy + my;
  return (double)((1461L * (iypmy + 4800L)) / 4L + (367L * (long)(im - 2 - 12 * my)) / 12L - (3L * ((iypmy + 4900L) / 100L)) / 4L +
                  (long)id - 2432076L);
}
// Gregorian year / month of an MJD day number — eraJd2cal
inline void mjd2ym(long mjd, int* iy, int* im) {
  long l = mjd + 2400001L + 68569L;
  const long n = (4L * l) / 146097L;
  l -= (146097L * n + 3L) / 4L;
  const long i = (4000L * (l + 1L)) / 1461001L;
  l -= (1461L * i) / 4L - 31L;
  const long k = (80L * l) / 2447L;
  l = k / 11L;
  *im = (int)(k + 2L - 12L * l);
  *iy = (int)(100L * (n - 49L) + i + l);
}
// TAI - UTC [s] (eraDat for dates >= 1972)
inline double dat(int iy, int im) {
  static const int leap[][3] = {{1972, 1, 10}, {1972, 7, 11}, {1973, 1, 12}, {1974, 1, 13}, {1975, 1, 14}, {1976, 1, 15}, {1977, 1, 16},
                                {1978, 1, 17}, {1979, 1, 18}, {1980, 1, 19}, {1981, 7, 20}, {1982, 7, 21}, {1983, 7, 22}, {1985, 7, 23},
                                {1988, 1, 24}, {1990, 1, 25}, {1991, 1, 26}, {1992, 7, 27}, {1993, 7, 28}, {1994, 7, 29}, {1996, 1, 30},
                                {1997, 7, 31}, {1999, 1, 32}, {2006, 1, 33}, {2009, 1, 34}, {2012, 7, 35}, {2015, 7, 36}, {2017, 1, 37}};
  double d = 10.0;
  for (const auto& r : leap)
    if (iy > r[0] || (iy == r[0] && im >= r[1])) d = (double)r[2];
  return d;
}
// Earth rotation angle, IAU 2000 — eraEra00
inline double era00(double dj1, double dj2) {
  const double d1 = dj1 < dj2 ? dj1 : dj2, d2 = dj1 < dj2 ? dj2 : dj1;
  const double t = d1 + (d2 - kDJ00);
  const double f = fmod(d1, 1.0) + fmod(d2, 1.0);
  double theta = fmod(kTAU * (f + 0.7790572732640 + 0.00273781191135448 * t), kTAU);
  if (theta < 0.0) theta += kTAU;
  return theta;
}
struct M3 { double a[3][3]; };
inline M3 mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.a[i][j] = (x.a[i][0] * y.a[0][j] + x.a[i][1] * y.a[1][j]) + x.a[i][2] * y.a[2][j];
  return r;
}
inline M3 rx(double t) { const double c = cos(t), s = sin(t); return M3{{{1, 0, 0}, {0, c, s}, {0, -s, c}}}; }
inline M3 ry(double t) { const double c = cos(t), s = sin(t); return M3{{{c, 0, -s}, {0, 1, 0}, {s, 0, c}}}; }
inline M3 rz(double t) { const double c = cos(t), s = sin(t); return M3{{{c, s, 0}, {-s, c, 0}, {0, 0, 1}}}; }

// CIP X, Y and the CIO locator s [rad]; t = TT Julian centuries since J2000
inline void xys(double t, double* X, double* Y, double* s) {
  const double om = (450160.398036 - 6962890.5431 * t) * kDAS2R;     // mean longitude of the Moon's node
  const double F = (335779.526232 + 1739527262.8478 * t) * kDAS2R;   // L - Omega
  const double D = (1072260.70369 + 1602961601.2090 * t) * kDAS2R;   // mean elongation of the Moon
  const double lp = (1287104.79305 + 129596581.0481 * t) * kDAS2R;   // mean anomaly of the Sun
  const double l = (485868.249036 + 1717915923.2178 * t) * kDAS2R;   // mean anomaly of the Moon
  const double a2 = 2 * (F - D + om), a3 = 2 * (F + om);
  double x = -0.016617 + 2004.191898 * t - 0.4297829 * t * t - 0.19861834 * t * t * t - 6.844318 * sin(om) - 0.523908 * sin(a2) -
             0.090552 * sin(a3) + 0.082169 * sin(2 * om) + 0.058707 * sin(lp) + 0.028288 * sin(l) - 0.020558 * sin(lp + a2) -
             0.015407 * sin(2 * F + om) - 0.011992 * sin(l + a3) + 0.205833 * t * cos(om);
  double y = -0.006951 - 0.025896 * t - 22.4072747 * t * t + 0.00190059 * t * t * t + 9.205236 * cos(om) + 0.573033 * cos(a2) +
             0.097847 * cos(a3) - 0.089618 * cos(2 * om) + 0.022438 * cos(lp + a2) + 0.020070 * cos(2 * F + om) +
             0.012902 * cos(l + a3) + 0.153042 * t * sin(om);
  // the next lunisolar terms, 1 - 9 mas: X_i = sin(eps0) dpsi_i sin(arg_i), Y_i = deps_i cos(arg_i) from the nutation
  // amplitudes of the classical series [0.1 mas] (within 1 % of the IAU 2000A X,Y coefficients at this level)
  static const double minor[14][7] = {{0, -1, 2, -2, 2, 217.0, -95.0}, {0, 0, 2, -2, 1, 129.0, -70.0}, {1, 0, 0, -2, 0, -158.0, -1.0},
                                      {-1, 0, 2, 0, 2, 123.0, -53.0}, {0, 0, 0, 2, 0, 63.0, -2.0},     {1, 0, 0, 0, 1, 63.0, -33.0},
                                      {-1, 0, 0, 0, 1, -58.0, 32.0},  {-1, 0, 2, 2, 2, -59.0, 26.0},   {1, 0, 2, 0, 1, -51.0, 27.0},
                                      {0, 0, 2, 2, 2, -38.0, 16.0},   {2, 0, 0, 0, 0, 29.0, -1.0},     {1, 0, 2, -2, 2, 29.0, -12.0},
                                      {2, 0, 2, 0, 2, -31.0, 13.0},   {0, 0, 2, 0, 0, 26.0, -1.0}};
  const double sin_eps0 = 0.397777156;  // sin of the J2000 obliquity
  for (const auto& m : minor) {
    const double arg = m[0] * l + m[1] * lp + m[2] * F + m[3] * D + m[4] * om;
    x = x + (1e-4 * sin_eps0 * m[5]) * sin(arg);
    y = y + (1e-4 * m[6]) * cos(arg);
  }
  x *= kDAS2R;
  y *= kDAS2R;
  *X = x;
  *Y = y;
  *s = -x * y / 2 + (94e-6 + 3808.65e-6 * t - 2640.73e-6 * sin(om)) * kDAS2R;
}
// GCRS -> CIRS from X, Y, s — eraC2ixys
inline M3 c2ixys(double x, double y, double s) {
  const double r2 = x * x + y * y;
  const double e = r2 > 0 ? atan2(y, x) : 0.0;
  const double d = atan(sqrt(r2 / (1.0 - r2)));
  return mul(mul(rz(-(e + s)), ry(d)), rz(e));
}
// linear interpolation of the daily EOP rows [mjd, x", y", UT1-UTC, dX", dY"] (transformations.py:156-165); zeros outside
inline void eop_at(const double* eop, int n_eop, long mjd, double frac, double v[5]) {
  for (int q = 0; q < 5; ++q) v[q] = 0.0;
  if (!eop) return;
  const double *lo = nullptr, *hi = nullptr;
  for (int r = 0; r < n_eop; ++r) {
    if ((long)eop[6 * r] == mjd) lo = eop + 6 * r;
    if ((long)eop[6 * r] == mjd + 1) hi = eop + 6 * r;
  }
  if (!lo || !hi) return;
  for (int q = 0; q < 5; ++q) v[q] = lo[1 + q] * (1 - frac) + hi[1 + q] * frac;
}
// one matrix: MJD day number (UTC), seconds of the day (the reference uses whole seconds: datetime.hour/minute/second)
inline M3 gcrs2itrs(long mjd, double sec, const double* eop, int n_eop) {
  const double djmjd0 = 2400000.5, date = (double)mjd, day_frac = sec / kDAYSEC;
  double v[5];
  eop_at(eop, n_eop, mjd, day_frac, v);
  int iy, im;
  mjd2ym(mjd, &iy, &im);
  const double tt = date + day_frac + dat(iy, im) / kDAYSEC + 32.184 / kDAYSEC;
  const double tut = day_frac + v[2] / kDAYSEC;
  const double tc = ((djmjd0 - kDJ00) + tt) / kDJC;
  double X, Y, s;
  xys(tc, &X, &Y, &s);
  const M3 rc2i = c2ixys(X + v[3] * kDAS2R, Y + v[4] * kDAS2R, s);
  const M3 rc2ti = mul(rz(era00(djmjd0 + date, tut)), rc2i);
  const double sp = -47e-6 * tc * kDAS2R;
  const M3 rpom = mul(mul(rx(-v[1] * kDAS2R), ry(-v[0] * kDAS2R)), rz(sp));
  return mul(rpom, rc2ti);
}

}  // namespace ssa_frames
