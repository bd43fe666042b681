/* ukf_oracle.c — CPU ORACLE (test infrastructure, not product code).
 *
 * A plain-C, libm-based restatement of the reference's per-step UKF hot path, written to follow the
 * reference's own operation order (no fused multiply-add, no hoisting, divisions where the reference
 * divides).  It is the checker for the CUDA path and the CPU baseline timed by bench.py; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (ssa_gym_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pinning: this file is checked against golden vectors produced by the reference's own
 * functions (envs/farnocchia.py and envs/transformations.py imported from /root/reference, plus a
 * numpy/scipy restatement of filterpy — see tests/golden/make_golden.py and tests/test_oracle_golden.py).
 * filterpy itself (requirements.txt:14, unpinned, 1.4.5 contemporary) is absent from the reference tree
 * and from this image, so the UKF algebra is restated from its published algorithm (SURVEY.md
 * Appendix B); no reference test asserts outputs of the default AER path, hence at the filterpy boundary
 * parity is "pinned to the reference's call sites and to reference-generated vectors", not to a
 * reference-run episode.
 *
 * Each function cites the reference file:line it follows (paths relative to the upstream repo).
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ssa_ukf.h" /* only for the ssa_ukf_cfg struct and flag/status constants */

#define MU 398600441800000.0
#define PI 3.141592653589793 /* numpy.pi */
#define NSIG 13

typedef struct { int exc; } octx; /* exc != 0: the reference would have raised inside numba */

/* Python float '%' as numba lowers it (sign follows the divisor). */
static double pymod(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0 && ((b < 0.0) != (m < 0.0))) m += b;
  return m;
}
static double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double* a) { return sqrt(dot3(a, a)); }

/* ---- envs/farnocchia.py:356-753 anomaly conversions ------------------------------------------ */
static double nu_to_E(double nu, double ecc) { return 2 * atan(sqrt((1 - ecc) / (1 + ecc)) * tan(nu / 2)); } /* :467 */
static double E_to_nu(double E, double ecc) { return 2 * atan(sqrt((1 + ecc) / (1 - ecc)) * tan(E / 2)); }    /* :539 */
static double nu_to_F(double nu, double ecc) { return 2 * atanh(sqrt((ecc - 1) / (ecc + 1)) * tan(nu / 2)); } /* :503 */
static double F_to_nu(double F, double ecc) { return 2 * atan(sqrt((ecc + 1) / (ecc - 1)) * tanh(F / 2)); }   /* :568 */
static double E_to_M(double E, double ecc) { return E - ecc * sin(E); }                                       /* :687 */
static double F_to_M(double F, double ecc) { return ecc * sinh(F) - F; }                                      /* :719 */
static double D_to_M(double D) { return D + D * D * D / 3; }                                                  /* :752 */
static double M_to_D(double M) {                                                                               /* :649-652 */
  double B = 3.0 * M / 2.0;
  double A = pow(B + sqrt(1.0 + B * B), 2.0 / 3.0);
  return 2 * A * B / (1 + A + A * A);
}
/* :336-353 newton(), elliptic and hyperbolic flavours */
static double newton_elliptic(double x0, double M, double ecc) {
  double p0 = 1.0 * x0;
  for (int i = 0; i < 50; ++i) {
    double fval = E_to_M(p0, ecc) - M;
    double fder = 1 - ecc * cos(p0);
    double p = p0 - fval / fder;
    if (fabs(p - p0) < 1.48e-08) return p;
    p0 = p;
  }
  return NAN;
}
static double newton_hyperbolic(double x0, double M, double ecc) {
  double p0 = 1.0 * x0;
  for (int i = 0; i < 100; ++i) {
    double fval = F_to_M(p0, ecc) - M;
    double fder = ecc * cosh(p0) - 1;
    double p = p0 - fval / fder;
    if (fabs(p - p0) < 1.48e-08) return p;
    p0 = p;
  }
  return NAN;
}
static double M_to_E(double M, double ecc, octx* c) { /* :572-601 */
  if (!(-PI <= M && M <= PI)) { c->exc = 1; return NAN; }
  double E0;
  if (ecc < 0.8) E0 = M;
  else E0 = PI * ((M > 0) - (M < 0));
  return newton_elliptic(E0, M, ecc);
}
static double M_to_F(double M, double ecc) { return newton_hyperbolic(asinh(M / ecc), M, ecc); } /* :625-627 */

/* :769-798 S_x and dS_x_alt */
static double series_S(double ecc, double x, int deriv, octx* c) {
  if (!(fabs(x) < 1)) { c->exc = 1; return NAN; }
  double S = 0, xk = 1.0;
  for (int k = 0; k < 200000; ++k) {
    double S_old = S;
    double term = ecc - 1.0 / (2 * k + 3);
    if (deriv) term = term * (2 * k + 3);
    S += term * xk;
    xk *= x;
    if (fabs(S - S_old) < 1e-12) return S;
  }
  c->exc = 1; /* the reference spins forever here */
  return NAN;
}
static double D_to_M_near_parabolic(double D, double ecc, octx* c) { /* :801-808 */
  double x = (ecc - 1.0) / (ecc + 1.0) * (D * D);
  double S = series_S(ecc, x, 0, c);
  double ope = 1.0 + ecc;
  return sqrt(2.0 / ope) * D + sqrt(2.0 / (ope * ope * ope)) * (D * D * D) * S;
}
static double M_to_D_near_parabolic(double M, double ecc, octx* c) { /* :811-843 */
  double D0 = M_to_D(M);
  for (int it = 0; it < 50; ++it) {
    double fval = D_to_M_near_parabolic(D0, ecc, c) - M;
    double x = (ecc - 1.0) / (ecc + 1.0) * (D0 * D0);
    double dS = series_S(ecc, x, 1, c);
    double ope = 1.0 + ecc;
    double fder = sqrt(2.0 / ope) + sqrt(2.0 / (ope * ope * ope)) * (D0 * D0) * dS;
    if (c->exc) return NAN;
    double D = D0 - fval / fder;
    if (fabs(D - D0) < 1.48e-08) return D;
    D0 = D;
  }
  return NAN;
}

/* envs/farnocchia.py:846-921 */
static double delta_t_from_nu(double nu, double ecc, double k, double q, octx* c) {
  const double delta = 1e-2;
  double M, n;
  if (!(-PI <= nu && nu < PI)) { c->exc = 1; return NAN; }
  if (ecc < 1 - delta) {
    double E = nu_to_E(nu, ecc);
    M = E_to_M(E, ecc);
    n = sqrt(k * ((1 - ecc) * (1 - ecc) * (1 - ecc)) / (q * q * q));
  } else if (1 - delta <= ecc && ecc < 1) {
    double E = nu_to_E(nu, ecc);
    if (delta <= 1 - ecc * cos(E)) {
      M = E_to_M(E, ecc);
      n = sqrt(k * ((1 - ecc) * (1 - ecc) * (1 - ecc)) / (q * q * q));
    } else {
      double D = tan(nu / 2.0);
      M = D_to_M_near_parabolic(D, ecc, c);
      n = sqrt(k / (2 * (q * q * q)));
    }
  } else if (ecc == 1) {
    double D = tan(nu / 2.0);
    M = D_to_M(D);
    n = sqrt(k / (2 * (q * q * q)));
  } else if (1 + ecc * cos(nu) < 0) {
    return NAN;
  } else if (1 < ecc && ecc <= 1 + delta) {
    double F = nu_to_F(nu, ecc);
    if (delta <= ecc * cosh(F) - 1) {
      M = F_to_M(F, ecc);
      n = sqrt(k * ((ecc - 1) * (ecc - 1) * (ecc - 1)) / (q * q * q));
    } else {
      double D = tan(nu / 2.0);
      M = D_to_M_near_parabolic(D, ecc, c);
      n = sqrt(k / (2 * (q * q * q)));
    }
  } else if (1 + delta < ecc) {
    double F = nu_to_F(nu, ecc);
    M = F_to_M(F, ecc);
    n = sqrt(k * ((ecc - 1) * (ecc - 1) * (ecc - 1)) / (q * q * q));
  } else {
    c->exc = 1; /* RuntimeError */
    return NAN;
  }
  if (n == 0.0) { c->exc = 1; return NAN; }
  return M / n;
}

/* envs/farnocchia.py:924-1006 */
static double nu_from_delta_t(double delta_t, double ecc, double k, double q, octx* c) {
  const double delta = 1e-2;
  double nu;
  if (ecc < 1 - delta) {
    double n = sqrt(k * ((1 - ecc) * (1 - ecc) * (1 - ecc)) / (q * q * q));
    double M = n * delta_t;
    double E = M_to_E(pymod(M + PI, 2 * PI) - PI, ecc, c);
    nu = E_to_nu(E, ecc);
  } else if (1 - delta <= ecc && ecc < 1) {
    double E_delta = acos((1 - delta) / ecc);
    double n = sqrt(k * ((1 - ecc) * (1 - ecc) * (1 - ecc)) / (q * q * q));
    double M = n * delta_t;
    if (E_to_M(E_delta, ecc) <= fabs(M)) {
      double E = M_to_E(pymod(M + PI, 2 * PI) - PI, ecc, c);
      nu = E_to_nu(E, ecc);
    } else {
      n = sqrt(k / (2 * (q * q * q)));
      M = n * delta_t;
      nu = 2.0 * atan(M_to_D_near_parabolic(M, ecc, c));
    }
  } else if (ecc == 1) {
    double n = sqrt(k / (2 * (q * q * q)));
    nu = 2.0 * atan(M_to_D(n * delta_t));
  } else if (1 < ecc && ecc <= 1 + delta) {
    double F_delta = acosh((1 + delta) / ecc);
    double n = sqrt(k * ((ecc - 1) * (ecc - 1) * (ecc - 1)) / (q * q * q));
    double M = n * delta_t;
    if (F_to_M(F_delta, ecc) <= fabs(M)) {
      nu = F_to_nu(M_to_F(M, ecc), ecc);
    } else {
      n = sqrt(k / (2 * (q * q * q)));
      M = n * delta_t;
      nu = 2.0 * atan(M_to_D_near_parabolic(M, ecc, c));
    }
  } else {
    double n = sqrt(k * ((ecc - 1) * (ecc - 1) * (ecc - 1)) / (q * q * q));
    double M = n * delta_t;
    nu = F_to_nu(M_to_F(M, ecc), ecc);
  }
  return nu;
}

/* envs/farnocchia.py:164-313 */
static void rv2coe(double k, const double* r, const double* v, double* coe, octx* c) {
  const double tol = 1e-8;
  double h[3], n[3], e[3], tmp[3];
  const double kz[3] = {0, 0, 1};
  cross3(r, v, h);
  cross3(kz, h, n);
  double rn = norm3(r), hn = norm3(h);
  if (rn == 0.0 || hn == 0.0) { c->exc = 1; for (int i = 0; i < 6; ++i) coe[i] = NAN; return; }
  double c1 = dot3(v, v) - k / rn, rv = dot3(r, v);
  for (int i = 0; i < 3; ++i) e[i] = (c1 * r[i] - rv * v[i]) / k;
  double ecc = norm3(e);
  double p = dot3(h, h) / k;
  double inc = acos(h[2] / hn);
  int circular = ecc < tol, equatorial = fabs(inc) < tol;
  double raan, argp, nu;
  if (equatorial && !circular) {
    raan = 0;
    argp = pymod(atan2(e[1], e[0]), 2 * PI);
    cross3(e, r, tmp);
    nu = atan2(dot3(h, tmp) / hn, dot3(r, e));
  } else if (!equatorial && circular) {
    raan = pymod(atan2(n[1], n[0]), 2 * PI);
    argp = 0;
    cross3(h, n, tmp);
    nu = atan2(dot3(r, tmp) / hn, dot3(r, n));
  } else if (equatorial && circular) {
    raan = 0;
    argp = 0;
    nu = pymod(atan2(r[1], r[0]), 2 * PI);
  } else {
    double ome2 = 1 - ecc * ecc;
    if (ome2 == 0.0) { c->exc = 1; for (int i = 0; i < 6; ++i) coe[i] = NAN; return; }
    double a = p / ome2;
    double ka = k * a;
    if (a > 0) {
      double e_se = rv / sqrt(ka);
      double e_ce = rn * dot3(v, v) / k - 1;
      nu = E_to_nu(atan2(e_se, e_ce), ecc);
    } else {
      double e_sh = rv / sqrt(-ka);
      double vn = norm3(v);
      double e_ch = rn * (vn * vn) / k - 1;
      if (e_ch - e_sh == 0.0) { c->exc = 1; for (int i = 0; i < 6; ++i) coe[i] = NAN; return; }
      nu = F_to_nu(log((e_ch + e_sh) / (e_ch - e_sh)) / 2, ecc);
    }
    raan = pymod(atan2(n[1], n[0]), 2 * PI);
    double px = dot3(r, n);
    cross3(h, n, tmp);
    double py = dot3(r, tmp) / hn;
    argp = pymod(atan2(py, px) - nu, 2 * PI);
  }
  nu = pymod(nu + PI, 2 * PI) - PI;
  coe[0] = p; coe[1] = ecc; coe[2] = inc; coe[3] = raan; coe[4] = argp; coe[5] = nu;
}

static void matmul3(const double a[3][3], const double b[3][3], double c[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i][j] = (a[i][0] * b[0][j] + a[i][1] * b[1][j]) + a[i][2] * b[2][j];
}
/* envs/farnocchia.py:100-161 (rv_pqw :14-73, rotation_matrix :76-87, coe_rotation_matrix :90-97) */
static void coe2rv(double k, const double* coe, double* out) {
  double p = coe[0], ecc = coe[1], inc = coe[2], raan = coe[3], argp = coe[4], nu = coe[5];
  double pqw[2][3] = {{cos(nu), sin(nu), 0}, {-sin(nu), ecc + cos(nu), 0}};
  double s0 = p / (1 + ecc * cos(nu)), s1 = sqrt(k / p);
  for (int j = 0; j < 3; ++j) { pqw[0][j] *= s0; pqw[1][j] *= s1; }
  double cO = cos(raan), sO = sin(raan), ci = cos(inc), si = sin(inc), cw = cos(argp), sw = sin(argp);
  double R3a[3][3] = {{cO, -sO, 0.0}, {sO, cO, 0.0}, {0.0, 0.0, 1.0}};
  double R1[3][3] = {{1.0, 0.0, 0.0}, {0.0, ci, -si}, {0.0, si, ci}};
  double R3b[3][3] = {{cw, -sw, 0.0}, {sw, cw, 0.0}, {0.0, 0.0, 1.0}};
  double t[3][3], rm[3][3];
  matmul3(R3a, R1, t);
  matmul3(t, R3b, rm);
  for (int a = 0; a < 2; ++a)
    for (int i = 0; i < 3; ++i)
      out[3 * a + i] = (pqw[a][0] * rm[i][0] + pqw[a][1] * rm[i][1]) + pqw[a][2] * rm[i][2];
}

/* envs/farnocchia.py:1009-1062 */
static int fx_farnocchia(const double* x, double dt, double* out) {
  octx c = {0};
  double coe[6];
  rv2coe(MU, x, x + 3, coe, &c);
  if (c.exc) { for (int i = 0; i < 6; ++i) out[i] = NAN; return 1; }
  double q = coe[0] / (1 + coe[1]);
  double delta_t0 = delta_t_from_nu(coe[5], coe[1], MU, q, &c);
  double delta_t = delta_t0 + dt;
  coe[5] = nu_from_delta_t(delta_t, coe[1], MU, q, &c);
  coe2rv(MU, coe, out);
  return c.exc;
}

/* ---- measurement side ---------------------------------------------------------------------- */
/* envs/dynamics.py:219-231 hx_aer_erfa + envs/transformations.py:329-352 ecef2aer.
 * T is trans_uvw_ecef (transformations.py:341-343), precomputed by the host with numpy. */
static void hx_aer(const double* x, const double* M, const double* obs_itrs, const double* T, double* aer) {
  double xi[3], d[3], e[3];
  for (int i = 0; i < 3; ++i) xi[i] = (M[3 * i] * x[0] + M[3 * i + 1] * x[1]) + M[3 * i + 2] * x[2];
  for (int i = 0; i < 3; ++i) d[i] = xi[i] - obs_itrs[i];
  for (int i = 0; i < 3; ++i) e[i] = (T[i] * d[0] + T[3 + i] * d[1]) + T[6 + i] * d[2];
  double r = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  double az = atan2(e[1], e[0]);
  if (az < 0) az = az + 2 * PI;
  aer[0] = az; aer[1] = asin(e[2] / r); aer[2] = r;
}
static void aer2uvw(const double* aer, double* uvw) { /* transformations.py:283-297 */
  uvw[0] = aer[2] * cos(aer[1]) * cos(aer[0]);
  uvw[1] = aer[2] * cos(aer[1]) * sin(aer[0]);
  uvw[2] = aer[2] * sin(aer[1]);
}
static void uvw2aer(const double* uvw, double* aer) { /* transformations.py:300-316 */
  double r = sqrt((uvw[0] * uvw[0] + uvw[1] * uvw[1]) + uvw[2] * uvw[2]);
  double az = atan2(uvw[1], uvw[0]);
  if (az < 0) az = az + 2 * PI;
  aer[0] = az; aer[1] = asin(uvw[2] / r); aer[2] = r;
}
static void residual_aer(const double* a, const double* b, double* c) { /* dynamics.py:260-267 */
  c[0] = atan2(sin(a[0] - b[0]), cos(a[0] - b[0]));
  c[1] = a[1] - b[1];
  c[2] = a[2] - b[2];
}

/* ---- filterpy algebra ------------------------------------------------- */
/* scipy.linalg.cholesky(a) (upper, reads the upper triangle; LAPACK dpotf2 'U' order). 0 = raised. */
static int cholesky_upper(const double A[6][6], double U[6][6]) {
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j)
      if (!isfinite(A[i][j])) return 0; /* check_finite -> ValueError */
  memset(U, 0, 36 * sizeof(double));
  for (int j = 0; j < 6; ++j) {
    double ajj = A[j][j];
    for (int k = 0; k < j; ++k) ajj -= U[k][j] * U[k][j];
    if (!(ajj > 0.0)) return 0; /* dpotrf info > 0 -> LinAlgError */
    ajj = sqrt(ajj);
    U[j][j] = ajj;
    double inv = 1.0 / ajj;
    for (int c = j + 1; c < 6; ++c) {
      double s = A[j][c];
      for (int k = 0; k < j; ++k) s -= U[k][j] * U[k][c];
      U[j][c] = s * inv;
    }
  }
  return 1;
}
/* envs/dynamics.py:402-417 robust_cholesky.  Returns attempt number, -1 if LinAlgError. */
static int robust_cholesky(const double A[6][6], double U[6][6]) {
  static const double p10[16] = {1e-06, 1e-05, 1e-04, 1e-03, 1e-02, 1e-01, 1.0, 10.0,
                                 1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8, 1e9};
  if (cholesky_upper(A, U)) return 0;
  for (int t = 0; t < 16; ++t) {
    double B[6][6];
    memcpy(B, A, sizeof(B));
    for (int i = 0; i < 6; ++i) B[i][i] = A[i][i] + p10[t];
    if (cholesky_upper(B, U)) return t + 1;
  }
  return -1;
}
/* MerweScaledSigmaPoints.sigma_points */
static int sigma_points(const double* x, const double P[6][6], double lam, double sig[NSIG][6]) {
  double A[6][6], U[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i][j] = lam * P[i][j];
  int r = robust_cholesky(A, U);
  if (r < 0) return r;
  for (int j = 0; j < 6; ++j) sig[0][j] = x[j];
  for (int k = 0; k < 6; ++k)
    for (int j = 0; j < 6; ++j) { sig[k + 1][j] = x[j] - (-U[k][j]); sig[6 + k + 1][j] = x[j] - U[k][j]; }
  return r;
}
/* numpy.linalg.inv of a 3x3 (LAPACK dgesv: partial-pivot LU, reciprocal column scaling, trsm solves) */
static int inv3(const double S[3][3], double SI[3][3]) {
  double a[3][3], b[3][3];
  int ok = 1;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = S[i][j]; b[i][j] = (i == j); }
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int i = c + 1; i < 3; ++i) if (fabs(a[i][c]) > fabs(a[p][c])) p = i;
    if (p != c) for (int j = 0; j < 3; ++j) {
      double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t;
      t = b[c][j]; b[c][j] = b[p][j]; b[p][j] = t;
    }
    if (a[c][c] == 0.0) ok = 0;
    double rp = 1.0 / a[c][c];
    for (int i = c + 1; i < 3; ++i) {
      a[i][c] *= rp;
      for (int j = c + 1; j < 3; ++j) a[i][j] -= a[i][c] * a[c][j];
    }
  }
  for (int col = 0; col < 3; ++col) {
    double y0 = b[0][col];
    double y1 = b[1][col] - a[1][0] * y0;
    double y2 = (b[2][col] - a[2][0] * y0) - a[2][1] * y1;
    double x2 = y2 / a[2][2];
    double x1 = (y1 - a[1][2] * x2) / a[1][1];
    double x0 = ((y0 - a[0][2] * x2) - a[0][1] * x1) / a[0][0];
    SI[0][col] = x0; SI[1][col] = x1; SI[2][col] = x2;
  }
  return ok;
}

typedef struct {
  double* x_true; double* x; double* P; int32_t* status; int32_t* infl; const double* z_noise;
  double* obs; double* dpos; double* dvel; double* spos; double* svel; double* trace;
  double* z_true; double* y; double* S; double* sigmas_h; uint8_t* visible; uint8_t* updated;
} obj_io;

static void fail_object(obj_io* o, int code) { /* SS2:369-382, sentinels :157-158 */
  for (int i = 0; i < 6; ++i) o->x[i] = i < 3 ? 1e20 : 1e12;
  memset(o->P, 0, 36 * sizeof(double));
  for (int i = 0; i < 6; ++i) o->P[7 * i] = i < 3 ? 1e20 : 1e12;
  *o->status |= SSA_STATUS_FAILED | code;
}

static void object_step(const ssa_ukf_cfg* cfg, const double* M, int flags, int tasked, obj_io* o) {
  double sig[NSIG][6];
  int have_sig = 0;
  double(*P)[6] = (double(*)[6])o->P;

  if (flags & SSA_STEP_TRUTH) { /* SS2:265-266 */
    double xt[6];
    if (fx_farnocchia(o->x_true, cfg->dt, xt)) *o->status |= SSA_STATUS_TRUTHEXC;
    memcpy(o->x_true, xt, sizeof(xt));
  }
  if ((flags & SSA_STEP_PREDICT) && !(*o->status & SSA_STATUS_FAILED)) { /* SS2:271-287, UKF.predict */
    int r = sigma_points(o->x, P, cfg->lam_plus_n, sig);
    if (r < 0) {
      fail_object(o, SSA_STATUS_LINALG);
    } else {
      if (r > 0) *o->infl += 1;
      double f[NSIG][6];
      int exc = 0;
      for (int k = 0; k < NSIG; ++k) exc |= fx_farnocchia(sig[k], cfg->dt, f[k]);
      if (exc) {
        fail_object(o, SSA_STATUS_FXEXC);
      } else {
        /* unscented_transform, residual = np.subtract fast path */
        double xb[6], yv[NSIG][6], wy[NSIG][6];
        for (int i = 0; i < 6; ++i) {
          double acc = 0.0;
          for (int k = 0; k < NSIG; ++k) acc += cfg->Wm[k] * f[k][i];
          xb[i] = acc;
        }
        for (int k = 0; k < NSIG; ++k) for (int i = 0; i < 6; ++i) { yv[k][i] = f[k][i] - xb[i]; wy[k][i] = cfg->Wc[k] * yv[k][i]; }
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < NSIG; ++k) acc += yv[k][i] * wy[k][j];
            P[i][j] = acc + cfg->Q[6 * i + j];
          }
        int nan = 0;
        for (int i = 0; i < 6; ++i) { o->x[i] = xb[i]; nan |= isnan(xb[i]); }
        int code = nan ? SSA_STATUS_NAN : 0;
        if (cfg->resample_after_predict) {
          r = sigma_points(o->x, P, cfg->lam_plus_n, sig);
          if (r < 0) code |= SSA_STATUS_LINALG;
          else if (r > 0) *o->infl += 1;
        } else {
          memcpy(sig, f, sizeof(f));
        }
        if (code) fail_object(o, code);
        else have_sig = 1;
      }
    }
  }

  int want_upd = (flags & SSA_STEP_UPDATE_ALL) || ((flags & SSA_STEP_UPDATE_ACT) && tasked);
  int want_meas = want_upd || (flags & SSA_STEP_EPILOGUE);
  double zt_aer[3] = {0, 0, 0};
  int visible = 0;
  if (want_meas) { /* SS2:418-425 object_visible: elevation of the TRUE state */
    hx_aer(o->x_true, M, cfg->obs_itrs, cfg->T, zt_aer);
    visible = zt_aer[1] >= cfg->obs_limit;
    if (o->visible) *o->visible = (uint8_t)visible;
  }
  if (o->updated) *o->updated = 0;

  if (want_upd && !(*o->status & SSA_STATUS_FAILED)) { /* SS2:292-315, UKF.update */
    double zt[3];
    for (int a = 0; a < 3; ++a) zt[a] = (cfg->obs_type == SSA_OBS_AER) ? zt_aer[a] : o->x_true[a];
    if (o->z_true) memcpy(o->z_true, zt, sizeof(zt));
    if (visible) {
      int ok_sig = 1;
      if (!have_sig) {
        int r = sigma_points(o->x, P, cfg->lam_plus_n, sig);
        if (r < 0) { fail_object(o, SSA_STATUS_LINALG | SSA_STATUS_IN_UPDATE); ok_sig = 0; }
      }
      if (ok_sig) {
        double z[3], zs[NSIG][3], rz[NSIG][3], zp[3], S[3][3];
        for (int a = 0; a < 3; ++a) z[a] = zt[a] + (o->z_noise ? o->z_noise[a] : 0.0);
        if (cfg->obs_type == SSA_OBS_AER) {
          double uvw[NSIG][3], zm[3];
          for (int k = 0; k < NSIG; ++k) { hx_aer(sig[k], M, cfg->obs_itrs, cfg->T, zs[k]); aer2uvw(zs[k], uvw[k]); }
          for (int a = 0; a < 3; ++a) { /* dynamics.py:353 np.dot(Wm, aers) */
            double acc = 0.0;
            for (int k = 0; k < NSIG; ++k) acc += cfg->Wm[k] * uvw[k][a];
            zm[a] = acc;
          }
          uvw2aer(zm, zp);
          memset(S, 0, sizeof(S));
          for (int k = 0; k < NSIG; ++k) { /* P += Wc[k] * outer(y, y) */
            residual_aer(zs[k], zp, rz[k]);
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] += cfg->Wc[k] * (rz[k][a] * rz[k][b]);
          }
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] += cfg->R[3 * a + b];
        } else { /* hx_xyz / np.dot mean / np.subtract residual: fast path */
          for (int k = 0; k < NSIG; ++k) for (int a = 0; a < 3; ++a) zs[k][a] = sig[k][a];
          for (int a = 0; a < 3; ++a) {
            double acc = 0.0;
            for (int k = 0; k < NSIG; ++k) acc += cfg->Wm[k] * zs[k][a];
            zp[a] = acc;
          }
          for (int k = 0; k < NSIG; ++k) for (int a = 0; a < 3; ++a) rz[k][a] = zs[k][a] - zp[a];
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
              double acc = 0.0;
              for (int k = 0; k < NSIG; ++k) acc += rz[k][a] * (cfg->Wc[k] * rz[k][b]);
              S[a][b] = acc + cfg->R[3 * a + b];
            }
        }
        double SI[3][3], Pxz[6][3], K[6][3], T[3][6], yr[3];
        int ok = inv3(S, SI);
        memset(Pxz, 0, sizeof(Pxz));
        for (int k = 0; k < NSIG; ++k) { /* cross_variance */
          double dx[6];
          for (int i = 0; i < 6; ++i) dx[i] = sig[k][i] - o->x[i];
          for (int i = 0; i < 6; ++i) for (int a = 0; a < 3; ++a) Pxz[i][a] += cfg->Wc[k] * (dx[i] * rz[k][a]);
        }
        for (int i = 0; i < 6; ++i) for (int a = 0; a < 3; ++a)
          K[i][a] = (Pxz[i][0] * SI[0][a] + Pxz[i][1] * SI[1][a]) + Pxz[i][2] * SI[2][a];
        if (cfg->obs_type == SSA_OBS_AER) residual_aer(z, zp, yr);
        else for (int a = 0; a < 3; ++a) yr[a] = z[a] - zp[a];
        int nan = 0;
        for (int i = 0; i < 6; ++i) {
          o->x[i] = o->x[i] + ((K[i][0] * yr[0] + K[i][1] * yr[1]) + K[i][2] * yr[2]);
          nan |= isnan(o->x[i]);
        }
        for (int a = 0; a < 3; ++a) for (int j = 0; j < 6; ++j)
          T[a][j] = (S[a][0] * K[j][0] + S[a][1] * K[j][1]) + S[a][2] * K[j][2];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j)
          P[i][j] = P[i][j] - ((K[i][0] * T[0][j] + K[i][1] * T[1][j]) + K[i][2] * T[2][j]);
        if (o->y) memcpy(o->y, yr, sizeof(yr));
        if (o->S) memcpy(o->S, S, sizeof(S));
        if (o->sigmas_h) memcpy(o->sigmas_h, zs, sizeof(zs));
        if (o->updated) *o->updated = 1;
        if (!ok) fail_object(o, SSA_STATUS_LINALG | SSA_STATUS_IN_UPDATE);
        else if (nan) fail_object(o, SSA_STATUS_NAN | SSA_STATUS_IN_UPDATE);
      }
    }
  }

  if (flags & SSA_STEP_EPILOGUE) { /* results.py:60-72 observations, :36-47 error */
    for (int i = 0; i < 6; ++i) { o->obs[i] = o->x[i]; o->obs[6 + i] = P[i][i]; }
    double d[6];
    for (int i = 0; i < 6; ++i) d[i] = o->x[i] - o->x_true[i];
    *o->dpos = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    *o->dvel = sqrt((d[3] * d[3] + d[4] * d[4]) + d[5] * d[5]);
    *o->spos = sqrt((P[0][0] + P[1][1]) + P[2][2]);
    *o->svel = sqrt((P[3][3] + P[4][4]) + P[5][5]);
    *o->trace = ((((P[0][0] + P[1][1]) + P[2][2]) + P[3][3]) + P[4][4]) + P[5][5];
  }
}

/* Batch step over N objects. Host AoS layouts; P is the FULL 6x6 per object ([N][36]). */
int oracle_step(const ssa_ukf_cfg* cfg, const double* M, int flags, double* x_true, double* x, double* P,
                int32_t* status, int32_t* infl, const int32_t* actions, const double* z_noise, double* obs,
                double* dpos, double* dvel, double* spos, double* svel, double* trace, double* z_true,
                double* y, double* S, double* sigmas_h, uint8_t* visible, uint8_t* updated) {
  const int N = cfg->n_objects, m = cfg->m;
#pragma omp parallel for schedule(dynamic, 64)
  for (int n = 0; n < N; ++n) {
    obj_io o;
    o.x_true = x_true + 6 * (size_t)n; o.x = x + 6 * (size_t)n; o.P = P + 36 * (size_t)n;
    o.status = status + n; o.infl = infl + n;
    o.z_noise = z_noise ? z_noise + 3 * (size_t)n : NULL;
    o.obs = obs + 12 * (size_t)n;
    o.dpos = dpos + n; o.dvel = dvel + n; o.spos = spos + n; o.svel = svel + n; o.trace = trace + n;
    o.z_true = z_true ? z_true + 3 * (size_t)n : NULL;
    o.y = y ? y + 3 * (size_t)n : NULL;
    o.S = S ? S + 9 * (size_t)n : NULL;
    o.sigmas_h = sigmas_h ? sigmas_h + 39 * (size_t)n : NULL;
    o.visible = visible ? visible + n : NULL;
    o.updated = updated ? updated + n : NULL;
    int tasked = actions && (actions[n / m] == (n % m));
    object_step(cfg, (flags & SSA_STEP_M_PER_ENV) ? M + (size_t)(n / m) * 9 : M, flags, tasked, &o);
  }
  return 0;
}

/* ---- unit entry points --------------------------------------------------------------------- */
void oracle_fx(const double* x, double dt, double* out, int32_t* exc, int n) {
  for (int i = 0; i < n; ++i) exc[i] = fx_farnocchia(x + 6 * i, dt, out + 6 * i);
}
void oracle_rv2coe(const double* x, double* coe, int32_t* exc, int n) {
  for (int i = 0; i < n; ++i) { octx c = {0}; rv2coe(MU, x + 6 * i, x + 6 * i + 3, coe + 6 * i, &c); exc[i] = c.exc; }
}
void oracle_coe2rv(const double* coe, double* x, int n) { for (int i = 0; i < n; ++i) coe2rv(MU, coe + 6 * i, x + 6 * i); }
void oracle_hx_aer(const double* x, const double* M, const double* obs_itrs, const double* T, double* out, int n) {
  for (int i = 0; i < n; ++i) hx_aer(x + 6 * i, M, obs_itrs, T, out + 3 * i);
}
void oracle_aer2uvw(const double* a, double* u, int n) { for (int i = 0; i < n; ++i) aer2uvw(a + 3 * i, u + 3 * i); }
void oracle_uvw2aer(const double* u, double* a, int n) { for (int i = 0; i < n; ++i) uvw2aer(u + 3 * i, a + 3 * i); }
void oracle_residual_aer(const double* a, const double* b, double* c, int n) {
  for (int i = 0; i < n; ++i) residual_aer(a + 3 * i, b + 3 * i, c + 3 * i);
}
/* A full [n][36] (already scaled) -> U full [n][36]; ret = attempt number or -1 */
void oracle_robust_chol(const double* A, double* U, int32_t* ret, int n) {
  for (int i = 0; i < n; ++i) ret[i] = robust_cholesky((const double(*)[6])(A + 36 * i), (double(*)[6])(U + 36 * i));
}
void oracle_inv3(const double* S, double* SI, int32_t* ok, int n) {
  for (int i = 0; i < n; ++i) ok[i] = inv3((const double(*)[3])(S + 9 * i), (double(*)[3])(SI + 9 * i));
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}
