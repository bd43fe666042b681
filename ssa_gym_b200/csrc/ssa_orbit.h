// ssa_orbit.h — two-body propagation of one state vector by mean motion (Farnocchia et al. 2013
// regimes), the `fx` of the UKF.  One definition for the sm_100a kernels and the host twin.
//
// Reference behaviour being reproduced (read-only upstream, cited as file:line):
//   envs/farnocchia.py:1053-1062  fx_xyz_farnocchia(x, dt)   mu = 398600441800000.0
//   envs/farnocchia.py:1009-1050  farnocchia(k, r0, v0, tof)
//   envs/farnocchia.py:164-313    rv2coe        (tol = 1e-8, four branches, python '%')
//   envs/farnocchia.py:846-921    delta_t_from_nu (delta = 1e-2)
//   envs/farnocchia.py:924-1006   nu_from_delta_t
//   envs/farnocchia.py:336-353    newton (tol 1.48e-8, maxiter 50 / 100, NaN when not converged)
//   envs/farnocchia.py:572-601    M_to_E starting guess (M if e < 0.8 else pi*sign(M))
//   envs/farnocchia.py:756-843    near-parabolic series S_x / dS_x_alt, M_to_D_near_parabolic
//   envs/farnocchia.py:100-161    coe2rv = rv_pqw @ (R3(raan) R1(inc) R3(argp))^T
//
// This is a re-statement, not a translation: vectors live in registers, every repeated
// sub-expression of the reference (norm(r), norm(h), cos(nu), ...) is evaluated once (value
// preserving), sin and cos of one angle come from one sincos, the zero third column of the
// perifocal frame is never multiplied, and divisions by the constant mu become multiplications by
// its reciprocal.  All of that changes results only at the last-ulp level; parity against the
// reference's own numba functions is asserted in tests/test_oracle_golden.py (golden vectors) and
// tests/test_twin_vs_oracle.py (independent C oracle).
//
// Return value: SSA_FX_OK, or SSA_FX_EXC when the reference would have raised a Python exception
// inside numba (failed `assert`, ZeroDivisionError, RuntimeError) — the environment turns both an
// exception and a NaN result into a failed filter (ssa_tasker_simple_2.py:278-285).
#pragma once
#include "ssa_math.h"

#define SSA_MU 398600441800000.0
#define SSA_FX_OK 0
#define SSA_FX_EXC 1
#define SSA_SERIES_CAP 200000

SSA_HD double ssa_dot3(const double* a, const double* b) {
  return ssa_fma(a[2], b[2], ssa_fma(a[1], b[1], ssa_mul(a[0], b[0])));
}

// --- anomaly conversions (farnocchia.py:356-753) ------------------------------------------------
SSA_HD double ssa_nu_to_E(double nu, double ecc) {
  return ssa_mul(2.0, ssa_atan(ssa_mul(ssa_sqrt(ssa_div(1.0 - ecc, 1.0 + ecc)), ssa_tan(ssa_mul(nu, 0.5)))));
}
SSA_HD double ssa_E_to_nu(double E, double ecc) {
  return ssa_mul(2.0, ssa_atan(ssa_mul(ssa_sqrt(ssa_div(1.0 + ecc, 1.0 - ecc)), ssa_tan(ssa_mul(E, 0.5)))));
}
SSA_HD double ssa_nu_to_F(double nu, double ecc) {
  return ssa_mul(2.0, ssa_atanh(ssa_mul(ssa_sqrt(ssa_div(ecc - 1.0, ecc + 1.0)), ssa_tan(ssa_mul(nu, 0.5)))));
}
SSA_HD double ssa_F_to_nu(double F, double ecc) {
  return ssa_mul(2.0, ssa_atan(ssa_mul(ssa_sqrt(ssa_div(ecc + 1.0, ecc - 1.0)), ssa_tanh(ssa_mul(F, 0.5)))));
}
SSA_HD double ssa_E_to_M(double E, double ecc) { return ssa_fma(-ecc, ssa_sin(E), E); }
SSA_HD double ssa_F_to_M(double F, double ecc) { return ssa_fma(ecc, ssa_sinh(F), -F); }
SSA_HD double ssa_D_to_M(double D) { return D + ssa_div(ssa_mul(ssa_mul(D, D), D), 3.0); }
SSA_HD double ssa_M_to_D(double M) {  // Barker, farnocchia.py:649-652
  const double B = ssa_div(ssa_mul(3.0, M), 2.0);
  const double A = ssa_pow23(B + ssa_sqrt(ssa_fma(B, B, 1.0)));
  return ssa_div(ssa_mul(ssa_mul(2.0, A), B), (1.0 + A) + ssa_mul(A, A));
}

// Newton on Kepler's equation, elliptic (farnocchia.py:336-353 with regime "elliptic").
SSA_HD double ssa_newton_elliptic(double E0, double M, double ecc) {
  double p0 = E0;
  for (int it = 0; it < 50; ++it) {
    double s, c;
    ssa_sincos(p0, &s, &c);
    const double fval = ssa_fma(-ecc, s, p0) - M;
    const double fder = ssa_fma(-ecc, c, 1.0);
    const double p = p0 - ssa_div(fval, fder);
    if (ssa_fabs(p - p0) < SSA_C(NEWTON_TOL)) return p;
    p0 = p;
  }
  return ssa_nan();
}
SSA_HD double ssa_newton_hyperbolic(double F0, double M, double ecc) {
  double p0 = F0;
  for (int it = 0; it < 100; ++it) {
    const double fval = ssa_F_to_M(p0, ecc) - M;
    const double fder = ssa_fma(ecc, ssa_cosh(p0), -1.0);
    const double p = p0 - ssa_div(fval, fder);
    if (ssa_fabs(p - p0) < 1.48e-08) return p;
    p0 = p;
  }
  return ssa_nan();
}
SSA_HD double ssa_M_to_E(double M, double ecc, int* exc) {
  if (!(-SSA_C(PI) <= M && M <= SSA_C(PI))) { *exc = SSA_FX_EXC; return ssa_nan(); }  // assert, l.595
  double E0;
  if (ecc < 0.8) E0 = M;
  else E0 = (M > 0.0) ? SSA_C(PI) : ((M < 0.0) ? -SSA_C(PI) : ssa_mul(SSA_C(PI), M));  // pi*sign(M); sign(0)=0
  return ssa_newton_elliptic(E0, M, ecc);
}
SSA_HD double ssa_M_to_F(double M, double ecc) {
  return ssa_newton_hyperbolic(ssa_asinh(ssa_div(M, ecc)), M, ecc);
}

// near-parabolic series (farnocchia.py:769-798); `deriv` selects dS_x_alt.
SSA_HD double ssa_series_S(double ecc, double x, int deriv, int* exc) {
  if (!(ssa_fabs(x) < 1.0)) { *exc = SSA_FX_EXC; return ssa_nan(); }  // assert abs(x) < 1
  double S = 0.0, xk = 1.0;
  for (int k = 0; k < SSA_SERIES_CAP; ++k) {
    const double S_old = S;
    const double d = (double)(2 * k + 3);
    double term = ecc - ssa_div(1.0, d);
    if (deriv) term = ssa_mul(term, d);
    S = ssa_fma(term, xk, S);
    xk = ssa_mul(xk, x);
    if (ssa_fabs(S - S_old) < 1e-12) return S;
  }
  *exc = SSA_FX_EXC;  // the reference would spin here; report it as a failure instead of hanging the GPU
  return ssa_nan();
}
SSA_HD double ssa_D_to_M_near_parabolic(double D, double ecc, int* exc) {
  const double D2 = ssa_mul(D, D);
  const double x = ssa_mul(ssa_div(ecc - 1.0, ecc + 1.0), D2);
  const double S = ssa_series_S(ecc, x, 0, exc);
  const double ope = 1.0 + ecc;
  return ssa_fma(ssa_sqrt(ssa_div(2.0, ssa_mul(ssa_mul(ope, ope), ope))), ssa_mul(ssa_mul(D2, D), S),
                 ssa_mul(ssa_sqrt(ssa_div(2.0, ope)), D));
}
SSA_HD double ssa_M_to_D_near_parabolic(double M, double ecc, int* exc) {
  double D0 = ssa_M_to_D(M);
  const double ope = 1.0 + ecc;
  const double c1 = ssa_sqrt(ssa_div(2.0, ope));
  const double c3 = ssa_sqrt(ssa_div(2.0, ssa_mul(ssa_mul(ope, ope), ope)));
  for (int it = 0; it < 50; ++it) {
    const double fval = ssa_D_to_M_near_parabolic(D0, ecc, exc) - M;
    const double D2 = ssa_mul(D0, D0);
    const double x = ssa_mul(ssa_div(ecc - 1.0, ecc + 1.0), D2);
    const double dS = ssa_series_S(ecc, x, 1, exc);
    const double fder = ssa_fma(ssa_mul(c3, D2), dS, c1);
    if (*exc) return ssa_nan();
    const double D = D0 - ssa_div(fval, fder);
    if (ssa_fabs(D - D0) < 1.48e-08) return D;
    D0 = D;
  }
  return ssa_nan();
}

// --- time since periapsis <-> true anomaly (farnocchia.py:846-1006) ---------------------------
SSA_HD double ssa_delta_t_from_nu(double nu, double ecc, double k, double q, int* exc) {
  const double delta = 1e-2;
  if (!(-SSA_C(PI) <= nu && nu < SSA_C(PI))) { *exc = SSA_FX_EXC; return ssa_nan(); }  // assert, l.870
  const double q3 = ssa_mul(ssa_mul(q, q), q);
  double M, n;
  if (ecc < 1.0 - delta) {
    const double E = ssa_nu_to_E(nu, ecc);
    M = ssa_E_to_M(E, ecc);
    const double ome = 1.0 - ecc;
    n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(ome, ome), ome)), q3));
  } else if (1.0 - delta <= ecc && ecc < 1.0) {
    const double E = ssa_nu_to_E(nu, ecc);
    if (delta <= ssa_fma(-ecc, ssa_cos(E), 1.0)) {
      M = ssa_E_to_M(E, ecc);
      const double ome = 1.0 - ecc;
      n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(ome, ome), ome)), q3));
    } else {
      const double D = ssa_tan(ssa_mul(nu, 0.5));
      M = ssa_D_to_M_near_parabolic(D, ecc, exc);
      n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
    }
  } else if (ecc == 1.0) {
    const double D = ssa_tan(ssa_mul(nu, 0.5));
    M = ssa_D_to_M(D);
    n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
  } else if (ssa_fma(ecc, ssa_cos(nu), 1.0) < 0.0) {
    return ssa_nan();  // unfeasible region
  } else if (1.0 < ecc && ecc <= 1.0 + delta) {
    const double F = ssa_nu_to_F(nu, ecc);
    if (delta <= ssa_fma(ecc, ssa_cosh(F), -1.0)) {
      M = ssa_F_to_M(F, ecc);
      const double em1 = ecc - 1.0;
      n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(em1, em1), em1)), q3));
    } else {
      const double D = ssa_tan(ssa_mul(nu, 0.5));
      M = ssa_D_to_M_near_parabolic(D, ecc, exc);
      n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
    }
  } else if (1.0 + delta < ecc) {
    const double F = ssa_nu_to_F(nu, ecc);
    M = ssa_F_to_M(F, ecc);
    const double em1 = ecc - 1.0;
    n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(em1, em1), em1)), q3));
  } else {
    *exc = SSA_FX_EXC;  // RuntimeError (ecc is NaN)
    return ssa_nan();
  }
  if (n == 0.0) { *exc = SSA_FX_EXC; return ssa_nan(); }  // ZeroDivisionError in numba
  return ssa_div(M, n);
}

SSA_HD double ssa_wrap_pi(double a) {  // (a + pi) % (2 pi) - pi with python's % (farnocchia.py:311, 954, 967)
  double m = ssa_fmod_pos_inv(a + SSA_C(PI), SSA_C(TWOPI), SSA_C(INV_TWOPI));
  if (m != 0.0 && m < 0.0) m = ssa_add(m, SSA_C(TWOPI));
  return m - SSA_C(PI);
}

SSA_HD double ssa_nu_from_delta_t(double delta_t, double ecc, double k, double q, int* exc) {
  const double delta = 1e-2;
  const double q3 = ssa_mul(ssa_mul(q, q), q);
  double nu;
  if (ecc < 1.0 - delta) {
    const double ome = 1.0 - ecc;
    const double n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(ome, ome), ome)), q3));
    const double M = ssa_mul(n, delta_t);
    const double E = ssa_M_to_E(ssa_wrap_pi(M), ecc, exc);
    nu = ssa_E_to_nu(E, ecc);
  } else if (1.0 - delta <= ecc && ecc < 1.0) {
    const double E_delta = ssa_acos(ssa_div(1.0 - delta, ecc));
    const double ome = 1.0 - ecc;
    double n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(ome, ome), ome)), q3));
    double M = ssa_mul(n, delta_t);
    if (ssa_E_to_M(E_delta, ecc) <= ssa_fabs(M)) {
      const double E = ssa_M_to_E(ssa_wrap_pi(M), ecc, exc);
      nu = ssa_E_to_nu(E, ecc);
    } else {
      n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
      M = ssa_mul(n, delta_t);
      nu = ssa_mul(2.0, ssa_atan(ssa_M_to_D_near_parabolic(M, ecc, exc)));
    }
  } else if (ecc == 1.0) {
    const double n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
    nu = ssa_mul(2.0, ssa_atan(ssa_M_to_D(ssa_mul(n, delta_t))));
  } else if (1.0 < ecc && ecc <= 1.0 + delta) {
    const double F_delta = ssa_acosh(ssa_div(1.0 + delta, ecc));
    const double em1 = ecc - 1.0;
    double n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(em1, em1), em1)), q3));
    double M = ssa_mul(n, delta_t);
    if (ssa_F_to_M(F_delta, ecc) <= ssa_fabs(M)) {
      nu = ssa_F_to_nu(ssa_M_to_F(M, ecc), ecc);
    } else {
      n = ssa_sqrt(ssa_div(k, ssa_mul(2.0, q3)));
      M = ssa_mul(n, delta_t);
      nu = ssa_mul(2.0, ssa_atan(ssa_M_to_D_near_parabolic(M, ecc, exc)));
    }
  } else {
    // strong hyperbolic — also where a NaN eccentricity lands (farnocchia.py:999-1004)
    const double em1 = ecc - 1.0;
    const double n = ssa_sqrt(ssa_div(ssa_mul(k, ssa_mul(ssa_mul(em1, em1), em1)), q3));
    const double M = ssa_mul(n, delta_t);
    nu = ssa_F_to_nu(ssa_M_to_F(M, ecc), ecc);
  }
  return nu;
}

// --- the propagator -----------------------------------------------------------------------------
// coe[6] = p, ecc, inc, raan, argp, nu  (farnocchia.py:164-313)
SSA_HD int ssa_rv2coe(const double* x, double* coe) {
  const double k = SSA_C(MU);
  const double kinv = SSA_C(MU_INV);
  const double tol = SSA_C(TOL8);
  const double* r = x;
  const double* v = x + 3;
  double h[3];
  h[0] = ssa_fma(r[1], v[2], -ssa_mul(r[2], v[1]));
  h[1] = ssa_fma(r[2], v[0], -ssa_mul(r[0], v[2]));
  h[2] = ssa_fma(r[0], v[1], -ssa_mul(r[1], v[0]));
  const double nvec0 = -h[1], nvec1 = h[0];  // cross([0,0,1], h) = (-hy, hx, 0)
  const double rr = ssa_dot3(r, r);
  const double rn = ssa_sqrt(rr);
  const double vv = ssa_dot3(v, v);
  const double rv = ssa_dot3(r, v);
  const double hh = ssa_dot3(h, h);
  const double hn = ssa_sqrt(hh);
  if (rn == 0.0 || hn == 0.0) return SSA_FX_EXC;  // ZeroDivisionError in numba (l.273, l.276)
  const double c1 = vv - ssa_div(k, rn);
  double e[3];
  e[0] = ssa_mul(ssa_fma(c1, r[0], -ssa_mul(rv, v[0])), kinv);
  e[1] = ssa_mul(ssa_fma(c1, r[1], -ssa_mul(rv, v[1])), kinv);
  e[2] = ssa_mul(ssa_fma(c1, r[2], -ssa_mul(rv, v[2])), kinv);
  const double ecc = ssa_sqrt(ssa_dot3(e, e));
  const double p = ssa_mul(hh, kinv);
  const double inc = ssa_acos(ssa_div(h[2], hn));
  const int circular = ecc < tol;
  const int equatorial = ssa_fabs(inc) < tol;
  double raan, argp, nu;
  if (equatorial && !circular) {
    raan = 0.0;
    argp = ssa_pymod(ssa_atan2(e[1], e[0]), SSA_C(TWOPI));
    // h . cross(e, r) / |h|
    double c[3];
    c[0] = ssa_fma(e[1], r[2], -ssa_mul(e[2], r[1]));
    c[1] = ssa_fma(e[2], r[0], -ssa_mul(e[0], r[2]));
    c[2] = ssa_fma(e[0], r[1], -ssa_mul(e[1], r[0]));
    nu = ssa_atan2(ssa_div(ssa_dot3(h, c), hn), ssa_dot3(r, e));
  } else if (!equatorial && circular) {
    raan = ssa_pymod(ssa_atan2(nvec1, nvec0), SSA_C(TWOPI));
    argp = 0.0;
    double c[3];  // cross(h, n) = (-hz*hx, -hz*hy, hx^2+hy^2)
    c[0] = -ssa_mul(h[2], h[0]);
    c[1] = -ssa_mul(h[2], h[1]);
    c[2] = ssa_fma(h[1], h[1], ssa_mul(h[0], h[0]));
    nu = ssa_atan2(ssa_div(ssa_dot3(r, c), hn), ssa_fma(r[1], nvec1, ssa_mul(r[0], nvec0)));
  } else if (equatorial && circular) {
    raan = 0.0;
    argp = 0.0;
    nu = ssa_pymod(ssa_atan2(r[1], r[0]), SSA_C(TWOPI));
  } else {
    const double ome2 = ssa_fma(-ecc, ecc, 1.0);
    if (ome2 == 0.0) return SSA_FX_EXC;  // ZeroDivisionError (l.295)
    const double a = ssa_div(p, ome2);
    const double ka = ssa_mul(k, a);
    if (a > 0.0) {
      const double e_se = ssa_div(rv, ssa_sqrt(ka));
      const double e_ce = ssa_fma(ssa_mul(rn, vv), kinv, -1.0);
      nu = ssa_E_to_nu(ssa_atan2(e_se, e_ce), ecc);
    } else {
      const double e_sh = ssa_div(rv, ssa_sqrt(-ka));
      const double e_ch = ssa_fma(ssa_mul(rn, vv), kinv, -1.0);
      if (e_ch - e_sh == 0.0) return SSA_FX_EXC;
      nu = ssa_F_to_nu(ssa_mul(ssa_log(ssa_div(e_ch + e_sh, e_ch - e_sh)), 0.5), ecc);
    }
    raan = ssa_pymod(ssa_atan2(nvec1, nvec0), SSA_C(TWOPI));
    const double px = ssa_fma(r[1], nvec1, ssa_mul(r[0], nvec0));
    double c[3];
    c[0] = -ssa_mul(h[2], h[0]);
    c[1] = -ssa_mul(h[2], h[1]);
    c[2] = ssa_fma(h[1], h[1], ssa_mul(h[0], h[0]));
    const double py = ssa_div(ssa_dot3(r, c), hn);
    argp = ssa_pymod(ssa_atan2(py, px) - nu, SSA_C(TWOPI));
  }
  nu = ssa_wrap_pi(nu);
  coe[0] = p; coe[1] = ecc; coe[2] = inc; coe[3] = raan; coe[4] = argp; coe[5] = nu;
  return SSA_FX_OK;
}

// farnocchia.py:100-161 (rv_pqw 14-73, rotation matrices 76-97)
SSA_HD void ssa_coe2rv(const double* coe, double* out) {
  const double k = SSA_C(MU);
  const double p = coe[0], ecc = coe[1];
  double snu, cnu, sO, cO, si, ci, sw, cw;
  ssa_sincos(coe[5], &snu, &cnu);
  ssa_sincos(coe[3], &sO, &cO);
  ssa_sincos(coe[2], &si, &ci);
  ssa_sincos(coe[4], &sw, &cw);
  const double rp = ssa_div(p, ssa_fma(ecc, cnu, 1.0));
  const double vp = ssa_sqrt(ssa_div(k, p));
  const double rx = ssa_mul(cnu, rp), ry = ssa_mul(snu, rp);
  const double vx = ssa_mul(-snu, vp), vy = ssa_mul(ecc + cnu, vp);
  // m1 = R3(raan) R1(inc): columns 0 and 1
  const double m00 = cO, m01 = ssa_mul(-sO, ci);
  const double m10 = sO, m11 = ssa_mul(cO, ci);
  const double m21 = si;  // m20 = 0
  // rm = m1 R3(argp): rm[i][0] = m_i0 cw + m_i1 sw ; rm[i][1] = -m_i0 sw + m_i1 cw
  const double a00 = ssa_fma(m00, cw, ssa_mul(m01, sw)), a01 = ssa_fma(m01, cw, -ssa_mul(m00, sw));
  const double a10 = ssa_fma(m10, cw, ssa_mul(m11, sw)), a11 = ssa_fma(m11, cw, -ssa_mul(m10, sw));
  const double a20 = ssa_mul(m21, sw), a21 = ssa_mul(m21, cw);
  out[0] = ssa_fma(rx, a00, ssa_mul(ry, a01));
  out[1] = ssa_fma(rx, a10, ssa_mul(ry, a11));
  out[2] = ssa_fma(rx, a20, ssa_mul(ry, a21));
  out[3] = ssa_fma(vx, a00, ssa_mul(vy, a01));
  out[4] = ssa_fma(vx, a10, ssa_mul(vy, a11));
  out[5] = ssa_fma(vx, a20, ssa_mul(vy, a21));
}

// The literal restatement: rv2coe -> time since periapsis -> true anomaly -> coe2rv, every regime.
SSA_HD_NOINLINE int ssa_fx_general(const double* x, double tof, double* out) {
  double coe[6];
  int exc = ssa_rv2coe(x, coe);
  if (exc) {
    for (int i = 0; i < 6; ++i) out[i] = ssa_nan();
    return exc;
  }
  const double ecc = coe[1];
  const double q = ssa_div(coe[0], 1.0 + ecc);
  const double dt0 = ssa_delta_t_from_nu(coe[5], ecc, SSA_C(MU), q, &exc);
  const double dt1 = dt0 + tof;
  coe[5] = ssa_nu_from_delta_t(dt1, ecc, SSA_C(MU), q, &exc);
  ssa_coe2rv(coe, out);
  return exc;
}

// fx.  The overwhelmingly common case — a strong-elliptic (1e-8 <= e < 0.99), non-equatorial orbit, i.e. the
// "general" branch of rv2coe (farnocchia.py:294-309) followed by the strong-elliptic regime (:871-875,
// :948-955) — is evaluated in a streamlined, mathematically identical form; every other case goes through
// ssa_fx_general.  What the fast path changes relative to the literal sequence (all at the last-ulp level,
// pinned against the reference's numba function in tests/test_oracle_golden.py):
//   * the classical angles raan, inc, argp are never formed: the reference only uses them through sin/cos
//     in coe2rv's rotation matrix, and those follow directly from the vectors
//       cos(raan) = -h_y/h_xy   sin(raan) = h_x/h_xy   cos(inc) = h_z/|h|   sin(inc) = h_xy/|h|
//       argp = u0 - nu0  ->  cos/sin(argp) by the angle-difference formulas, u0 = atan2(py, px);
//   * the E0 -> nu0 -> E round trip (E_to_nu then nu_to_E, :300 and :873) is the identity up to rounding:
//     M0 = E0 - e sin E0 is taken from E0 directly, and cos/sin of nu0 and of the propagated nu come from
//     cos nu = (cos E - e)/(1 - e cos E), sin nu = sqrt(1 - e^2) sin E/(1 - e cos E) instead of two
//     half-angle tangent/arctangent conversions;
//   * `equatorial` (|acos(h_z/|h|)| < 1e-8) is decided as h_z/|h| == 1.0 — the only double whose acos is
//     below 1e-8 (acos(1 - 2^-53) = 1.49e-8).
//   * algebraically equal forms that save divisions: n = sqrt(k/a^3), M = M0 + n dt, |r'| = a(1 - e cos E),
//     sqrt(px^2 + py^2) = |r| h_xy, reciprocals of |r|, |h|, h_xy formed once.
//   * sin / cos of E0 follow from (e sin E0, e cos E0) by normalisation, those of the propagated E from the last Newton
//     iterate by a second-order step: 2.2 sincos evaluations (the Newton iterations) remain of 4.2;
//   * reciprocals shared: 1/|r|, 1/|h|, 1/h_xy come from one division, 1/p = k/|h|^2, 1/a = (1-e^2)/p,
//     sqrt(k/p) = k/|h|, r.v/sqrt(k a) and sqrt(k/a^3) are products with sqrt(k a) and 1/a.
// 1 atan2 + ~5.5 sincos + ~6 divisions + 6 square roots instead of 7 atan2 + acos + 11 sincos + ~30 divisions.
// Streamlined strong-hyperbolic propagation (ecc > 1 + delta, farnocchia.py:911-917, 994-1001).  Filter estimates that
// an update pushed beyond escape speed are the usual way a state leaves the elliptic fast path (two or three objects
// of the C2 catalog from step ~100 on, with eccentricities up to 1e4); through the literal restatement each of their
// propagations costs ~3x a normal one and its warp lengthens k_fx by 6-14 us.  Same construction as the elliptic
// path: e sinh F0 = r.v / sqrt(k |a|), e cosh F0 = r v^2 / k - 1 give F0 (the reference's logarithm), M0 = e sinh F0 -
// F0, M = M0 + n tof, Newton from asinh(M / e) exactly as M_to_F, then cos nu = (e - cosh F)/(e cosh F - 1),
// sin nu = sqrt(e^2 - 1) sinh F / (e cosh F - 1), |r'| = |a| (e cosh F - 1), and the rotation from the vectors.
// Called with the quantities ssa_fx has already formed; a real function: only stray lanes come here.
SSA_HD_NOINLINE int ssa_fx_hyperbolic(const double* x, double tof, double* out) {
  const double k = SSA_C(MU), kinv = SSA_C(MU_INV);
  const double* r = x;
  const double* v = x + 3;
  double h[3];
  h[0] = ssa_fma(r[1], v[2], -ssa_mul(r[2], v[1]));
  h[1] = ssa_fma(r[2], v[0], -ssa_mul(r[0], v[2]));
  h[2] = ssa_fma(r[0], v[1], -ssa_mul(r[1], v[0]));
  const double rr = ssa_dot3(r, r), vv = ssa_dot3(v, v), rv = ssa_dot3(r, v), hh = ssa_dot3(h, h);
  const double rn = ssa_sqrt(rr), hn = ssa_sqrt(hh);
  const double hxy2 = ssa_fma(h[1], h[1], ssa_mul(h[0], h[0]));
  const bool planar = (hxy2 == 0.0);
  const double inv_rn = ssa_div(1.0, rn), inv_hn = ssa_div(1.0, hn);
  const double c1 = vv - ssa_mul(k, inv_rn);
  const double e0 = ssa_mul(ssa_fma(c1, r[0], -ssa_mul(rv, v[0])), kinv);
  const double e1 = ssa_mul(ssa_fma(c1, r[1], -ssa_mul(rv, v[1])), kinv);
  const double e2 = ssa_mul(ssa_fma(c1, r[2], -ssa_mul(rv, v[2])), kinv);
  const double ecc = ssa_sqrt(ssa_fma(e2, e2, ssa_fma(e1, e1, ssa_mul(e0, e0))));
  const double ci = ssa_mul(h[2], inv_hn);
  const double p = ssa_mul(hh, kinv);
  const double em2 = ssa_fma(ecc, ecc, -1.0);          // e^2 - 1 > 0
  const double na = ssa_div(p, em2);                    // |a|
  const double e_sh = ssa_div(rv, ssa_sqrt(ssa_mul(k, na)));
  const double e_ch = ssa_fma(ssa_mul(rn, vv), kinv, -1.0);
  const double F0 = ssa_mul(ssa_log(ssa_div(e_ch + e_sh, e_ch - e_sh)), 0.5);   // farnocchia.py:299-301
  const double sq = ssa_sqrt(em2);
  const double inv_e = ssa_div(1.0, ecc);
  const double d0 = ssa_div(1.0, e_ch - 1.0);           // 1 / (e cosh F0 - 1)
  const double cnu0 = ssa_mul(ecc - ssa_mul(e_ch, inv_e), d0);
  const double snu0 = ssa_mul(ssa_mul(sq, ssa_mul(e_sh, inv_e)), d0);
  const double n = ssa_sqrt(ssa_div(k, ssa_mul(ssa_mul(na, na), na)));
  const double M0 = e_sh - F0;
  const double M = ssa_fma(n, tof, M0);
  const double F1 = ssa_newton_hyperbolic(ssa_asinh(ssa_div(M, ecc)), M, ecc);   // M_to_F, farnocchia.py:604-622
  const double sh = ssa_sinh(F1), ch = ssa_cosh(F1);
  const double den = ssa_fma(ecc, ch, -1.0);
  const double d1 = ssa_div(1.0, den);
  const double cnu = ssa_mul(ecc - ch, d1), snu = ssa_mul(ssa_mul(sq, sh), d1);
  const double px = planar ? r[0] : ssa_fma(r[1], h[0], -ssa_mul(r[0], h[1]));
  const double py =
      planar ? r[1] : ssa_mul(ssa_fma(r[2], hxy2, -ssa_mul(h[2], ssa_fma(r[1], h[1], ssa_mul(r[0], h[0])))), inv_hn);
  const double hxy = ssa_sqrt(hxy2);
  const double inv_hxy = ssa_div(1.0, planar ? 1.0 : hxy);
  const double inv_rho = ssa_mul(inv_rn, inv_hxy);
  const double cu0 = ssa_mul(px, inv_rho), su0 = ssa_mul(py, inv_rho);
  const double cw = ssa_fma(cu0, cnu0, ssa_mul(su0, snu0)), sw = ssa_fma(su0, cnu0, -ssa_mul(cu0, snu0));
  const double cO = planar ? 1.0 : -ssa_mul(h[1], inv_hxy), sO = planar ? 0.0 : ssa_mul(h[0], inv_hxy);
  const double si = ssa_mul(hxy, inv_hn);
  const double rp = ssa_mul(na, den);
  const double vp = ssa_sqrt(ssa_div(k, p));
  const double rx = ssa_mul(cnu, rp), ry = ssa_mul(snu, rp);
  const double vx = ssa_mul(-snu, vp), vy = ssa_mul(ecc + cnu, vp);
  const double m00 = cO, m01 = ssa_mul(-sO, ci);
  const double m10 = sO, m11 = ssa_mul(cO, ci);
  const double a00 = ssa_fma(m00, cw, ssa_mul(m01, sw)), a01 = ssa_fma(m01, cw, -ssa_mul(m00, sw));
  const double a10 = ssa_fma(m10, cw, ssa_mul(m11, sw)), a11 = ssa_fma(m11, cw, -ssa_mul(m10, sw));
  const double a20 = ssa_mul(si, sw), a21 = ssa_mul(si, cw);
  out[0] = ssa_fma(rx, a00, ssa_mul(ry, a01));
  out[1] = ssa_fma(rx, a10, ssa_mul(ry, a11));
  out[2] = ssa_fma(rx, a20, ssa_mul(ry, a21));
  out[3] = ssa_fma(vx, a00, ssa_mul(vy, a01));
  out[4] = ssa_fma(vx, a10, ssa_mul(vy, a11));
  out[5] = ssa_fma(vx, a20, ssa_mul(vy, a21));
  return 0;
}

SSA_HD int ssa_fx(const double* x, double tof, double* out) {
  const double k = SSA_C(MU), kinv = SSA_C(MU_INV);
  const double* r = x;
  const double* v = x + 3;
  double h[3];
  h[0] = ssa_fma(r[1], v[2], -ssa_mul(r[2], v[1]));
  h[1] = ssa_fma(r[2], v[0], -ssa_mul(r[0], v[2]));
  h[2] = ssa_fma(r[0], v[1], -ssa_mul(r[1], v[0]));
  const double rr = ssa_dot3(r, r), vv = ssa_dot3(v, v), rv = ssa_dot3(r, v), hh = ssa_dot3(h, h);
  const double rn = ssa_sqrt_i(rr), hn = ssa_sqrt_i(hh);
  const double hxy2 = ssa_fma(h[1], h[1], ssa_mul(h[0], h[0]));
  // planar: the orbit plane is the reference plane (h along +z exactly: the GEO class of the reference's catalog,
  // inclination drawn from uniform(0, 0)).  rv2coe then calls the orbit equatorial and puts the node on the x axis.
  const bool planar = (hxy2 == 0.0);
  bool fast = (rn > 0.0) && (hn > 0.0) && (hxy2 > 0.0 || h[2] > 0.0);
  double ecc = 0.0, e_ce = 0.0, ci = 0.0, inv_hn = 0.0, inv_rn = 0.0, inv_hxy = 0.0, hxy = 0.0;
  if (fast) {
    // the three reciprocals 1/|r|, 1/|h|, 1/h_xy from ONE division (t = 1 / (|r| |h| h_xy)); h_xy -> 1 for a planar orbit.
    // Divisions and square roots are a third of the instructions of this function (~16 each, 8 of them FP64): every
    // quotient below that has an algebraically equal product form uses it (all value-preserving to an ulp or two).
    hxy = ssa_sqrt_i(hxy2);
    const double hxy_s = planar ? 1.0 : hxy;
    const double rh = ssa_mul(rn, hn);
    const double t3 = ssa_div_i(1.0, ssa_mul(rh, hxy_s));
    inv_hxy = ssa_mul(t3, rh);
    const double t2 = ssa_mul(t3, hxy_s);  // 1 / (|r| |h|)
    inv_rn = ssa_mul(t2, hn);
    inv_hn = ssa_mul(t2, rn);
    // one Newton step each (y += y (1 - d y)): the shared reciprocals are good to an ulp again
    inv_rn = ssa_fma(inv_rn, ssa_fma(-rn, inv_rn, 1.0), inv_rn);
    inv_hn = ssa_fma(inv_hn, ssa_fma(-hn, inv_hn, 1.0), inv_hn);
    inv_hxy = ssa_fma(inv_hxy, ssa_fma(-hxy_s, inv_hxy, 1.0), inv_hxy);
    const double c1 = vv - ssa_mul(k, inv_rn);
    const double e0 = ssa_mul(ssa_fma(c1, r[0], -ssa_mul(rv, v[0])), kinv);
    const double e1 = ssa_mul(ssa_fma(c1, r[1], -ssa_mul(rv, v[1])), kinv);
    const double e2 = ssa_mul(ssa_fma(c1, r[2], -ssa_mul(rv, v[2])), kinv);
    ecc = ssa_sqrt_i(ssa_fma(e2, e2, ssa_fma(e1, e1, ssa_mul(e0, e0))));
    ci = ssa_mul(h[2], inv_hn);
    e_ce = ssa_fma(ssa_mul(rn, vv), kinv, -1.0);
    // Circular orbits: rv2coe drops the periapsis direction when ecc < 1e-8 (argp = 0, nu = argument of latitude)
    // but keeps ecc in the anomaly conversions, an O(a ecc) inconsistency; keeping the direction, as this path
    // does, agrees with it to a * ecc, so only ecc < 1e-12 (< 5e-5 m) may take this path; [1e-12, 1e-8) goes to
    // the literal restatement.
    fast = ((ecc >= SSA_C(TOL8)) || (ecc < SSA_C(TOL12))) && (ecc < SSA_C(DELTA99)) && (planar || ci < 1.0);
  }
  if (!fast) {
    // strong hyperbolic regime with a well-defined plane: the streamlined hyperbolic path (finite e only; the band
    // [0.99, 1.01], degenerate and non-finite states keep the literal restatement with its exact failure semantics)
    const bool hyper = (rn > 0.0) && (hn > 0.0) && (hxy2 > 0.0 || h[2] > 0.0) && (ecc > SSA_C(HYP101)) && (ecc < 1e12) &&
                       (planar || ci < 1.0);
    // (through copies: handing x / out themselves to the out-of-line functions would make the caller's arrays escape and
    // pin them to local memory on the fast path too — measured in k_predict_tile: 25 local loads / stores per propagation)
    double xs[6], fo[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) xs[j] = x[j];
    const int e = hyper ? ssa_fx_hyperbolic(xs, tof, fo) : ssa_fx_general(xs, tof, fo);
#pragma unroll
    for (int j = 0; j < 6; ++j) out[j] = fo[j];
    return e;
  }

  const double p = ssa_mul(hh, kinv);
  const double inv_p = ssa_mul(k, ssa_mul(inv_hn, inv_hn));  // 1/p = k / |h|^2
  const double ome2 = ssa_fma(-ecc, ecc, 1.0);
  const double inv_ome2 = ssa_div_i(1.0, ome2);
  double a = ssa_mul(p, inv_ome2);
  a = ssa_fma(ssa_fma(-a, ome2, p), inv_ome2, a);  // residual correction: a = p / (1 - e^2) to half an ulp
  double inv_a = ssa_mul(ome2, inv_p);
  inv_a = ssa_fma(inv_a, ssa_fma(-a, inv_a, 1.0), inv_a);  // Newton step: 1/a to an ulp (n = sqrt(k a) / a^2 multiplies tof)
  const double s_ka = ssa_sqrt_i(ssa_mul(k, a));
  const double e_se = ssa_mul(ssa_mul(rv, s_ka), ssa_mul(inv_a, kinv));  // r.v / sqrt(k a)
  const double E0 = ssa_atan2_i(e_se, e_ce);
  // sin E0, cos E0 without evaluating them: (e_se, e_ce) = e (sin E0, cos E0), normalised by its own length (not by ecc:
  // for a near-circular orbit the two differ by the cancellation error of the eccentricity vector)
  ssa_sc sc0;
  {
    const double hh2 = ssa_fma(e_se, e_se, ssa_mul(e_ce, e_ce));
    const double inv_h = ssa_div_i(1.0, ssa_sqrt_i(hh2));
    const bool zero = (hh2 == 0.0);  // atan2(0, 0) = 0
    const double s_ = ssa_mul(e_se, inv_h), c_ = ssa_mul(e_ce, inv_h);
    // one renormalisation step: s^2 + c^2 = 1 to half an ulp (the perifocal rotation built from it must be orthonormal)
    const double k2 = ssa_fma(-0.5, ssa_fma(s_, s_, ssa_mul(c_, c_)), 1.5);
    sc0.s = zero ? 0.0 : ssa_mul(s_, k2);
    sc0.c = zero ? 1.0 : ssa_mul(c_, k2);
  }
  const double sq = ssa_sqrt_i(ome2);
  const double d0 = ssa_div_i(1.0, ssa_fma(-ecc, sc0.c, 1.0));
  const double cnu0 = ssa_mul(sc0.c - ecc, d0), snu0 = ssa_mul(ssa_mul(sq, sc0.s), d0);
  // mean motion and mean anomaly (farnocchia.py:874-875, 950-951)
  // n = sqrt(k (1-e)^3 / q^3) with q = p/(1+e) = a (1-e)  ->  sqrt(k / a^3) = sqrt(k a) / a^2;  M = n (M0/n + tof) -> M0 + n tof
  const double n = ssa_mul(s_ka, ssa_mul(inv_a, inv_a));
  const double M0 = ssa_fma(-ecc, sc0.s, E0);
  const double M = ssa_fma(n, tof, M0);
  int exc = 0;
  const double Mw = ssa_wrap_pi(M);
  // newton(), farnocchia.py:336-353.  sin / cos of the converged anomaly E1 = p0 + d come from those of the last iterate
  // p0 (|d| < 1.48e-8, so the terms beyond d^2 are below 1e-24): sin E1 = s + d (c - d s / 2), cos E1 = c - d (s + d c / 2).
  ssa_sc sc1;
  sc1.s = ssa_nan();
  sc1.c = ssa_nan();
  if (!(-SSA_C(PI) <= Mw && Mw <= SSA_C(PI))) {  // assert of M_to_E (farnocchia.py:595): only a NaN gets here
    exc = SSA_FX_EXC;
  } else {
    double p0 = (ecc < 0.8) ? Mw : ((Mw > 0.0) ? SSA_C(PI) : ((Mw < 0.0) ? -SSA_C(PI) : ssa_mul(SSA_C(PI), Mw)));
    for (int it = 0; it < 50; ++it) {
      const ssa_sc sn = ssa_sincos_i(p0);
      const double fval = ssa_fma(-ecc, sn.s, p0) - Mw;
      const double fder = ssa_fma(-ecc, sn.c, 1.0);
      const double pn = p0 - ssa_div_i(fval, fder);
      const double d = pn - p0;
      if (ssa_fabs(d) < SSA_C(NEWTON_TOL)) {
        const double hd = ssa_mul(0.5, d);
        sc1.s = ssa_fma(d, ssa_fma(-hd, sn.s, sn.c), sn.s);
        sc1.c = ssa_fma(-d, ssa_fma(hd, sn.c, sn.s), sn.c);
        break;
      }
      p0 = pn;
    }
  }
  const double den1 = ssa_fma(-ecc, sc1.c, 1.0);
  const double d1 = ssa_div_i(1.0, den1);
  const double cnu = ssa_mul(sc1.c - ecc, d1), snu = ssa_mul(ssa_mul(sq, sc1.s), d1);
  // argument of latitude of the initial position: px = r.n, py = r.(h x n)/|h|, n = (-h_y, h_x, 0)
  const double px = planar ? r[0] : ssa_fma(r[1], h[0], -ssa_mul(r[0], h[1]));
  const double py =
      planar ? r[1] : ssa_mul(ssa_fma(r[2], hxy2, -ssa_mul(h[2], ssa_fma(r[1], h[1], ssa_mul(r[0], h[0])))), inv_hn);
  const double inv_rho = ssa_mul(inv_rn, inv_hxy);  // sqrt(px^2 + py^2) = |r| h_xy: r lies in the orbital plane
  const double cu0 = ssa_mul(px, inv_rho), su0 = ssa_mul(py, inv_rho);
  const double cw = ssa_fma(cu0, cnu0, ssa_mul(su0, snu0)), sw = ssa_fma(su0, cnu0, -ssa_mul(cu0, snu0));
  // rotation (farnocchia.py:90-97) from the vectors
  const double cO = planar ? 1.0 : -ssa_mul(h[1], inv_hxy), sO = planar ? 0.0 : ssa_mul(h[0], inv_hxy);
  const double si = ssa_mul(hxy, inv_hn);
  // perifocal position / velocity (farnocchia.py:70-72)
  const double rp = ssa_mul(a, den1);     // p/(1 + e cos nu) = a (1 - e cos E)
  const double vp = ssa_mul(k, inv_hn);   // sqrt(k/p) = k / |h|
  const double rx = ssa_mul(cnu, rp), ry = ssa_mul(snu, rp);
  const double vx = ssa_mul(-snu, vp), vy = ssa_mul(ecc + cnu, vp);
  const double m00 = cO, m01 = ssa_mul(-sO, ci);
  const double m10 = sO, m11 = ssa_mul(cO, ci);
  const double a00 = ssa_fma(m00, cw, ssa_mul(m01, sw)), a01 = ssa_fma(m01, cw, -ssa_mul(m00, sw));
  const double a10 = ssa_fma(m10, cw, ssa_mul(m11, sw)), a11 = ssa_fma(m11, cw, -ssa_mul(m10, sw));
  const double a20 = ssa_mul(si, sw), a21 = ssa_mul(si, cw);
  out[0] = ssa_fma(rx, a00, ssa_mul(ry, a01));
  out[1] = ssa_fma(rx, a10, ssa_mul(ry, a11));
  out[2] = ssa_fma(rx, a20, ssa_mul(ry, a21));
  out[3] = ssa_fma(vx, a00, ssa_mul(vy, a01));
  out[4] = ssa_fma(vx, a10, ssa_mul(vy, a11));
  out[5] = ssa_fma(vx, a20, ssa_mul(vy, a21));
  return exc;
}
