// ssa_meas.h — measurement model of the UKF: GCRS state -> ITRS -> topocentric az/el/range, the
// Cartesian (uvw) mean of angular measurements and the wrapped residual.  Host+device, see
// ssa_math.h for the bit-exactness contract.
//
// Reference behaviour reproduced (file:line in the read-only upstream):
//   envs/dynamics.py:219-231          hx_aer_erfa : x_itrs = trans_matrix @ x[:3]; ecef2aer(...)
//   envs/transformations.py:329-352   ecef2aer    : T(lat,lon)^T (x_itrs - obs_itrs); az = atan2(E1,E0) (+2pi if <0);
//                                                   el = asin(E2/r); r = |x_itrs - obs_itrs|
//   envs/transformations.py:283-297   aer2uvw     : (r cos el cos az, r cos el sin az, r sin el)
//   envs/transformations.py:300-316   uvw2aer
//   envs/dynamics.py:342-354          mean_z_uvw  : uvw2aer(Wm . aer2uvw(sigmas))
//   envs/dynamics.py:260-267          residual_z_aer : [atan2(sin d, cos d), d_el, d_r]
//   envs/dynamics.py:207-217,270-278  hx_xyz / residual_xyz / mean_xyz (Cartesian variant, tests.py Test 6/7)
//
// The observer rotation T (transformations.py:341-343) only depends on the observer latitude and
// longitude; the reference rebuilds it on every call, here the host computes its nine entries once
// with the same expressions and ships them in the constant block (`ssa_obs`).
#pragma once
#include "ssa_math.h"

typedef struct {
  double M[9];         // GCRS -> ITRS rotation for the current step (row-major), host input (SS2:137)
  double obs_itrs[3];  // observer ECEF [m]  (lla2ecef, transformations.py:216-235)
  double T[9];         // trans_uvw_ecef, row-major, as in transformations.py:341-343
} ssa_obs;

// hx for the 'aer' observation type.  out = [az, el, range].  INL: inlined math (k_hx), same arithmetic.
// `enz` (optional): the topocentric vector e = T^T (x_itrs - obs_itrs) itself.  It IS the Cartesian image the update's
// mean_z needs: aer2uvw(ecef2aer(.)) = (r cos el cos az, r cos el sin az, r sin el) with az = atan2(e1, e0),
// el = asin(e2 / r), r = |e| is e again (transformations.py:283-297 after :329-352), so the two sincos of the round trip
// are not evaluated for the 13 measurement sigma points (the identity holds to the rounding of that round trip, 1e-16).
template <bool INL>
SSA_HD void ssa_hx_aer_m(const double* x, const double* M, const double* obs_itrs, const double* T, double* out, double* enz = nullptr) {
  // x_itrs = M @ x[:3]
  double xi[3], d[3], e[3];
  for (int i = 0; i < 3; ++i)
    xi[i] = ssa_fma(M[3 * i + 2], x[2], ssa_fma(M[3 * i + 1], x[1], ssa_mul(M[3 * i], x[0])));
  for (int i = 0; i < 3; ++i) d[i] = xi[i] - obs_itrs[i];
  // R_enz = T^T @ delta
  for (int i = 0; i < 3; ++i)
    e[i] = ssa_fma(T[6 + i], d[2], ssa_fma(T[3 + i], d[1], ssa_mul(T[i], d[0])));
  const double r = ssa_sqrt_t<INL>(ssa_fma(d[2], d[2], ssa_fma(d[1], d[1], ssa_mul(d[0], d[0]))));
  double az = INL ? ssa_atan2_i(e[1], e[0]) : ssa_atan2(e[1], e[0]);
  if (az < 0.0) az = az + SSA_C(TWOPI);
  out[0] = az;
  out[1] = INL ? ssa_asin_t<true>(ssa_div_i(e[2], r)) : ssa_asin(ssa_div(e[2], r));
  out[2] = r;
  if (enz) { enz[0] = e[0]; enz[1] = e[1]; enz[2] = e[2]; }
}
template <bool INL>
SSA_HD void ssa_hx_aer_t(const double* x, const ssa_obs* o, double* out, double* enz = nullptr) { ssa_hx_aer_m<INL>(x, o->M, o->obs_itrs, o->T, out, enz); }
SSA_HD void ssa_hx_aer(const double* x, const ssa_obs* o, double* out, double* enz = nullptr) { ssa_hx_aer_t<false>(x, o, out, enz); }

// Geodetic altitude of an ECEF position (transformations.py:239-279 `ecef2lla`, the closed form of You (2000);
// only the altitude is needed by the catalog generator's 300 km rule, envs/orbit_gen.py:62).  WGS-84.
SSA_HD double ssa_ecef_altitude(const double* ecef) {
  const double a = 6378137.0, f = 1.0 / 298.257223563;
  const double b = ssa_mul(1.0 - f, a);
  const double x = ecef[0], y = ecef[1], z = ecef[2];
  const double r2 = ssa_fma(z, z, ssa_fma(y, y, ssa_mul(x, x)));
  const double E2 = ssa_fma(a, a, -ssa_mul(b, b));
  const double E = ssa_sqrt(E2);
  const double w = r2 - E2;
  const double u = ssa_sqrt(ssa_fma(0.5, w, ssa_mul(0.5, ssa_sqrt(ssa_fma(ssa_mul(4.0, E2), ssa_mul(z, z), ssa_mul(w, w))))));
  const double Q = ssa_sqrt(ssa_fma(y, y, ssa_mul(x, x)));
  const double huE = ssa_sqrt(ssa_fma(u, u, E2));
  double beta;
  if (!(Q == 0.0 || u == 0.0)) beta = ssa_atan(ssa_mul(ssa_div(huE, u), ssa_div(z, Q)));
  else beta = (z >= 0.0) ? SSA_C(PIO2_A) : -SSA_C(PIO2_A);
  double sb, cb;
  ssa_sincos(beta, &sb, &cb);
  const double eps = ssa_div(ssa_mul(ssa_fma(b, u, -ssa_mul(a, huE)) + E2, sb),
                             ssa_fma(ssa_mul(a, huE), ssa_div(1.0, cb), -ssa_mul(E2, cb)));
  beta = beta + eps;
  ssa_sincos(beta, &sb, &cb);
  const double dz = ssa_fma(-b, sb, z), dq = ssa_fma(-a, cb, Q);
  double alt = ssa_sqrt(ssa_fma(dq, dq, ssa_mul(dz, dz)));
  const double inside = ssa_div(ssa_fma(y, y, ssa_mul(x, x)), ssa_mul(a, a)) + ssa_div(ssa_mul(z, z), ssa_mul(b, b));
  if (inside < 1.0) alt = -alt;
  return alt;
}

template <bool INL>
SSA_HD void ssa_aer2uvw_t(const double* aer, double* uvw) {
  const ssa_sc a = INL ? ssa_sincos_i(aer[0]) : ssa_sincos_v(aer[0]);
  const ssa_sc e = INL ? ssa_sincos_i(aer[1]) : ssa_sincos_v(aer[1]);
  const double rc = ssa_mul(aer[2], e.c);
  uvw[0] = ssa_mul(rc, a.c);
  uvw[1] = ssa_mul(rc, a.s);
  uvw[2] = ssa_mul(aer[2], e.s);
}
SSA_HD void ssa_aer2uvw(const double* aer, double* uvw) { ssa_aer2uvw_t<false>(aer, uvw); }

template <bool INL>
SSA_HD void ssa_uvw2aer_t(const double* uvw, double* aer) {
  const double r = ssa_sqrt_t<INL>(ssa_fma(uvw[2], uvw[2], ssa_fma(uvw[1], uvw[1], ssa_mul(uvw[0], uvw[0]))));
  double az = INL ? ssa_atan2_i(uvw[1], uvw[0]) : ssa_atan2(uvw[1], uvw[0]);
  if (az < 0.0) az = az + SSA_C(TWOPI);
  aer[0] = az;
  aer[1] = INL ? ssa_asin_t<true>(ssa_div_i(uvw[2], r)) : ssa_asin(ssa_div(uvw[2], r));
  aer[2] = r;
}
SSA_HD void ssa_uvw2aer(const double* uvw, double* aer) { ssa_uvw2aer_t<false>(uvw, aer); }

// residual_z_aer.  The reference wraps the azimuth difference through atan2(sin d, cos d)
// (dynamics.py:263).  For |d| < pi that expression IS d (to within an ulp of d, and d is the exact value), so
// the wrap — one sincos and one atan2, 13 times per update — is only evaluated when it does something.
SSA_HD void ssa_residual_aer(const double* a, const double* b, double* c) {
  const double d = a[0] - b[0];
  if (ssa_fabs(d) < SSA_C(PI)) {
    c[0] = d;
  } else {
    double s, co;
    ssa_sincos(d, &s, &co);
    c[0] = ssa_atan2(s, co);
  }
  c[1] = a[1] - b[1];
  c[2] = a[2] - b[2];
}
