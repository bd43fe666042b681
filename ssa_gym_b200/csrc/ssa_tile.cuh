// ssa_tile.cuh — the TILE kernels of the step (included by ssa_ukf.cu inside its anonymous namespace).
//
// A CTA owns a tile of T consecutive objects and 14 T threads; everything between the object's state in HBM and its
// next state lives in shared memory:
//   k_predict_tile   (x, U, xt) -> 14 T propagations fx into the shared sigma tile F[78][T] -> unscented transform
//                    spread over the 6 mean and 21 covariance elements -> (xt', x', P')
//   k_refactor       thread / object: robust Cholesky of the predicted P' (filterpy's predict() re-draws sigmas_f, which
//                    calls sqrt_method -> robust_cholesky and may fail the filter), failure sentinels
//   k_update_tile    (x', P', U', xt', z_noise) -> 14 T measurements hx into shared ZS/UVW[39][T] -> mean_z, residuals,
//                    S, Pxz (one thread per row / element, k = 0..12 in order), 3x3 inverse, gain, state and
//                    covariance update, observation / error epilogue
//   k_step_tile      (SSA_UKF_KERNEL=fused) the whole catalog step in one launch: factor -> propagate -> transform ->
//                    factor -> the update body of k_update_tile; tile_robust_chol is the factorisation spread over the
//                    tile's threads, also used by the FACTOR / REFACTOR variants of the two kernels above (tile2).
//                    Bit-identical and measured slower than the four-kernel chain.
// They replace k_fx + k_ut and k_hx + k_update(_staged): the propagated sigma set (624 B / object), the measurement
// sigma set and its Cartesian image (648 B / object) never travel through L2 / HBM, and the per-object linear
// algebra that used to be one ~2 700-instruction chain per thread is spread over the tile's threads.
//
// Thread <-> work mapping.  In the propagation / measurement phase thread t owns (object t / 14, sigma index t % 14;
// 13 = the TRUE state), so a warp holds the sigma points of at most four objects: its lanes run the same eccentricity
// regime and nearly the same Newton trip count (the 13 points of one filter are metres to kilometres apart), where the
// split kernels' warps of 32 unrelated orbits ran every lane for the slowest orbit's count (ncu: 23.5 of 32 lanes
// active).  In the algebra phases thread t owns (element t / T, object t % T): objects are the fastest index of every
// shared array, so all shared accesses are conflict-free and all global accesses are coalesced row segments.
//
// Every output element is produced by the same sequence of rounded operations as in the split kernels and the host
// twin (sums over the 13 sigma points are sequential k = 0..12 everywhere); tests/test_gpu_bitexact.py holds all
// three to bit equality.
#pragma once

#ifndef SSA_TILE
#define SSA_TILE 32     // objects per CTA
#endif
#ifndef SSA_TILE_ROUNDS
#define SSA_TILE_ROUNDS 2  // propagation / measurement tasks per thread: a CTA has 14 * SSA_TILE / SSA_TILE_ROUNDS threads
#endif
#ifndef SSA_LB_PT
#define SSA_LB_PT 4  // resident CTAs per SM the predict tile kernel is compiled for (224 threads: 4 -> 72 registers)
#endif
#ifndef SSA_LB_UTILE
#define SSA_LB_UTILE 5  // 56 registers, 28 bytes of spills: 0.530 vs 0.531 ms at 1 M objects, 73.5 vs 76.2 us at 125 000, 21.0 vs 24.6 us at
#endif                  // 20 000 (625 tiles: one wave of 740 CTA slots instead of 592 + 33).  k_step_tile keeps 4 (its propagation phase).
constexpr int kTileThreads = 14 * SSA_TILE / SSA_TILE_ROUNDS;
static_assert((14 * SSA_TILE) % SSA_TILE_ROUNDS == 0 && kTileThreads % SSA_TILE == 0 && kTileThreads % 32 == 0 && kTileThreads % 14 == 0, "tile shape");

// FACTOR: the kernel factors (lambda + n) P itself (k_factor folded in): rows xt 0..5 | x 6..11 | P 12..32 | U 33..53
template <int T, bool FACTOR>
struct PredictTile {
  alignas(128) double S[FACTOR ? 54 : 33][T];  // staged state rows: xt 0..5 | x 6..11 | U 12..32 (T = 32: two TMA tile loads)
  double F[78][T + 1];  // propagated sigma points, then deviations; +1: the 14 sigma lanes of one object write rows 6 apart
  uint64_t bar;
  int live[T], exc[T], nan[T], texc[T];
  int res[T], bad[T], fcode[T];  // FACTOR
};
constexpr int TS_XT = 0, TS_X = 6, TS_U = 12;

constexpr int CHOL_PENDING = -2;
template <int T, int NT>
__device__ __forceinline__ void tile_robust_chol(const double (*A)[T], double (*U)[T], const double lam, int* res, int* bad,
                                                 const int r_a, const int o_a);

// sigma point k (0..12) of the object in column o of a tile: s = x +- U[r, :], r = (k - 1) % 6 (rows of the upper factor).
// x = rows of the mean, U = the 21 packed rows of the factor (row-major upper triangle), both [row][T].
template <int T>
__device__ __forceinline__ void tile_sigma(const double (*x)[T], const double (*U)[T], int o, int k, double* s) {
  const int r = (k == 0) ? -1 : (k - 1) % 6;
  const int rb = r * 5 - (r * (r - 1)) / 2;  // packed index of U[r][j] is rb + j (j >= r)
  const bool minus = k > 6;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double u = 0.0;
    if (r >= 0 && r <= j) u = U[rb + j][o];
    s[j] = minus ? (x[j][o] - u) : (x[j][o] + u);
  }
}

// covariance rows I and 5 - I of one object (7 elements): acc[0 .. 5-I] = (I, I..5), acc[6-I .. 6] = (5-I, 5-I..5)
template <int I, int T>
__device__ __forceinline__ void tile_cov_rows(const double (*F)[T + 1], int o, const double* Wc, double* acc) {
  constexpr int I2 = 5 - I;
#pragma unroll
  for (int q = 0; q < 7; ++q) acc[q] = 0.0;
#pragma unroll
  for (int k = 0; k < SSA_NSIG; ++k) {
    double y[6];
#pragma unroll
    for (int j = I; j < 6; ++j) y[j] = F[k * 6 + j][o];
    const double wk = Wc[k];
#pragma unroll
    for (int j = I; j < 6; ++j) {
      const double m = ssa_mul(wk, y[j]);
      acc[j - I] = ssa_fma(y[I], m, acc[j - I]);
      if (j >= I2) acc[(6 - I) + (j - I2)] = ssa_fma(y[I2], m, acc[(6 - I) + (j - I2)]);
    }
  }
}

// FACTOR = false: the factor U comes from k_factor (tm_x = box of xt | x, tm_u = box of U).  FACTOR = true: the kernel
// stages xt | x | P (tm_x = box of all 33 state rows) and factors (lambda + n) P across the tile's threads first —
// k_factor folded in, the factor never travels through HBM.
// MINB = resident CTAs per SM the kernel is compiled for: 4 (72 registers) everywhere except for catalogs whose tiles fill
// one wave only at 5 CTAs per SM (56 registers, spills in the propagation: 0.689 vs 0.677 ms at 1 M objects, but 27.6 vs
// 29.3 us at 20 000 objects = 625 tiles).
template <int T, int NT, bool FACTOR, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_predict_tile(const KParams p, const __grid_constant__ CUtensorMap tm_x,
                                                                const __grid_constant__ CUtensorMap tm_u) {
  constexpr int NR = NT / T;  // rows of the algebra mapping
  static_assert(NR >= 7, "the algebra phases use up to 7 thread rows");
  static_assert(!FACTOR || T == 32, "the factoring variant stages its tile with one TMA load");
  constexpr int UO = FACTOR ? 33 : TS_U;  // first row of the factor
  __shared__ PredictTile<T, FACTOR> sm;
  pdl_prologue();
  const int tid = threadIdx.x;
  const int o_a = tid % T, r_a = tid / T;  // algebra mapping: (row / element r_a in 0..NR-1, object o_a)
  const long loc0 = (long)blockIdx.x * T;
  const long ld = p.ld, lds = p.lds;
  const bool predict = (p.flags & SSA_STEP_PREDICT) != 0, truth = (p.flags & SSA_STEP_TRUTH) != 0;
  const long loc_a = loc0 + o_a;
  const bool valid_a = loc_a < p.Nc;
  const long obj_a = p.obj0 + loc_a;

  // ---- stage the tile: 33 rows of T consecutive objects ----
  if (T == 32) {  // two TMA tile loads (box 32 x 12 of the state tensor, 32 x 21 of the scratch tensor), one mbarrier
    if (tid == 0) {
      mbar_init(&sm.bar, 1);
      mbar_expect_tx(&sm.bar, 33 * 256);
      tma_load_2d(&sm.S[0][0], &tm_x, (int)(p.obj0 + loc0), 0, &sm.bar);  // FACTOR: all 33 rows xt | x | P
      if (!FACTOR) tma_load_2d(&sm.S[12][0], &tm_u, (int)loc0, SC_U, &sm.bar);
    }
  } else if (valid_a) {  // thread (row, object): coalesced row segments
#pragma unroll
    for (int row = r_a; row < 33; row += NR) {
      const double* src = row < 6 ? p.xt + row * ld + obj_a : (row < 12 ? p.x + (row - 6) * ld + obj_a : p.U + (row - 12) * lds + loc_a);
      sm.S[row][o_a] = *src;
    }
  }
  if (tid < T) {
    int live = 0;
    if (valid_a && predict) live = !(p.status[obj_a] & SSA_ST_FAILED) && (FACTOR || !p.code[obj_a]);
    sm.live[tid] = live; sm.exc[tid] = 0; sm.nan[tid] = 0; sm.texc[tid] = 0;
    if (FACTOR) { sm.res[tid] = live ? CHOL_PENDING : 0; sm.bad[tid] = 0; sm.fcode[tid] = 0; }
  }
  __syncthreads();
  if (T == 32) mbar_wait(&sm.bar, 0);
  if (FACTOR) {  // sigma points of the prior: U^T U = (lambda + n) P with the inflation fallback (what k_factor does)
    tile_robust_chol<T, NT>(sm.S + 12, sm.S + UO, p.lam, sm.res, sm.bad, r_a, o_a);
    if (tid < T && sm.live[tid]) {
      const int r1 = sm.res[tid];
      if (r1 == CHOL_PENDING) { sm.fcode[tid] = SSA_ST_LINALG; sm.live[tid] = 0; }
      else if (r1 > 0) p.infl[obj_a] += 1;
    }
    __syncthreads();
  }

  // ---- 14 T propagations: task = (object, sigma index), 13 = the TRUE state; ONE call site of fx ----
#pragma unroll 1
  for (int o = tid / 14; o < T; o += NT / 14) {  // NT is a multiple of 14: a thread keeps its sigma index in every round
    const int k = tid % 14;
    const bool is_truth = (k == 13);
    const bool run = (loc0 + o < p.Nc) && (is_truth ? truth : (sm.live[o] != 0));
    if (run) {
      double s[6], f[6];
      tile_sigma<T>(sm.S + TS_X, sm.S + UO, o, is_truth ? 0 : k, s);
      if (is_truth) {
#pragma unroll
        for (int j = 0; j < 6; ++j) s[j] = sm.S[TS_XT + j][o];
      }
      const int exc = ssa_fx(s, p.dt, f);
      if (is_truth) {
#pragma unroll
        for (int j = 0; j < 6; ++j) sm.S[TS_XT + j][o] = f[j];
        if (exc) sm.texc[o] = 1;
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) sm.F[k * 6 + j][o] = f[j];
        if (exc) sm.exc[o] = 1;
      }
    }
  }
  __syncthreads();

  // ---- unscented transform.  Thread (component i, object): mean over k = 0..12 in order, then the 13 deviations of
  //      that component in place (no other thread touches them); the remaining threads write the true state back ----
  const bool ut_a = valid_a && sm.live[o_a] && !sm.exc[o_a];
  if (r_a < 6) {
    if (ut_a) {
      double f[SSA_NSIG];
#pragma unroll
      for (int k = 0; k < SSA_NSIG; ++k) f[k] = sm.F[k * 6 + r_a][o_a];
      double acc = ssa_mul(p.Wm[0], f[0]);
#pragma unroll
      for (int k = 1; k < SSA_NSIG; ++k) acc = ssa_fma(p.Wm[k], f[k], acc);
      if (ssa_isnan(acc)) sm.nan[o_a] = 1;
      p.x[r_a * ld + obj_a] = acc;  // a failure of this predict overwrites it with the sentinel in k_refactor
#pragma unroll
      for (int k = 0; k < SSA_NSIG; ++k) sm.F[k * 6 + r_a][o_a] = f[k] - acc;
    }
  } else if (valid_a && truth) {
#pragma unroll
    for (int row = r_a - 6; row < 6; row += NR - 6) p.xt[row * ld + obj_a] = sm.S[TS_XT + row][o_a];
  }
  __syncthreads();
  // covariance + Q: thread (row pair {i, 5 - i}, object) owns the 7 elements (i, i..5) and (5-i, 5-i..5), each summed
  // over k = 0..12 in order.  (T = 16: a warp holds two values of r_a; the three row pairs go to three different warps so
  // that none diverges.)
  const bool cov_thread = (T >= 32) ? (r_a < 3) : (r_a < 6 && !(r_a & 1));
  if (cov_thread) {
    if (ut_a) {
      const int i = (T >= 32) ? r_a : (r_a >> 1), i2 = 5 - i;
      double acc[7];
      if (i == 0) tile_cov_rows<0, T>(sm.F, o_a, p.Wc, acc);
      else if (i == 1) tile_cov_rows<1, T>(sm.F, o_a, p.Wc, acc);
      else tile_cov_rows<2, T>(sm.F, o_a, p.Wc, acc);
      const int b1 = i * 6 - (i * (i - 1)) / 2, n1 = 6 - i;  // packed index of (i, j): i*6 - i(i-1)/2 + (j - i)
      const int b2 = i2 * 6 - (i2 * (i2 - 1)) / 2;
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int e = (q < n1) ? (b1 + q) : (b2 + q - n1);
        p.P[e * ld + obj_a] = acc[q] + __ldg(p.qr + e);
      }
    }
  } else if (r_a == NR - 1) {  // per-object words: truth exception, failure code of this predict (k_refactor applies it)
    if (valid_a) {
      if (sm.texc[o_a]) p.status[obj_a] = p.status[obj_a] | SSA_ST_TRUTHEXC;
      if (sm.live[o_a]) p.code[obj_a] = sm.exc[o_a] ? SSA_ST_FXEXC : (sm.nan[o_a] ? SSA_ST_NAN : 0);
      else if (FACTOR) p.code[obj_a] = sm.fcode[o_a];  // k_factor's word: LinAlgError of the prior's factorisation, or 0
    }
  }
}

// After the unscented transform: factor the predicted covariance (the re-drawn sigma points of filterpy's predict();
// their factor is what the update reads), apply the failure of this predict to the state.  Thread per object.
__global__ void __launch_bounds__(kObjThreads, SSA_LB_FAC * 128 / kObjThreads) k_refactor(const KParams p) {
  pdl_prologue();
  const long loc = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (loc >= p.Nc) return;
  const long obj = p.obj0 + loc;
  const int st = p.status[obj];
  if (st & SSA_ST_FAILED) return;
  int code = p.code[obj];
  if ((code == 0 || code == SSA_ST_NAN) && p.resample) {
    double U[SSA_NP];
    const int r2 = ssa_robust_chol6_t<SSA_CHOL_INLINE>(p.P + obj, p.ld, p.lam, U);
    if (r2 < 0) code |= SSA_ST_LINALG;
    else {
      if (r2 > 0) p.infl[obj] += 1;
#pragma unroll
      for (int e = 0; e < SSA_NP; ++e) p.U[e * p.lds + loc] = U[e];
    }
  }
  if (code) {
    store_sentinel(p, obj);
    p.status[obj] = st | SSA_ST_FAILED | code;
  }
  p.code[obj] = code;
}

template <int T>
struct UpdateTile {
  alignas(128) double S[54][T];  // staged rows: xt 0..5 | x 6..11 | P 12..32 | U 33..53 (T = 32: two TMA tile loads)
  double zn[3][T];
  double ZS[39][T + 1];   // measurement sigma points (az, el, range); after the residuals: Pxz 18 | Sm 9 rows of [T]
  double UVW[39][T + 1];  // their Cartesian images, then the 13 residuals; after S / Pxz: K 18 | Tm 18 rows of [T]
  double zt[3][T];        // measurement of the TRUE state
  double zm[3][T];        // Cartesian mean
  double zp[3][T];        // predicted measurement
  double yr[3][T];        // innovation
  double SI[9][T];
  double xn[6][T];
  uint64_t bar;
  int st[T], code[T], vis[T], ok[T], nan[T];
  int want[T];  // the object takes the update this step: every object (SSA_STEP_UPDATE_ALL) or the tasked one of its environment (_ACT)
  int live0[T], live[T], exc[T], texc[T], pnan[T], res[T], bad[T];  // predict phases of the fused kernel k_step_tile
};

// Row I of the cross covariance Pxz = sum_k Wc_k dx_k r_k^T of one object.  dx_k[I] = (x_I +- U[r][I]) - x_I is exactly
// zero unless r = (k - 1) % 6 <= I (U is upper triangular), and a zero term leaves the running sum unchanged bit for bit,
// so only k = 1..I+1 and 7..I+7 are evaluated, in increasing k like the full sum.
template <int I, int T>
__device__ __forceinline__ void tile_pxz_row(const double (*S)[T], const double (*RZ)[T + 1], int o, const double* Wc, double* a) {
  const double xi = S[6 + I][o];  // US_X + I
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int r = 0; r <= I; ++r) {
      const int k = 1 + r + 6 * half;
      const double u = S[33 + r * 5 - (r * (r - 1)) / 2 + I][o];  // US_U + packed index of U[r][I]
      const double sk = half ? (xi - u) : (xi + u);
      const double dx = sk - xi;
      const double wk = Wc[k];
      a0 = ssa_fma(wk, ssa_mul(dx, RZ[k * 3 + 0][o]), a0);
      a1 = ssa_fma(wk, ssa_mul(dx, RZ[k * 3 + 1][o]), a1);
      a2 = ssa_fma(wk, ssa_mul(dx, RZ[k * 3 + 2][o]), a2);
    }
  }
  a[0] = a0; a[1] = a1; a[2] = a2;
}
constexpr int US_XT = 0, US_X = 6, US_P = 12, US_U = 33;

// Everything of the update after the tile is staged (sm.S rows xt | x | P | U, sm.zn, sm.st, sm.code; sm.ok = 1,
// sm.nan = sm.vis = 0): 14 T measurements, mean_z, residuals, S, Pxz, 3x3 inverse, gain, state / covariance update,
// write-back, epilogue.  FUSED (k_step_tile): the tile arrives from the predict phases of the same kernel — every object
// that was not failed when the step began (sm.live0) writes its mean and covariance back, updated or not.
template <int T, int NT, bool FUSED>
__device__ __forceinline__ void tile_update_body(const KParams& p, UpdateTile<T>& sm, const long loc0) {
  constexpr int NR = NT / T;  // rows of the algebra mapping
  double(*sPxz)[T] = reinterpret_cast<double(*)[T]>(&sm.ZS[0][0]);  // [18][T], valid after the residual phase
  double(*sSm)[T] = sPxz + 18;                                      // [9][T]
  double(*sK)[T] = reinterpret_cast<double(*)[T]>(&sm.UVW[0][0]);   // [18][T], valid after the S / Pxz phase
  double(*sTm)[T] = sK + 18;                                        // [18][T]
  static_assert(27 * T <= 39 * (T + 1) && 36 * T <= 39 * (T + 1), "aliases fit");
  const int tid = threadIdx.x;
  const int o_a = tid % T, r_a = tid / T;  // algebra mapping: (row / element r_a in 0..NR-1, object o_a)
  const long ld = p.ld;
  const int flags = p.flags;
  const bool aer = (p.obs_type == SSA_OBS_AER);
  const long loc_a = loc0 + o_a;
  const bool valid_a = loc_a < p.Nc;
  const long obj_a = p.obj0 + loc_a;

  // ---- 14 T measurements: task = (object, sigma index); 13 = the TRUE state (visibility, z_true) ----
#pragma unroll 1
  for (int o = tid / 14; o < T; o += NT / 14) {  // NT is a multiple of 14: a thread keeps its sigma index in every round
    const int k = tid % 14;
    const long loc = loc0 + o;
    const bool is_truth = (k == 13);
    const bool live = !(sm.st[o] & SSA_ST_FAILED) && !sm.code[o] && sm.want[o];
    if (loc < p.Nc && (is_truth || live)) {
      const long obj = p.obj0 + loc;
      double s[6], z[3], uvw[3];
      tile_sigma<T>(sm.S + US_X, sm.S + US_U, o, is_truth ? 0 : k, s);
      if (is_truth) {
#pragma unroll
        for (int j = 0; j < 3; ++j) s[j] = sm.S[US_XT + j][o];
      }
      if (aer || is_truth) {
        if (p.Menv) {  // device-resident trans_matrix (per environment / per step): fetched once into registers
          double M[9];
          const double* Mg = env_M(p, obj / p.m);
#pragma unroll
          for (int i = 0; i < 9; ++i) M[i] = Mg[i];
          ssa_hx_aer_m<true>(s, M, p.ob.obs_itrs, p.ob.T, z, uvw);
        } else {
          ssa_hx_aer_m<true>(s, p.ob.M, p.ob.obs_itrs, p.ob.T, z, uvw);
        }
      }
      if (is_truth) {
        const int visible = z[1] >= p.obs_limit;  // SS2:424
        p.visible[obj] = (uint8_t)visible;
        sm.vis[o] = visible;
#pragma unroll
        for (int a = 0; a < 3; ++a) sm.zt[a][o] = aer ? z[a] : s[a];
      } else if (aer) {  // the Cartesian image of the measurement is the topocentric vector itself (ssa_meas.h)
#pragma unroll
        for (int a = 0; a < 3; ++a) { sm.ZS[k * 3 + a][o] = z[a]; sm.UVW[k * 3 + a][o] = uvw[a]; }
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) sm.ZS[k * 3 + a][o] = s[a];
      }
    }
  }
  __syncthreads();

  // the object of this thread's algebra column takes the update: wanted, not failed, factor available, visible
  const bool upd_a = valid_a && !(sm.st[o_a] & SSA_ST_FAILED) && !sm.code[o_a] && sm.vis[o_a] && sm.want[o_a];
  // ---- mean_z: Cartesian (uvw) mean of the angular measurements, or the plain mean (xyz) ----
  if (r_a < 3) {
    if (upd_a) {
      const double(*V)[T + 1] = aer ? sm.UVW : sm.ZS;
      double acc = ssa_mul(p.Wm[0], V[r_a][o_a]);
#pragma unroll
      for (int k = 1; k < SSA_NSIG; ++k) acc = ssa_fma(p.Wm[k], V[k * 3 + r_a][o_a], acc);
      sm.zm[r_a][o_a] = acc;
    }
  } else if (r_a == 3) {
    if (valid_a && !(sm.st[o_a] & SSA_ST_FAILED) && sm.want[o_a] && p.z_true) {
#pragma unroll
      for (int a = 0; a < 3; ++a) p.z_true[obj_a * 3 + a] = sm.zt[a][o_a];  // SS2:298
    }
  }
  __syncthreads();
  if (r_a == 0 && upd_a) {
    double zm[3], zp[3], z[3], yr[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { zm[a] = sm.zm[a][o_a]; z[a] = sm.zt[a][o_a] + sm.zn[a][o_a]; }
    if (aer) {
      ssa_uvw2aer_t<true>(zm, zp);
      ssa_residual_aer(z, zp, yr);
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) { zp[a] = zm[a]; yr[a] = z[a] - zp[a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { sm.zp[a][o_a] = zp[a]; sm.yr[a][o_a] = yr[a]; }
  }
  __syncthreads();
  // ---- the 13 residuals: thread (sigma index, object); they replace the Cartesian images ----
  if (upd_a) {
#pragma unroll
    for (int k = r_a; k < SSA_NSIG; k += NR) {
      double zk[3], zp[3], rz[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { zk[a] = sm.ZS[k * 3 + a][o_a]; zp[a] = sm.zp[a][o_a]; }
      if (aer) ssa_residual_aer(zk, zp, rz);
      else {
#pragma unroll
        for (int a = 0; a < 3; ++a) rz[a] = zk[a] - zp[a];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) sm.UVW[k * 3 + a][o_a] = rz[a];
      if (p.sigmas_h) {
#pragma unroll
        for (int a = 0; a < 3; ++a) p.sigmas_h[obj_a * 39 + k * 3 + a] = zk[a];
      }
    }
  }
  __syncthreads();
  // ---- cross covariance (thread per state row: dx_k formed once, three accumulators) and innovation covariance
  //      (thread per element); k = 0..12 in order.  Pxz / Sm overwrite the (dead) measurement sigma points ----
  if (upd_a) {
    const int nS = aer ? 6 : 9;  // aer: the reference's loop makes S symmetric bit for bit; xyz: np.dot form, 9 elements
    for (int task = r_a; task < 6 + nS; task += NR) {
      if (task < 6) {
        double a3[3];
        switch (task) {
          case 0: tile_pxz_row<0, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
          case 1: tile_pxz_row<1, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
          case 2: tile_pxz_row<2, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
          case 3: tile_pxz_row<3, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
          case 4: tile_pxz_row<4, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
          default: tile_pxz_row<5, T>(sm.S, sm.UVW, o_a, p.Wc, a3); break;
        }
        sPxz[3 * task + 0][o_a] = a3[0]; sPxz[3 * task + 1][o_a] = a3[1]; sPxz[3 * task + 2][o_a] = a3[2];
      } else if (aer) {
        const int e = task - 6;
        const int a = c_sa[e], b = c_sb[e];
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < SSA_NSIG; ++k) acc = ssa_fma(p.Wc[k], ssa_mul(sm.UVW[k * 3 + a][o_a], sm.UVW[k * 3 + b][o_a]), acc);
        sSm[3 * a + b][o_a] = acc + __ldg(p.qr + 21 + 3 * a + b);
        sSm[3 * b + a][o_a] = acc + __ldg(p.qr + 21 + 3 * b + a);
      } else {
        const int e = task - 6;
        const int a = e / 3, b = e % 3;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < SSA_NSIG; ++k) acc = ssa_fma(sm.UVW[k * 3 + a][o_a], ssa_mul(p.Wc[k], sm.UVW[k * 3 + b][o_a]), acc);
        sSm[e][o_a] = acc + __ldg(p.qr + 21 + e);
      }
    }
  }
  __syncthreads();
  // ---- S^-1 (numpy.linalg.inv: LU with partial pivoting): thread per object ----
  if (r_a == 0 && upd_a) {
    double Sm[9], SI[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Sm[e] = sSm[e][o_a];
    sm.ok[o_a] = ssa_inv3_t<true>(Sm, SI);
#pragma unroll
    for (int e = 0; e < 9; ++e) sm.SI[e][o_a] = SI[e];
  } else if (r_a == 1 && upd_a) {
    if (p.y) {
#pragma unroll
      for (int a = 0; a < 3; ++a) p.y[obj_a * 3 + a] = sm.yr[a][o_a];
    }
    if (p.S) {
#pragma unroll
      for (int e = 0; e < 9; ++e) p.S[obj_a * 9 + e] = sSm[e][o_a];
    }
  }
  __syncthreads();
  // ---- gain row, S K^T column, new mean: thread (state row, object).  K / Tm overwrite the (dead) residuals ----
  if (r_a < 6 && upd_a) {
    const int i = r_a;
    double K[3];
    const double px0 = sPxz[3 * i][o_a], px1 = sPxz[3 * i + 1][o_a], px2 = sPxz[3 * i + 2][o_a];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      K[a] = ssa_fma(px2, sm.SI[6 + a][o_a], ssa_fma(px1, sm.SI[3 + a][o_a], ssa_mul(px0, sm.SI[a][o_a])));
      sK[3 * i + a][o_a] = K[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
      sTm[a * 6 + i][o_a] = ssa_fma(sSm[3 * a + 2][o_a], K[2], ssa_fma(sSm[3 * a + 1][o_a], K[1], ssa_mul(sSm[3 * a][o_a], K[0])));
    const double xn = sm.S[US_X + i][o_a] + ssa_fma(K[2], sm.yr[2][o_a], ssa_fma(K[1], sm.yr[1][o_a], ssa_mul(K[0], sm.yr[0][o_a])));
    sm.xn[i][o_a] = xn;
    if (ssa_isnan(xn)) sm.nan[o_a] = 1;
  }
  __syncthreads();
  // ---- P -= K (S K^T) and the new mean go back to HBM and into the tile (the epilogue reads them): thread (element,
  //      object), elements 0..20 of P then 21..26 = the mean.  A failure of this update (or of the stand-alone
  //      factorisation before it) stores the sentinels instead, SS2:369-382 ----
  const bool alive_a = valid_a && !(sm.st[o_a] & SSA_ST_FAILED) && sm.want[o_a];  // (objects that do not take the update only get the epilogue)
  const int fail_a = !alive_a ? 0
                              : (sm.code[o_a] ? sm.code[o_a]
                                              : (sm.vis[o_a] ? (!sm.ok[o_a] ? (SSA_ST_LINALG | SSA_ST_IN_UPDATE)
                                                                            : (sm.nan[o_a] ? (SSA_ST_NAN | SSA_ST_IN_UPDATE) : 0))
                                                             : 0));
  // (not FUSED: an object that is written and did not fail took the update; FUSED: predicted-only objects — not visible —
  // and the sentinels of a failed predict go back as they stand in the tile)
  const bool changed_a = FUSED ? (valid_a && sm.live0[o_a] != 0) : (alive_a && (fail_a || sm.vis[o_a]));
  const bool took_a = !FUSED || upd_a;
  if (changed_a) {
#pragma unroll
    for (int e = r_a; e < SSA_NP + 6; e += NR) {
      if (e < SSA_NP) {
        const int i = c_pi[e], j = c_pj[e];
        double pn;
        if (fail_a) pn = (i == j) ? (i < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL) : 0.0;
        else if (took_a) {
          const double kt = ssa_fma(sK[3 * i + 2][o_a], sTm[12 + j][o_a],
                                    ssa_fma(sK[3 * i + 1][o_a], sTm[6 + j][o_a], ssa_mul(sK[3 * i][o_a], sTm[j][o_a])));
          pn = sm.S[US_P + e][o_a] - kt;
        } else pn = sm.S[US_P + e][o_a];
        sm.S[US_P + e][o_a] = pn;
        p.P[e * ld + obj_a] = pn;
      } else {
        const int i = e - SSA_NP;
        const double xn = fail_a ? (i < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL) : (took_a ? sm.xn[i][o_a] : sm.S[US_X + i][o_a]);
        sm.S[US_X + i][o_a] = xn;  // nobody reads the prior mean any more
        p.x[i * ld + obj_a] = xn;
      }
    }
  }
  if (r_a == NR - 1 && valid_a) {  // per-object words
    int st = sm.st[o_a];
    if (fail_a) { st |= SSA_ST_FAILED | fail_a; p.status[obj_a] = st; }
    if (p.updated) p.updated[obj_a] = (uint8_t)(alive_a && !sm.code[o_a] && sm.vis[o_a]);
    if (p.status_out) p.status_out[obj_a] = st;
  }
  if (!(flags & SSA_STEP_EPILOGUE)) return;
  __syncthreads();
  // ---- epilogue (results.py:36-72): obs row, errors, trace ----
  for (int idx = tid; idx < 12 * T; idx += NT) {  // obs is [N][12]: the tile's 12 T doubles are contiguous
    const int o = idx / 12, c = idx % 12;
    if (loc0 + o < p.Nc)
      p.obs[(p.obj0 + loc0) * 12 + idx] = (c < 6) ? sm.S[US_X + c][o] : sm.S[US_P + ssa_pidx(c - 6, c - 6)][o];
  }
  if (valid_a) {
    if (r_a == NR - 1) {
      const double d0 = sm.S[US_X + 0][o_a] - sm.S[US_XT + 0][o_a], d1 = sm.S[US_X + 1][o_a] - sm.S[US_XT + 1][o_a];
      const double d2 = sm.S[US_X + 2][o_a] - sm.S[US_XT + 2][o_a], d3 = sm.S[US_X + 3][o_a] - sm.S[US_XT + 3][o_a];
      const double d4 = sm.S[US_X + 4][o_a] - sm.S[US_XT + 4][o_a], d5 = sm.S[US_X + 5][o_a] - sm.S[US_XT + 5][o_a];
      p.dpos[obj_a] = ssa_sqrt_t<true>(ssa_fma(d2, d2, ssa_fma(d1, d1, ssa_mul(d0, d0))));
      p.dvel[obj_a] = ssa_sqrt_t<true>(ssa_fma(d5, d5, ssa_fma(d4, d4, ssa_mul(d3, d3))));
    } else if (r_a == NR - 2) {
      const double g0 = sm.S[US_P + 0][o_a], g1 = sm.S[US_P + 6][o_a], g2 = sm.S[US_P + 11][o_a];
      const double g3 = sm.S[US_P + 15][o_a], g4 = sm.S[US_P + 18][o_a], g5 = sm.S[US_P + 20][o_a];
      p.spos[obj_a] = ssa_sqrt_t<true>((g0 + g1) + g2);
      p.svel[obj_a] = ssa_sqrt_t<true>((g3 + g4) + g5);
      p.trace[obj_a] = ((((g0 + g1) + g2) + g3) + g4) + g5;
    }
  }
}

// REFACTOR = false: the factor of the predicted covariance comes from k_refactor (tm_u).  REFACTOR = true: the kernel
// stages xt | x | P only, factors the predicted covariance across the tile's threads and applies the failure of the
// predict (sentinels, status) — k_refactor folded in, the factor never travels through HBM.
template <int T, int NT, bool REFACTOR>
__global__ void __launch_bounds__(NT, SSA_LB_UTILE) k_update_tile(const KParams p, const __grid_constant__ CUtensorMap tm_s,
                                                                  const __grid_constant__ CUtensorMap tm_u) {
  constexpr int NR = NT / T;  // rows of the algebra mapping
  static_assert(NR >= 7, "the algebra phases use up to 7 thread rows");
  static_assert(!REFACTOR || T == 32, "the factoring variant stages its tile with one TMA load");
  extern __shared__ __align__(128) unsigned char tile_raw[];
  UpdateTile<T>& sm = *reinterpret_cast<UpdateTile<T>*>(tile_raw);
  pdl_prologue();
  const int tid = threadIdx.x;
  const int o_a = tid % T, r_a = tid / T;  // algebra mapping: (row / element r_a in 0..NR-1, object o_a)
  const long loc0 = (long)blockIdx.x * T;
  const long ld = p.ld, lds = p.lds;
  const long loc_a = loc0 + o_a;
  const bool valid_a = loc_a < p.Nc;
  const long obj_a = p.obj0 + loc_a;

  // ---- stage the tile: 54 rows (xt 6 | x 6 | P 21 | U 21) + z_noise ----
  if (T == 32) {  // two TMA tile loads (box 32 x 33 of the state tensor, 32 x 21 of the scratch tensor), one mbarrier
    if (tid == 0) {
      mbar_init(&sm.bar, 1);
      mbar_expect_tx(&sm.bar, (REFACTOR ? 33 : 54) * 256);
      tma_load_2d(&sm.S[0][0], &tm_s, (int)(p.obj0 + loc0), 0, &sm.bar);
      if (!REFACTOR) tma_load_2d(&sm.S[33][0], &tm_u, (int)loc0, SC_U, &sm.bar);
    }
  } else if (valid_a) {  // thread (row, object): coalesced row segments
#pragma unroll
    for (int row = r_a; row < 54; row += NR) {
      const double* src = row < 6 ? p.xt + row * ld + obj_a
                                  : (row < 12 ? p.x + (row - 6) * ld + obj_a
                                              : (row < 33 ? p.P + (row - 12) * ld + obj_a : p.U + (row - 33) * lds + loc_a));
      sm.S[row][o_a] = *src;
    }
  }
  if (tid < 3 * T) {  // z_noise is [N][3]: the tile's 3 T doubles are contiguous
    const long loc = loc0 + tid / 3;
    sm.zn[tid % 3][tid / 3] = (p.z_noise && loc < p.Nc) ? p.z_noise[(p.obj0 + loc0) * 3 + tid] : 0.0;
  }
  int st_r = SSA_ST_FAILED, code_r = 0;  // threads 3 T .. 4 T - 1 keep the words of object tid - 3 T
  if (tid >= 3 * T && tid < 4 * T) {
    const int o = tid - 3 * T;
    int want = 0;
    if (loc0 + o < p.Nc) {
      const long obj = p.obj0 + loc0 + o;
      st_r = p.status[obj]; code_r = p.code[obj];
      want = (p.flags & SSA_STEP_UPDATE_ALL) ? 1 : (((p.flags & SSA_STEP_UPDATE_ACT) && p.actions[obj / p.m] == (int)(obj % p.m)) ? 1 : 0);
    }
    sm.st[o] = st_r; sm.code[o] = code_r; sm.ok[o] = 1; sm.nan[o] = 0; sm.vis[o] = 0; sm.want[o] = want;
    if (REFACTOR) {
      sm.res[o] = (!(st_r & SSA_ST_FAILED) && (code_r == 0 || code_r == SSA_ST_NAN)) ? CHOL_PENDING : 0;
      sm.bad[o] = 0;
    }
  }
  __syncthreads();
  if (T == 32) mbar_wait(&sm.bar, 0);
  if (REFACTOR) {  // the re-drawn sigma points of filterpy's predict(), failure of the predict (what k_refactor does)
    tile_robust_chol<T, NT>(sm.S + US_P, sm.S + US_U, p.lam, sm.res, sm.bad, r_a, o_a);
    if (tid >= 3 * T && tid < 4 * T && !(st_r & SSA_ST_FAILED)) {
      const int o = tid - 3 * T;
      const long obj = p.obj0 + loc0 + o;
      if (code_r == 0 || code_r == SSA_ST_NAN) {
        const int r2 = sm.res[o];
        if (r2 == CHOL_PENDING) code_r |= SSA_ST_LINALG;
        else if (r2 > 0) p.infl[obj] += 1;
      }
      if (code_r) {
        store_sentinel(p, obj);
#pragma unroll
        for (int i = 0; i < 6; ++i) sm.S[US_X + i][o] = i < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL;
#pragma unroll
        for (int e = 0; e < SSA_NP; ++e) sm.S[US_P + e][o] = (c_pi[e] == c_pj[e]) ? (c_pi[e] < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL) : 0.0;
        st_r |= SSA_ST_FAILED | code_r;
        p.status[obj] = st_r;
        sm.st[o] = st_r;
      }
      p.code[obj] = code_r;
      sm.code[o] = code_r;
    }
    __syncthreads();
  }

  tile_update_body<T, NT, false>(p, sm, loc0);
}

// ---- the FUSED step kernel (catalog mode: truth + predict + update of every object + epilogue in ONE launch) ----------
//
// Robust Cholesky of lam * A for the T objects of a tile, spread over the tile's threads: the right-looking elimination of
// ssa_chol6_t with row r of the trailing matrix owned by thread row r (objects across lanes), one barrier per column.
// Every element sees the same sequence of rounded operations as in the one-thread routine (column j: inv = 1 / d_j,
// l = a[j][r] * inv, a[r][c] = fma(-l, a[j][c], a[r][c]); at the end row j is scaled by inv and then by sqrt(d_j)), so
// the factor is bit-identical to k_factor's / k_refactor's; the per-thread chain of a column is one division and at most
// five FMAs, and the 400-instruction dependent chain that made the factorisation a kernel of its own (one warp of seven
// busy when folded into a tile kernel as it stood) is spread over six thread rows.
// The inflation retries +10^i, i = -6..9 (dynamics.py:402-417) loop over the whole tile while any of its objects still
// has a non-positive pivot; objects that already have their factor sit the retries out.
// res[o]: in CHOL_PENDING for the objects to factor (anything else: skipped), out the number of the successful attempt
// (0 plain, t + 1 with +10^(t-6)) or CHOL_PENDING when all 17 attempts failed.  bad[] must be zero on entry.
template <int T, int NT>
__device__ __forceinline__ void tile_robust_chol(const double (*A)[T], double (*U)[T], const double lam, int* res, int* bad,
                                                 const int r_a, const int o_a) {
  constexpr int NR = NT / T;
#pragma unroll 1
  for (int t = -1; t < 16; ++t) {
    const bool need = res[o_a] == CHOL_PENDING;
    if (need) {
      const double eps = (t >= 0) ? ssa_pow10_infl(t) : 0.0;
      int fin = 1;
#pragma unroll
      for (int e = r_a; e < SSA_NP; e += NR) {
        double u = ssa_mul(lam, A[e][o_a]);
        if (t >= 0 && c_pi[e] == c_pj[e]) u = u + eps;
        fin &= (ssa_fabs(u) <= 1.79769313486231570815e+308);
        U[e][o_a] = u;
      }
      if (!fin) bad[o_a] = 1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (need && r_a > j && r_a < 6) {
        const int r = r_a;
        const int bj = j * 6 - (j * (j - 1)) / 2 - j;  // packed index of (j, c) is bj + c
        const int br = r * 6 - (r * (r - 1)) / 2 - r;
        const double inv = ssa_div_t<true>(1.0, U[bj + j][o_a]);
        const double l = ssa_mul(U[bj + r][o_a], inv);
#pragma unroll
        for (int c = j + 1; c < 6; ++c)
          if (c >= r) U[br + c][o_a] = ssa_fma(-l, U[bj + c][o_a], U[br + c][o_a]);
      }
      __syncthreads();
    }
    if (need && r_a < 6) {
      const int j = r_a;
      const int bj = j * 6 - (j * (j - 1)) / 2 - j;
      const double d = U[bj + j][o_a];
      if (!(d > 0.0)) bad[o_a] = 1;
      const double inv = ssa_div_t<true>(1.0, d);
      const double sj = ssa_sqrt_t<true>(d);
      U[bj + j][o_a] = sj;
#pragma unroll
      for (int c = 1; c < 6; ++c)
        if (c > j) U[bj + c][o_a] = ssa_mul(ssa_mul(U[bj + c][o_a], inv), sj);
    }
    __syncthreads();
    int again = 0;
    if (need && r_a == NR - 1) {
      if (bad[o_a]) { again = 1; bad[o_a] = 0; }
      else res[o_a] = t + 1;
    }
    if (!__syncthreads_or(again)) break;
  }
}

// One launch per step: a CTA carries its T objects from their state in HBM to their next state.  Stage (xt, x, P) with
// one TMA tile load -> factor (lambda + n) P -> 14 T propagations into the shared sigma tile -> unscented transform + Q
// (mean and covariance stay in the tile) -> factor the predicted covariance (the re-drawn sigma points of filterpy's
// predict()), failures of the predict -> the update body of k_update_tile -> (xt', x, P, obs, errors).  Replaces the chain
// k_factor -> k_predict_tile -> k_refactor -> k_update_tile for the flag combination PREDICT | UPDATE_ALL: the factor U,
// the predicted mean / covariance and the propagated truth never travel through HBM (1.89 KB -> ~0.7 KB per object-step).
// The propagated sigma tile F[78][T + 1] lives in the storage of ZS | UVW (dead until the measurement phase).
template <int T, int NT>
__global__ void __launch_bounds__(NT, 4) k_step_tile(const KParams p, const __grid_constant__ CUtensorMap tm_s,
                                                                const __grid_constant__ CUtensorMap tm_u) {
  constexpr int NR = NT / T;
  static_assert(NR >= 7 && T == 32, "the algebra phases use 7 thread rows; the tile is staged by one TMA load");
  extern __shared__ __align__(128) unsigned char tile_raw[];
  UpdateTile<T>& sm = *reinterpret_cast<UpdateTile<T>*>(tile_raw);
  double(*sF)[T + 1] = reinterpret_cast<double(*)[T + 1]>(&sm.ZS[0][0]);  // [78][T + 1] = ZS | UVW
  static_assert(sizeof(sm.ZS) + sizeof(sm.UVW) == sizeof(double) * 78 * (T + 1) && offsetof(UpdateTile<T>, UVW) == offsetof(UpdateTile<T>, ZS) + sizeof(sm.ZS), "F aliases ZS | UVW");
  (void)tm_u;
  pdl_prologue();
  const int tid = threadIdx.x;
  const int o_a = tid % T, r_a = tid / T;
  const long loc0 = (long)blockIdx.x * T;
  const long ld = p.ld;
  const bool truth = (p.flags & SSA_STEP_TRUTH) != 0;
  const long loc_a = loc0 + o_a;
  const bool valid_a = loc_a < p.Nc;
  const long obj_a = p.obj0 + loc_a;

  // ---- stage: 33 rows (xt 6 | x 6 | P 21) by one TMA tile load, z_noise, status ----
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_expect_tx(&sm.bar, 33 * 256);
    tma_load_2d(&sm.S[0][0], &tm_s, (int)(p.obj0 + loc0), 0, &sm.bar);
  }
  if (tid < 3 * T) {  // z_noise is [N][3]: the tile's 3 T doubles are contiguous
    const long loc = loc0 + tid / 3;
    sm.zn[tid % 3][tid / 3] = (p.z_noise && loc < p.Nc) ? p.z_noise[(p.obj0 + loc0) * 3 + tid] : 0.0;
  } else if (tid < 4 * T) {
    const int o = tid - 3 * T;
    int st = SSA_ST_FAILED;
    if (loc0 + o < p.Nc) st = p.status[p.obj0 + loc0 + o];
    const int alive = (loc0 + o < p.Nc) && !(st & SSA_ST_FAILED);
    sm.st[o] = st; sm.code[o] = 0; sm.ok[o] = 1; sm.nan[o] = 0; sm.vis[o] = 0; sm.want[o] = 1;
    sm.live0[o] = alive; sm.live[o] = 0; sm.exc[o] = 0; sm.texc[o] = 0; sm.pnan[o] = 0; sm.bad[o] = 0;
    sm.res[o] = alive ? CHOL_PENDING : 0;
  }
  __syncthreads();
  mbar_wait(&sm.bar, 0);

  // ---- sigma points of the prior: U^T U = (lambda + n) P with the inflation fallback (k_factor) ----
  tile_robust_chol<T, NT>(sm.S + US_P, sm.S + US_U, p.lam, sm.res, sm.bad, r_a, o_a);
  int infl_a = 0, code_a = 0;  // thread row NR - 1 keeps the per-object words of its column
  if (r_a == NR - 1) {
    if (sm.live0[o_a]) {
      const int r1 = sm.res[o_a];
      if (r1 == CHOL_PENDING) code_a = SSA_ST_LINALG;
      else if (r1 > 0) infl_a = 1;
      sm.live[o_a] = (code_a == 0);
    }
  }
  __syncthreads();

  // ---- 14 T propagations: task = (object, sigma index), 13 = the TRUE state; ONE call site of fx ----
#pragma unroll 1
  for (int o = tid / 14; o < T; o += NT / 14) {  // NT is a multiple of 14: a thread keeps its sigma index in every round
    const int k = tid % 14;
    const bool is_truth = (k == 13);
    const bool run = (loc0 + o < p.Nc) && (is_truth ? truth : (sm.live[o] != 0));
    if (run) {
      double s[6], f[6];
      tile_sigma<T>(sm.S + US_X, sm.S + US_U, o, is_truth ? 0 : k, s);
      if (is_truth) {
#pragma unroll
        for (int j = 0; j < 6; ++j) s[j] = sm.S[US_XT + j][o];
      }
      const int exc = ssa_fx(s, p.dt, f);
      if (is_truth) {
#pragma unroll
        for (int j = 0; j < 6; ++j) sm.S[US_XT + j][o] = f[j];
        if (exc) sm.texc[o] = 1;
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) sF[k * 6 + j][o] = f[j];
        if (exc) sm.exc[o] = 1;
      }
    }
  }
  __syncthreads();

  // ---- unscented transform.  Thread (component i, object): mean over k = 0..12 in order, then the 13 deviations of that
  //      component in place; the mean replaces the prior mean in the tile (the sigma points are drawn).  The seventh
  //      thread row writes the propagated true state back ----
  const bool ut_a = valid_a && sm.live[o_a] && !sm.exc[o_a];
  if (r_a < 6) {
    if (ut_a) {
      double f[SSA_NSIG];
#pragma unroll
      for (int k = 0; k < SSA_NSIG; ++k) f[k] = sF[k * 6 + r_a][o_a];
      double acc = ssa_mul(p.Wm[0], f[0]);
#pragma unroll
      for (int k = 1; k < SSA_NSIG; ++k) acc = ssa_fma(p.Wm[k], f[k], acc);
      if (ssa_isnan(acc)) sm.pnan[o_a] = 1;
      sm.S[US_X + r_a][o_a] = acc;
#pragma unroll
      for (int k = 0; k < SSA_NSIG; ++k) sF[k * 6 + r_a][o_a] = f[k] - acc;
    }
  } else if (valid_a && truth) {
#pragma unroll
    for (int row = r_a - 6; row < 6; row += NR - 6) p.xt[row * ld + obj_a] = sm.S[US_XT + row][o_a];
  }
  __syncthreads();
  // covariance + Q: thread (row pair {i, 5 - i}, object), each element summed over k = 0..12 in order
  if (r_a < 3 && ut_a) {
    const int i = r_a, i2 = 5 - i;
    double acc[7];
    if (i == 0) tile_cov_rows<0, T>(sF, o_a, p.Wc, acc);
    else if (i == 1) tile_cov_rows<1, T>(sF, o_a, p.Wc, acc);
    else tile_cov_rows<2, T>(sF, o_a, p.Wc, acc);
    const int b1 = i * 6 - (i * (i - 1)) / 2, n1 = 6 - i;  // packed index of (i, j): i*6 - i(i-1)/2 + (j - i)
    const int b2 = i2 * 6 - (i2 * (i2 - 1)) / 2;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const int e = (q < n1) ? (b1 + q) : (b2 + q - n1);
      sm.S[US_P + e][o_a] = acc[q] + __ldg(p.qr + e);
    }
  } else if (r_a == NR - 1) {  // failure code of the propagation / transform; which objects re-draw their sigma points
    if (sm.live[o_a]) code_a = sm.exc[o_a] ? SSA_ST_FXEXC : (sm.pnan[o_a] ? SSA_ST_NAN : 0);
    sm.res[o_a] = (sm.live[o_a] && (code_a == 0 || code_a == SSA_ST_NAN)) ? CHOL_PENDING : 0;
  }
  __syncthreads();

  // ---- the re-drawn sigma points of filterpy's predict(): factor of the predicted covariance (k_refactor) ----
  tile_robust_chol<T, NT>(sm.S + US_P, sm.S + US_U, p.lam, sm.res, sm.bad, r_a, o_a);
  if (r_a == NR - 1 && valid_a) {  // per-object words of the predict; a failure stores the sentinels, SS2:369-382
    int st = sm.st[o_a];
    const int st0 = st;
    if (sm.texc[o_a]) st |= SSA_ST_TRUTHEXC;
    if (sm.live0[o_a]) {
      if (sm.live[o_a] && (code_a == 0 || code_a == SSA_ST_NAN)) {
        const int r2 = sm.res[o_a];
        if (r2 == CHOL_PENDING) code_a |= SSA_ST_LINALG;
        else if (r2 > 0) infl_a += 1;
      }
      if (code_a) {
        st |= SSA_ST_FAILED | code_a;
#pragma unroll
        for (int i = 0; i < 6; ++i) sm.S[US_X + i][o_a] = i < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL;
#pragma unroll
        for (int e = 0; e < SSA_NP; ++e) sm.S[US_P + e][o_a] = (c_pi[e] == c_pj[e]) ? (c_pi[e] < 3 ? SSA_XFAIL_POS : SSA_XFAIL_VEL) : 0.0;
      }
      if (infl_a) p.infl[obj_a] += infl_a;
    }
    p.code[obj_a] = code_a;
    if (st != st0) p.status[obj_a] = st;
    sm.st[o_a] = st;
  }
  __syncthreads();

  // ---- measurement update of every object + epilogue ----
  tile_update_body<T, NT, true>(p, sm, loc0);
}
