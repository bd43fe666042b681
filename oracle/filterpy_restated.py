"""numpy restatement of the filterpy pieces the reference uses (ORACLE — test infrastructure).

filterpy is a third-party dependency of the reference (requirements.txt:14, unpinned; release 1.4.5
is the one contemporary with the repo) and is neither vendored under /root/reference nor installed in
this image.  Its published algorithm is restated here with the same numpy/scipy calls filterpy makes
(np.dot, np.outer, np.linalg.inv, user-supplied sqrt/mean/residual callables), so that the reference's
own callables (envs/farnocchia.py fx, envs/transformations.py geometry) can be plugged in unchanged
and golden vectors generated (tests/golden/make_golden.py).

Reference call sites (file:line in upstream): envs/ssa_tasker_simple_2.py:4-6 (imports), :110
(Q_discrete_white_noise), :211-218 (construction), :275 (predict), :301-306 (update, .y .S .sigmas_h);
tests.py:119-151 (Test 6), :161-186 (Test 7).

One behavioural fork cannot be verified offline: whether predict() re-draws `sigmas_f`
from the prior after the unscented transform (filterpy >= 1.4.5 / master) or keeps the propagated
points (book version).  `resample_after_predict` selects; default True.
"""
import numpy as np


def Q_discrete_white_noise(dim, dt=1.0, var=1.0, block_size=1, order_by_dim=True):
    if dim != 2:
        raise NotImplementedError("the reference only uses dim=2 (SS2:110)")
    Q = [[.25 * dt ** 4, .5 * dt ** 3],
         [.5 * dt ** 3, dt ** 2]]
    if order_by_dim:
        raise NotImplementedError("the reference passes order_by_dim=False")
    # order_by_block: each scalar of Q expands to eye(block_size) * value
    Q = np.array(Q, dtype=float)
    out = np.zeros((dim * block_size, dim * block_size))
    for i in range(dim):
        for j in range(dim):
            out[i * block_size:(i + 1) * block_size, j * block_size:(j + 1) * block_size] = \
                np.eye(block_size) * Q[i, j]
    return out * var


class MerweScaledSigmaPoints:
    def __init__(self, n, alpha, beta, kappa, sqrt_method=None, subtract=None):
        self.n = n
        self.alpha = alpha
        self.beta = beta
        self.kappa = kappa
        if sqrt_method is None:
            import scipy.linalg
            sqrt_method = scipy.linalg.cholesky
        self.sqrt = sqrt_method
        self.subtract = np.subtract if subtract is None else subtract
        self._compute_weights()

    def num_sigmas(self):
        return 2 * self.n + 1

    def _compute_weights(self):
        n = self.n
        lambda_ = self.alpha ** 2 * (n + self.kappa) - n
        c = .5 / (n + lambda_)
        self.Wc = np.full(2 * n + 1, c)
        self.Wm = np.full(2 * n + 1, c)
        self.Wc[0] = lambda_ / (n + lambda_) + (1 - self.alpha ** 2 + self.beta)
        self.Wm[0] = lambda_ / (n + lambda_)

    def sigma_points(self, x, P):
        n = self.n
        x = np.asarray(x, dtype=float)
        P = np.atleast_2d(P)
        lambda_ = self.alpha ** 2 * (n + self.kappa) - n
        U = self.sqrt((lambda_ + n) * P)
        sigmas = np.zeros((2 * n + 1, n))
        sigmas[0] = x
        for k in range(n):
            sigmas[k + 1] = self.subtract(x, -U[k])
            sigmas[n + k + 1] = self.subtract(x, U[k])
        return sigmas


def unscented_transform(sigmas, Wm, Wc, noise_cov=None, mean_fn=None, residual_fn=None):
    kmax, n = sigmas.shape
    if mean_fn is None:
        x = np.dot(Wm, sigmas)
    else:
        x = mean_fn(sigmas, Wm)
    if residual_fn is np.subtract or residual_fn is None:
        y = sigmas - x[np.newaxis, :]
        P = np.dot(y.T, np.dot(np.diag(Wc), y))
    else:
        P = np.zeros((n, n))
        for k in range(kmax):
            y = residual_fn(sigmas[k], x)
            P += Wc[k] * np.outer(y, y)
    if noise_cov is not None:
        P += noise_cov
    return x, P


class UnscentedKalmanFilter:
    def __init__(self, dim_x, dim_z, dt, hx, fx, points, sqrt_fn=None, x_mean_fn=None, z_mean_fn=None,
                 residual_x=None, residual_z=None, resample_after_predict=True):
        self.x = np.zeros(dim_x)
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.R = np.eye(dim_z)
        self._dim_x = dim_x
        self._dim_z = dim_z
        self.points_fn = points
        self._dt = dt
        self._num_sigmas = points.num_sigmas()
        self.hx = hx
        self.fx = fx
        self.x_mean = x_mean_fn
        self.z_mean = z_mean_fn
        self.Wm, self.Wc = points.Wm, points.Wc
        self.residual_x = np.subtract if residual_x is None else residual_x
        self.residual_z = np.subtract if residual_z is None else residual_z
        self.sigmas_f = np.zeros((self._num_sigmas, dim_x))
        self.sigmas_h = np.zeros((self._num_sigmas, dim_z))
        self.K = np.zeros((dim_x, dim_z))
        self.y = np.zeros(dim_z)
        self.S = np.zeros((dim_z, dim_z))
        self.SI = np.zeros((dim_z, dim_z))
        self.inv = np.linalg.inv
        self.resample_after_predict = resample_after_predict

    def predict(self, dt=None, **fx_args):
        if dt is None:
            dt = self._dt
        sigmas = self.points_fn.sigma_points(self.x, self.P)
        for i, s in enumerate(sigmas):
            self.sigmas_f[i] = self.fx(s, dt, **fx_args)
        self.x, self.P = unscented_transform(self.sigmas_f, self.Wm, self.Wc, self.Q, self.x_mean, self.residual_x)
        if self.resample_after_predict:
            # "update sigma points to reflect the new variance of the points" (filterpy >= 1.4.5)
            self.sigmas_f = self.points_fn.sigma_points(self.x, self.P)
        self.x_prior = np.copy(self.x)
        self.P_prior = np.copy(self.P)

    def update(self, z, R=None, **hx_args):
        if R is None:
            R = self.R
        elif np.isscalar(R):
            R = np.eye(self._dim_z) * R
        sigmas_h = []
        for s in self.sigmas_f:
            sigmas_h.append(self.hx(s, **hx_args))
        self.sigmas_h = np.atleast_2d(sigmas_h)
        zp, self.S = unscented_transform(self.sigmas_h, self.Wm, self.Wc, R, self.z_mean, self.residual_z)
        self.SI = self.inv(self.S)
        Pxz = self.cross_variance(self.x, zp, self.sigmas_f, self.sigmas_h)
        self.K = np.dot(Pxz, self.SI)
        self.y = self.residual_z(z, zp)
        self.x = self.x + np.dot(self.K, self.y)
        self.P = self.P - np.dot(self.K, np.dot(self.S, self.K.T))
        self.z = np.copy(z)
        self.x_post = self.x.copy()
        self.P_post = self.P.copy()

    def cross_variance(self, x, z, sigmas_f, sigmas_h):
        Pxz = np.zeros((sigmas_f.shape[1], sigmas_h.shape[1]))
        N = sigmas_f.shape[0]
        for i in range(N):
            dx = self.residual_x(sigmas_f[i], x)
            dz = self.residual_z(sigmas_h[i], z)
            Pxz += self.Wc[i] * np.outer(dx, dz)
        return Pxz
