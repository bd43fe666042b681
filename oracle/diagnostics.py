"""TEST INFRASTRUCTURE — numpy restatement of the reference's consistency diagnostics, the checker of
`ssa_ukf_diagnostics`.  Never imported by the product.

  nees(x_true, x_filter, P)       SS2:436-446  `anees`: delta @ np.linalg.inv(P) @ delta  (mean over all = ANEES)
  nis(y, S)                       SS2:564-569  `plot_NIS`: y @ np.linalg.inv(S) @ y
  innovation_bounds(y, S)         SS2:598-604  fraction of innovations inside 1 / 2 standard deviations
  durbin_watson(e)                SS2:782-832  statsmodels.stats.stattools.durbin_watson: sum(diff(e)^2) / sum(e^2), axis 0
  acf_conservative(x, nlags)      SS2:655-668  statsmodels.tsa.stattools.acf(x, missing='conservative', fft=False)
                                               (statsmodels is not in the image: its published acovf algorithm restated)
"""
import numpy as np


def nees(x_true, x_filter, P):
    d = np.asarray(x_true) - np.asarray(x_filter)
    return np.array([d[j] @ np.linalg.inv(P[j]) @ d[j] for j in range(len(d))])


def nis(y, S):
    return np.array([y[j] @ np.linalg.inv(S[j]) @ y[j] for j in range(len(y))])


def innovation_flags(y, S):
    """per innovation: inside one sigma [.,3] bool, inside two sigmas [.,3] bool (strict, like SS2:601-602)."""
    sd = np.sqrt(np.array([np.diag(s) for s in S]))
    y = np.asarray(y)
    return (y < sd) * (y > -sd), (y < 2 * sd) * (y > -2 * sd)


def innovation_bounds(y, S):
    """SS2:598-604: percentages, rounded to 2 decimals, rows = (sigma, two sigmas), columns = measurement components."""
    one, two = innovation_flags(y, S)
    return np.round(np.stack((np.mean(one, axis=0), np.mean(two, axis=0))) * 100, 2)


def durbin_watson(e):
    """statsmodels.stats.stattools.durbin_watson along axis 0 (the reference calls the same formula inline, SS2:817-818)."""
    e = np.asarray(e, dtype=float)
    return np.sum(np.diff(e, 1, axis=0) ** 2, axis=0) / np.sum(e ** 2, axis=0)


def acf_conservative(x, nlags=40):
    """statsmodels acf(x, missing='conservative', fft=False, adjusted=False): acovf demeans by the mean of the non-NaN
    entries, sets the NaN entries to zero, correlates the series with itself (np.correlate 'full'), divides by the number
    of valid entries; acf = acov[:nlags + 1] / acov[0]."""
    x = np.array(x, dtype=float)
    ok = ~np.isnan(x)
    xo = x.copy()
    xo[~ok] = 0.0
    xo = xo - xo.sum() / ok.sum()
    xo[~ok] = 0.0
    n = len(x)
    acov = np.correlate(xo, xo, "full")[n - 1:] / float(ok.sum())
    return acov[:nlags + 1] / acov[0]
