"""Host-side helpers with the reference's names and argument meaning (reference utils.py).

Only what the RHMC path's callers need is mirrored here: unit conversions, the Gaussian PSF used to build mock
data and the frozen Fisher constants, Poisson realisation, power-law sampling and the exponential scheduler.
These run once at set-up time on the host, exactly as upstream; everything evaluated inside the sampler loops
runs in the CUDA kernels.  The chain statistics (convergence_stats: SURVEY.md 8f row 4) run on the device as well;
plotting and NUTS bookkeeping of utils.py are out of scope.
"""
from __future__ import annotations

import numpy as np

__all__ = ["mag2flux", "flux2mag", "gauss_PSF", "factors", "poisson_realization", "gen_pow_law_sample",
           "integrate_pow_law", "scheduler", "convergence_stats", "acceptance_rate", "np"]


def mag2flux(mag):
    """Magnitude -> flux in nanomaggies (utils.py:24)."""
    return 10 ** (0.4 * (22.5 - mag))


def flux2mag(flux):
    """utils.py:27."""
    return 22.5 - 2.5 * np.log10(flux)


def gauss_PSF(num_rows, num_cols, x, y, FWHM):
    """Normalised circular Gaussian over the whole image, pixel centres at i+0.5, x along rows (utils.py:475-486)."""
    sigma = FWHM / 2.354
    rows = np.arange(0.5, num_rows)[:, None]
    cols = np.arange(0.5, num_cols)[None, :]
    return np.exp(-(np.square(rows - x) + np.square(cols - y)) / (2 * sigma**2)) / (np.pi * 2 * sigma**2)


def factors(num_rows, num_cols, x, y, PSF_FWHM_pix):
    """g0 = sum PSF^2, g1 = sum PSF (x-l-.5)^2/s^4, g2 = sum PSF^2 (x-l-.5)^2/s^4 (utils.py:623-644)."""
    rowidx = np.arange(0, num_rows)[:, None] * np.ones((1, num_cols))
    var = (PSF_FWHM_pix / 2.354) ** 2
    psf = gauss_PSF(num_rows, num_cols, x, y, FWHM=PSF_FWHM_pix)
    psf_sq = np.square(psf)
    off_sq = (x - rowidx - 0.5) ** 2
    return np.sum(psf_sq), np.sum(psf * off_sq) / float(var**2), np.sum(psf_sq * off_sq) / float(var**2)


def poisson_realization(D0):
    """Poisson draw per pixel from the legacy global np.random stream, row-major like the reference's double loop
    (utils.py:488-496) so that a seeded script produces the same image."""
    D0 = np.asarray(D0, dtype=float)
    return np.random.poisson(lam=D0).astype(float).reshape(D0.shape)


def integrate_pow_law(alpha, A, fmin, fmax):
    """utils.py:452-457."""
    return A * (fmax ** (1 + alpha) - fmin ** (1 + alpha)) / (1 + alpha)


def gen_pow_law_sample(alpha, fmin, fmax, Nsample=1):
    """Draws from f^-alpha on [fmin, fmax] by inverse CDF (utils.py:460-471)."""
    assert alpha > 1
    alpha = float(alpha)
    u = np.random.random(size=Nsample)
    lmbda = fmin ** (1 - alpha) + u * (fmax ** (1 - alpha) - fmin ** (1 - alpha))
    return np.exp(np.log(lmbda) / (1 - alpha))


def scheduler(val_init, val_final, Niter=10):
    """Exponential schedule val_init -> val_final over Niter values (utils.py:649-660)."""
    c = np.exp(np.log(val_final / float(val_init)) / float(Niter - 1))
    return c ** np.arange(0, Niter, 1) * val_init


def convergence_stats(q_chain, thin_rate=5, warm_up_num=0, device=0):
    """Gelman-Rubin R and effective sample size per variable of q_chain [Nchain, Niter, D] (utils.py:86-167), computed
    by the CUDA library (srhmc_convergence_stats); the reference's Python-2 `n = L_chain/2` is an integer division.
    For chains that are still on the device use RHMCContext.run_stats instead."""
    from . import _capi

    q = np.ascontiguousarray(q_chain, dtype=np.float64)
    Nchain, Niter, D = q.shape
    assert Nchain > 1  # utils.py:94
    lib = _capi.load_library()
    R, n_eff = np.empty(D), np.empty(D)
    _capi.check(lib.srhmc_convergence_stats(int(device), _capi.dptr(q), Nchain, Niter, D, 1, int(thin_rate), int(warm_up_num),
                                            _capi.dptr(R), _capi.dptr(n_eff)))
    return R, n_eff


def acceptance_rate(decision_chain, start=None, end=None):
    """Fraction of accepted proposals per chain; decision_chain [Nchain, Niter, 1] (utils.py:192-209)."""
    decision_chain = np.asarray(decision_chain)
    _, Niter, _ = decision_chain.shape
    if start is None and end is None:
        return np.sum(decision_chain, axis=(1, 2)) / Niter
    Niter = (end - start) if end > 0 else (Niter - start)
    return np.sum(decision_chain[:, start:end, :], axis=(1, 2)) / Niter
