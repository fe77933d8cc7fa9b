"""Parity of the one-star kernel's variants that the goldens alone do not reach: the full-width column variant
(SRHMC_CHAIN_WINDOW=0, and a PSF wider than 1.5 px that selects it automatically), stars next to / on / outside the
image edges where the 24-column window is clamped (chain_kernel.cuh: `jb`), and a chain that reflects off the image
boundary (sampler_RHMC.py:554-564).  The checker is the NumPy oracle (oracle/stellar_oracle.py, full-image PSF as
utils.py:475-486) and, where they exist, the reference's recorded chains."""
import numpy as np
import pytest

import stellar_oracle as so
from helpers import first_divergence, golden, relerr, setup_from
from test_gpu_parity import RTOL, grad_relerr, make_ctx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("name", ["chain_one_star_m19", "chain_one_star_m21", "chain_one_star_m20_sep1"])
def test_reference_chain_through_both_column_variants(name, window, monkeypatch):
    """The recorded reference chains through the 24-column window (default) and the full-width variant."""
    monkeypatch.setenv("SRHMC_CHAIN_WINDOW", window)
    g = golden(name)
    S = setup_from(g)
    q0 = so.format_q(S, g["q_model"])
    with make_ctx(S, max_stars=1) as ctx:
        ctx.set_data(S.D)
        r = ctx.run(q0[None], int(g["niter"]), int(g["nsteps"]), float(g["dt"]), normals=g["normals"][None],
                    lnu=g["lnu"][None], g_ff2=S.g_ff2, beta=S.beta, f_pos=True)
    assert np.array_equal(r.A_chain[0].astype(bool), g["A_chain"]), "accept decisions differ"
    for arr, ref, tol, what in ((r.q_chain[0], g["q_chain"], 1e-9, "q"), (r.p_chain[0], g["p_chain"], 1e-9, "p"),
                                (r.E_chain[0], g["E_chain"], RTOL, "E"), (r.T_chain[0], g["T_chain"], 1e-9, "T")):
        i = first_divergence(arr, ref, tol)
        assert i == -1, "%s first differs from the reference at iteration %d (tolerance %g)" % (what, i, tol)


def _one_star_setup(fwhm, seed, q_true):
    S = so.Setup(num_rows=32, num_cols=32, dt=0.2, g_xx=1.0, g_ff=1.0, g_ff2=1.0, PSF_FWHM_pix=fwhm)
    rng = np.random.RandomState(seed)
    S.D = rng.poisson(so.model_image(S, q_true)).astype(float)
    return S


EDGE_STARS = [(0.3, 31.6), (31.2, 0.4), (15.5, 0.0), (-0.4, 16.2), (16.0, 32.7), (31.9, 31.9), (11.3, 13.0), (20.49, 3.51)]


@pytest.mark.parametrize("window", ["0", "1"])
def test_eval_with_stars_at_the_image_edges(window, monkeypatch):
    """V, dV/dq, H for stars near, on and just outside every edge: the column window is clamped there and the row window
    is one-sided.  One field per position, all in one launch (so the warp's row window is the union over very different
    chains)."""
    monkeypatch.setenv("SRHMC_CHAIN_WINDOW", window)
    f = so.Setup().mag2flux_converter(18.0)
    F = len(EDGE_STARS)
    S = so.Setup(num_rows=32, num_cols=32, dt=0.2, g_xx=1.0, g_ff=1.0, g_ff2=1.0)
    rng = np.random.RandomState(4)
    q = np.array([[f * (0.7 + 0.1 * i), x, y] for i, (x, y) in enumerate(EDGE_STARS)])
    D = np.stack([rng.poisson(so.model_image(S, qi)).astype(float) for qi in q])
    with make_ctx(S, n_fields=F, max_stars=1) as ctx:
        ctx.set_data(D)
        V, grad, H, Hg = ctx.eval(q, f_pos=True, g_ff2=S.g_ff2)
    for i in range(F):
        Si = S.clone(D=D[i])
        assert relerr(V[i], so.potential(Si, q[i], True)) < RTOL, EDGE_STARS[i]
        gi = so.grad_potential(Si, q[i])
        assert np.max(np.abs(grad[i] - gi) / np.maximum(np.abs(gi), 1e-6 * np.max(np.abs(gi)))) < RTOL, EDGE_STARS[i]
        assert relerr(H[i], so.metric(Si, q[i])) < RTOL


@pytest.mark.parametrize("window", ["0", "1"])
def test_chain_that_reflects_off_the_boundary(window, monkeypatch):
    """A faint star started half a pixel from the corner with a large step: the trajectory crosses x < 0 / y < 0, the
    position momenta flip (sampler_RHMC.py:554-564) and the column window slides along the edge.  Decisions and
    trajectory against the oracle with the same Philox draws."""
    monkeypatch.setenv("SRHMC_CHAIN_WINDOW", window)
    S0 = so.Setup(num_rows=32, num_cols=32, dt=0.35, g_xx=1.0, g_ff=1.0, g_ff2=1.0)
    q_true = np.array([S0.mag2flux_converter(21.0), 0.6, 0.7])
    S = _one_star_setup(S0.PSF_FWHM_pix, 12, q_true)
    S.dt = 0.35
    niter, nsteps = 40, 10
    with make_ctx(S, n_fields=1, max_stars=1) as ctx:
        ctx.set_data(S.D)
        normals, lnu = ctx.philox_draws(99, niter)
        r = ctx.run(q_true[None], niter, nsteps, S.dt, seed=99, g_ff2=S.g_ff2, f_pos=True)
    out = so.run_rhmc(S, q_true, normals[0], lnu[0], niter, nsteps, S.dt)
    assert (out.q[:, 1:] < 0.0).any() or (np.diff(np.sign(out.p[:, 1])) != 0).any()  # the boundary is actually reached
    assert np.array_equal(r.A_chain[0].astype(bool), out.A)
    assert first_divergence(r.q_chain[0], out.q, 1e-9) == -1
    assert first_divergence(r.E_chain[0], out.E, RTOL) == -1
    # the leapfrog reflects at least once inside these trajectories: replay one step at a time where x or y is negative
    assert out.A.sum() > 0


def test_wide_psf_selects_the_full_width_kernel_and_matches_the_oracle():
    """sigma = 1.8 px (FWHM 4.2372 px): every column carries weight above 2^-46, so the host must pick the full-width
    variant; eval and a Philox chain against the oracle."""
    fwhm = 1.8 * 2.354
    S0 = so.Setup(num_rows=32, num_cols=32, PSF_FWHM_pix=fwhm)
    q_true = np.array([S0.mag2flux_converter(19.0), 15.7, 16.4])
    S = _one_star_setup(fwhm, 3, q_true)
    q0 = q_true * np.array([1.02, 1.0, 1.0]) + np.array([0.0, 0.3, -0.2])
    niter, nsteps = 30, 10
    with make_ctx(S, n_fields=1, max_stars=1) as ctx:
        ctx.set_data(S.D)
        V, grad, H, _ = ctx.eval(q0[None], f_pos=True, g_ff2=S.g_ff2)
        normals, lnu = ctx.philox_draws(5, niter)
        r = ctx.run(q0[None], niter, nsteps, S.dt, seed=5, g_ff2=S.g_ff2, f_pos=True)
    assert relerr(V[0], so.potential(S, q0, True)) < RTOL
    assert grad_relerr(grad[0], so.grad_potential(S, q0)) < RTOL
    out = so.run_rhmc(S, q0, normals[0], lnu[0], niter, nsteps, S.dt)
    assert np.array_equal(r.A_chain[0].astype(bool), out.A) and out.A.sum() > 0
    assert first_divergence(r.q_chain[0], out.q, 1e-9) == -1
    assert first_divergence(r.p_chain[0], out.p, 1e-9) == -1
    assert relerr(r.E_chain[0], out.E) < RTOL


def test_faint_star_gradients_at_the_1e10_class_in_a_crowded_field():
    """The 204-star field of the reference golden with the scale floor of the gradient comparison lowered from 1e-3 to
    1e-6 of the per-coordinate maximum: faint stars' components are then checked to ~1e-10 of their own size unless they
    vanish by cancellation."""
    g = golden("field_eval_204")
    S = setup_from(g)
    with make_ctx(S, max_stars=204) as ctx:
        ctx.set_data(S.D)
        _, grad, _, _ = ctx.eval(g["q"], f_pos=True, g_ff2=S.g_ff2)
    a, b = grad[0].reshape(-1, 3), g["dVdq"].reshape(-1, 3)
    scale = np.maximum(np.abs(b), 1e-6 * np.max(np.abs(b), axis=0, keepdims=True))
    err = np.abs(a - b) / scale
    assert err.max() < 5e-10, "worst component %s: %.2e" % (np.unravel_index(err.argmax(), err.shape), err.max())
