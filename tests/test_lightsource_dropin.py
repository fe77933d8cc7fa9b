"""samplers.lightsource_gym on the device (csrc/ls_kernel.cuh through the C ABI): the reference's own script flows,
same seeds and calls as tests/golden/make_golden.py used on the unmodified reference, replayed through the drop-in
class and compared with the recorded reference outputs; plus kernel-level checks against the NumPy oracle."""
import contextlib
import io

import numpy as np
import pytest

import stellar_oracle as so
from helpers import golden, relerr

RTOL = 1e-10


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def test_lightsource_call_surface():
    import inspect

    from hmc_stellar_toy_model_b200 import samplers as m

    g = m.lightsource_gym()
    assert (g.num_rows, g.num_cols, g.mB) == (48, 48, 23) and g.B_count == 24.98145266935892
    sig = inspect.signature(m.lightsource_gym.HMC_random)
    assert list(sig.parameters) == ["self", "q_model_0", "Nchain", "Niter", "thin_rate", "Nwarmup", "steps_min",
                                    "steps_max", "f_lim", "f_lim_default"]
    sig = inspect.signature(m.lightsource_gym.HMC_find_best_dt)
    assert list(sig.parameters) == ["self", "q_model_0", "steps_min", "steps_max", "Niter_per_trial", "Ntrial",
                                    "dt_f_coeff", "dt_xy_coeff", "default", "A_target_f", "A_target_xy"]
    sig = inspect.signature(m.lightsource_gym.RHMC_random)
    assert sig.parameters["dt_RHMC_xy"].default == 1. and sig.parameters["dt_RHMC_f"].default == 0.1
    g.num_rows = g.num_cols = 32
    g.compute_factors()
    assert np.allclose([g.factor0, g.factor1, g.factor2],
                       [0.035997054345069765, 0.45235232653061236, 0.008141675878296744], rtol=1e-12)
    g.HMC_find_best_dt(np.array([[1000.0, 16.0, 16.0]]), default=True, dt_f_coeff=0.05, dt_xy_coeff=2.0)
    assert np.allclose(g.dt, [50.0, 0.002, 0.002]) and g.d == 3
    sig = inspect.signature(m.lightsource_gym.find_peaks)
    assert list(sig.parameters) == ["self", "linear_pix_density", "dr_tol", "dmag_tol", "mag_lim", "Nstep", "dt_f_coeff",
                                    "dt_xy_coeff", "no_perturb"]
    with pytest.raises(AssertionError):
        g.find_peaks()  # no image yet (samplers.py:152-154)


@pytest.mark.gpu
def test_gym_functions_match_reference_kat3():
    from hmc_stellar_toy_model_b200.samplers import lightsource_gym

    g = golden("kat3")
    gym = lightsource_gym()
    gym.num_rows = gym.num_cols = 32
    gym.D = g["D"]
    gym.f_lim = 0.0
    gym.Nobjs, gym.d = 1, 3
    gym.compute_factors()
    q, p = g["q"], g["p"]
    dqdt, dpdt, E = gym.RHMC_efficient_computation(q, p, debug=False)
    assert relerr(dqdt, g["dqdt"]) < RTOL and relerr(dpdt, g["dpdt"]) < 1e-9 and relerr(E, g["E_eff"]) < RTOL
    assert relerr(gym.RHMC_efficient_computation(q, p, debug=False, dVdqq_only=True), g["dVdqq"]) < RTOL
    M = gym.mass_matrix(q)
    assert relerr(M, g["mass"]) < 1e-14
    assert np.allclose(gym.dlnDetdq(q), g["dlnDetdq"], rtol=1e-13) and np.allclose(gym.dpMpdq(q, p), g["dpMpdq"], rtol=1e-13)
    assert relerr(gym.K(p, M), g["K"]) < RTOL and relerr(gym.E(q, p, M), g["E"]) < RTOL
    assert relerr(gym.V(q), g["V"]) < RTOL
    assert np.allclose(gym.dVdq(q), g["dVdq"], rtol=1e-9, atol=1e-10 * np.max(np.abs(g["dVdq"])))
    gym.f_lim = 1e9
    assert gym.RHMC_efficient_computation(q, p, debug=False) == (np.inf, np.inf, np.inf)
    assert np.isinf(gym.E(q, p))


@pytest.mark.gpu
def test_one_star_inference_script_hmc_and_diag():
    """one-star-inference-single.py flow (make_golden.light_chains): default step sizes, HMC_random, RHMC_random_diag."""
    from hmc_stellar_toy_model_b200.samplers import lightsource_gym, mag2flux

    g = golden("light_chains")
    niter = int(g["niter"])
    np.random.seed(11)
    gym = lightsource_gym()
    gym.num_rows = gym.num_cols = 32
    q0 = np.array([[mag2flux(19.0) * gym.flux_to_count, 16.0 + np.random.randn(), 16.0 + np.random.randn()]])
    gym.gen_mock_data(q_true=q0)
    assert np.array_equal(gym.D, g["D"])
    gym.HMC_find_best_dt(q0, default=True, dt_f_coeff=0.05, dt_xy_coeff=2.0)
    assert np.array_equal(gym.dt, g["dt_vec"])
    f_lim = mag2flux(22.0) * gym.flux_to_count
    with quiet():
        gym.HMC_random(np.copy(q0), Nchain=1, Niter=niter, steps_max=20, steps_min=5, f_lim=f_lim)
    assert gym.q_chain.shape == (1, niter + 1, 3) and gym.A_chain.shape == (1, niter, 1)
    assert np.array_equal(gym.A_chain[0, :, 0], g["hmc_A"])
    assert relerr(gym.q_chain[0], g["hmc_q"]) < 1e-9
    assert relerr(gym.E_chain[0, :, 0], g["hmc_E"]) < RTOL
    assert np.allclose(gym.dE_chain[0, :, 0], g["hmc_dE"], rtol=0, atol=1e-7)
    gym.compute_factors()
    with quiet():
        gym.RHMC_random_diag(np.copy(q0), Nchain=1, Niter=niter, steps_max=20, steps_min=5, f_lim=f_lim, dt_global=5e-2)
    assert np.array_equal(gym.A_chain[0, :, 0], g["diag_A"])
    assert relerr(gym.q_chain[0], g["diag_q"]) < 1e-9
    assert relerr(gym.E_chain[0, :, 0], g["diag_E"]) < RTOL


@pytest.mark.gpu
def test_hessian_metric_chain_script():
    """make_golden.light_hess: RHMC_efficient_computation and a 40-iteration RHMC_random chain, state advanced in
    place, global stream left where the reference leaves it."""
    from hmc_stellar_toy_model_b200.samplers import lightsource_gym, mag2flux

    g = golden("light_hess")
    niter = int(g["niter"])
    np.random.seed(int(g["seed"]))
    gym = lightsource_gym()
    gym.num_rows = gym.num_cols = 32
    q0 = np.array([[mag2flux(19.0) * gym.flux_to_count, 16.0 + 0.3 * np.random.randn(), 16.0 + 0.3 * np.random.randn()]])
    gym.gen_mock_data(q_true=q0)
    assert np.array_equal(gym.D, g["D"])
    f_lim = mag2flux(22.0) * gym.flux_to_count
    gym.f_lim = f_lim
    gym.Nobjs, gym.d = 1, 3
    dqdt, dpdt, E = gym.RHMC_efficient_computation(g["qe"], g["pe"], debug=False)
    assert relerr(dqdt, g["dqdt"]) < RTOL and relerr(dpdt, g["dpdt"]) < 1e-9 and relerr(E, g["E_eff"]) < RTOL
    assert relerr(gym.RHMC_efficient_computation(g["qe"], g["pe"], debug=False, dVdqq_only=True), g["dVdqq"]) < RTOL
    q_start = np.copy(q0)
    with quiet():
        gym.RHMC_random(q_start, Nchain=1, Niter=niter, steps_max=12, steps_min=4, f_lim=f_lim, dt_RHMC_xy=0.3,
                        dt_RHMC_f=0.3, debug=False)
    assert np.array_equal(gym.A_chain[0, :, 0], g["A_chain"])
    assert relerr(gym.q_chain[0], g["q_chain"]) < 1e-9
    assert relerr(gym.E_chain[0, :, 0], g["E_chain"]) < RTOL
    assert relerr(q_start.reshape(-1), g["q_after"]) < 1e-9       # caller's array mutated like upstream
    assert np.random.random(1)[0] == g["next_uniform"][0]


@pytest.mark.gpu
def test_best_dt_search_script():
    """make_golden.best_dt: V_single / dVdq_single, the full HMC_find_best_dt search (device trials, host bisection)
    and the HMC_random chain that uses its step sizes."""
    from hmc_stellar_toy_model_b200.samplers import lightsource_gym, mag2flux

    g = golden("best_dt")
    np.random.seed(int(g["seed"]))
    gym = lightsource_gym()
    gym.num_rows = gym.num_cols = 32
    q0 = np.array([[mag2flux(19.0) * gym.flux_to_count, 16.0 + np.random.randn(), 16.0 + np.random.randn()]])
    gym.gen_mock_data(q_true=q0)
    assert np.array_equal(gym.D, g["D"])
    qs = q0.reshape(-1)
    assert relerr(gym.V_single(qs, g["model_data"]), g["V_single"]) < RTOL
    assert np.allclose(gym.dVdq_single(qs, g["model_data"], return_all=True), g["dVdq_single"], rtol=1e-9,
                       atol=1e-10 * np.max(np.abs(g["dVdq_single"])))
    assert relerr(gym.dVdq_single(qs, g["model_data"]), g["dVdq_single"][0]) < 1e-9
    with quiet():
        gym.HMC_find_best_dt(q0, default=False, dt_f_coeff=1, dt_xy_coeff=10, Niter_per_trial=20, Ntrial=6,
                             steps_min=5, steps_max=12, A_target_f=0.99, A_target_xy=0.5)
    assert relerr(gym.dt, g["dt_vec"]) < 1e-12
    with quiet():
        gym.HMC_random(np.copy(q0), Nchain=1, Niter=25, steps_max=20, steps_min=5, f_lim=float(g["f_lim"]))
    assert np.array_equal(gym.A_chain[0, :, 0], g["A_chain"])
    assert relerr(gym.q_chain[0], g["q_chain"]) < 1e-9
    assert np.random.random(1)[0] == g["next_uniform"][0]


@pytest.mark.gpu
def test_hessian_reductions_crowded_field_vs_oracle():
    """K7 on the 204-star 64x64 field: 17 separable sums per star against the oracle's full-image formulas."""
    from hmc_stellar_toy_model_b200 import RHMCContext

    g = golden("field_eval_204")
    L = so.LightSetup(num_rows=64, num_cols=64, D=g["D"], f_lim=0.0)
    q = g["q"]
    rng = np.random.RandomState(2)
    p = rng.randn(q.size) * np.sqrt(np.abs(so.ls_efficient(L, q, np.zeros_like(q), dVdqq_only=True)))
    dqdt, dpdt, E, d1, d2, d3 = so.ls_efficient(L, q, p, parts=True)
    with RHMCContext(n_fields=2, num_rows=64, num_cols=64, max_stars=204, psf_fwhm_pix=L.PSF_FWHM_pix,
                     B_count=L.B_count, f_lim=0.0, f_low=1.0, g0=1.0, g1=1.0, g2=1.0, g_xx=1.0, g_ff=1.0,
                     enable_hessian=True) as ctx:
        ctx.set_data(np.stack([g["D"], g["D"]]))
        a1, a2, a3, adq, adp, aE = ctx.hessian(np.stack([q, q]), np.stack([p, p]))
        only = ctx.hessian(np.stack([q, q]), d2_only=True)

    def close(a, b, tol):
        a, b = a.reshape(-1, 3), b.reshape(-1, 3)
        scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b), axis=0, keepdims=True))
        return np.max(np.abs(a - b) / scale) < tol

    assert close(a1[0], d1, 1e-9) and close(a2[0], d2, 1e-9) and close(a3[0], d3, 1e-8)
    assert np.array_equal(a2[0], only[1]) and np.array_equal(a1[0], a1[1])
    assert close(adq[0], dqdt, 1e-9) and close(adp[0], dpdt, 1e-8)
    assert (np.isnan(E) and np.isnan(aE[0])) or relerr(aE[0], E) < RTOL


@pytest.mark.gpu
def test_step_size_trial_kernel_vs_oracle():
    from hmc_stellar_toy_model_b200 import RHMCContext

    g = golden("best_dt")
    L = so.LightSetup(num_rows=32, num_cols=32, D=g["D"])
    rng = np.random.RandomState(4)
    n = 30
    normals = rng.randn(n + 1, 3)
    steps = rng.randint(5, 12, n)
    lnu = np.log(rng.random_sample(n))
    q = g["q0"]
    psf = so.gauss_psf(32, 32, q[1], q[2], L.PSF_FWHM_pix)
    model_data = g["D"] - q[0] * psf * 0.9 + 3.0   # a data-like background with the star mostly removed
    with RHMCContext(n_fields=1, num_rows=32, num_cols=32, max_stars=1, psf_fwhm_pix=L.PSF_FWHM_pix,
                     B_count=L.B_count, f_lim=0.0, f_low=1.0, g0=1.0, g1=1.0, g2=1.0, g_xx=1.0, g_ff=1.0) as ctx:
        ctx.set_data(g["D"])
        for zero_xy, dt in ((True, np.array([q[0] * 0.06, 0.0, 0.0])), (False, np.array([q[0] * 0.05, 2e-3, 2e-3]))):
            out = ctx.ls_run(ctx.LS_TRIAL, q[None], dt, normals[None], steps[None], lnu[None],
                             background=model_data[None], zero_xy=zero_xy)
            ref = so.ls_trial(L, q, model_data, dt, normals, steps, lnu, zero_xy)
            assert int(out["accept_count"][0]) == ref
            assert 0 < ref < n
