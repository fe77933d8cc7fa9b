"""The drop-in gym classes (hmc_stellar_toy_model_b200.sampler_RHMC) driven exactly as the reference's scripts drive
the reference's classes -- same constructor arguments, attribute pokes, np.random seeds and method calls as
tests/golden/make_golden.py used on the unmodified reference -- must reproduce the recorded reference chains.
Nothing is injected: data generation, draw order and chain layout all come from the drop-in itself."""
import contextlib
import io

import numpy as np
import pytest

from helpers import first_divergence, golden, relerr

RTOL = 1e-10


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------ CPU: call surface only
def test_call_surface_matches_reference_names():
    import inspect

    from hmc_stellar_toy_model_b200 import sampler_RHMC as m

    sig = inspect.signature(m.multi_gym.run_RHMC)
    assert list(sig.parameters) == ["self", "q_model_0", "f_pos", "delta", "Niter", "Nsteps", "dt", "save_traj",
                                    "counter_max", "verbose", "q_true", "schedule_g_ff2", "N_max", "P_move",
                                    "schedule_beta"]
    assert sig.parameters["N_max"].default == 50 and sig.parameters["counter_max"].default == 1000
    sig = inspect.signature(m.single_gym.run_single_RHMC)
    assert list(sig.parameters) == ["self", "q_model_0", "f_pos", "solver", "delta", "p_initial", "counter_max"]
    assert sig.parameters["counter_max"].default == 100
    for name in ("gen_mock_data", "gen_model", "gen_noise_profile", "mag2flux_converter", "flux2mag_converter",
                 "compute_factors", "default_exp_setup", "u_sample", "format_q", "reverse_format_q", "H", "H_xx",
                 "H_ff", "V", "T", "dVdq", "dphidq", "dtaudq", "dtaudp", "RHMC_single_step", "display_image"):
        assert callable(getattr(m.base_class, name)), name
    g = m.multi_gym(dt=0., Nsteps=0, g_xx=0.05, g_ff=4., g_ff2=4.)
    assert (g.num_rows, g.num_cols, g.mB) == (48, 48, 23)
    assert g.PSF_FWHM_pix == 1.4 / 0.4 and g.f_lim == g.B_count == 24.98145266935892
    # SURVEY.md KAT-3 factors, frozen at 48x48
    assert np.allclose([g.g0, g.g1, g.g2], [0.035997054345069765, 0.45235232653061236, 0.008141675878296744],
                       rtol=1e-14)
    s = m.single_gym(g_ff2=7.)
    assert s.g_ff2 == 1.  # the reference drops the argument (sampler_RHMC.py:578)
    q = np.array([[19.0, 3.0, 4.0]])
    flat = g.format_q(q)
    assert flat.shape == (3,) and q[0, 0] == flat[0] == g.mag2flux_converter(19.0)  # in-place conversion


def test_mock_data_follows_reference_stream():
    """gen_mock_data consumes the legacy np.random stream like the reference's per-pixel loop (utils.py:488-496):
    the golden D was produced by the reference from the same seed."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym

    g = golden("chain_one_star_m19")
    np.random.seed(1903)
    gym = multi_gym(dt=0.0, Nsteps=0, g_xx=1.0, g_ff=1.0)
    gym.num_rows = gym.num_cols = 32
    gym.gen_mock_data(np.array([[19.0, 16.0, 16.0]]))
    assert np.array_equal(gym.D, g["D"])


def test_unsupported_solvers_are_refused_loudly():
    s = __import__("hmc_stellar_toy_model_b200.sampler_RHMC", fromlist=["single_gym"]).single_gym()
    with pytest.raises(NotImplementedError):
        s.run_single_RHMC(np.array([[19.0, 8.0, 8.0]]), solver="naive")


# ------------------------------------------------------------------ GPU: scripts replayed through the drop-in
@pytest.mark.gpu
@pytest.mark.parametrize("name,mT,seed,kw", [
    ("chain_one_star_m19", 19.0, 1903, dict(sep=0.0)),
    ("chain_one_star_m21", 21.0, 2103, dict(sep=0.0)),
    ("chain_one_star_m20_sep1", 20.0, 2003, dict(mM=19.0, sep=1.0, dt=0.05, niter=40)),
    ("chain_one_star_m15", 15.0, 1503, dict(sep=0.0, niter=30)),
])
def test_one_star_inference_script(name, mT, seed, kw):
    """RHMC-single-full-inference-test.py flow (make_golden.chain_one_star)."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym

    g = golden(name)
    niter, nsteps, dt, size = kw.get("niter", 60), 10, kw.get("dt", 0.2), 32
    sep, mM = kw.get("sep", 1.0), kw.get("mM", None)
    np.random.seed(seed)
    gym = multi_gym(dt=0.0, Nsteps=0, g_xx=1.0, g_ff=1.0)
    gym.num_rows = gym.num_cols = size
    gym.V_prior_const = 0.0
    q_true = np.array([[mT, size / 2.0, size / 2.0]])
    gym.gen_mock_data(q_true)
    q_model = np.array([[mM if mM is not None else mT, size / 2.0 + sep, size / 2.0]])
    with quiet():
        gym.run_RHMC(np.copy(q_model), f_pos=True, delta=1e-6, Niter=niter, Nsteps=nsteps, dt=dt, save_traj=False,
                     N_max=1)
    assert np.array_equal(gym.D, g["D"])
    assert gym.q_chain.shape == g["q_chain"].shape and gym.A_chain.dtype == bool
    assert np.array_equal(gym.A_chain, g["A_chain"])
    assert first_divergence(gym.q_chain, g["q_chain"], 1e-9) == -1
    assert first_divergence(gym.p_chain, g["p_chain"], 1e-8) == -1
    assert relerr(gym.E_chain, g["E_chain"]) < RTOL and relerr(gym.V_chain, g["V_chain"]) < RTOL
    assert np.all(gym.move_chain == 0) and np.all(gym.N_chain == 1)


@pytest.mark.gpu
def test_big_sim_script_with_schedules():
    """RHMC-big-sim2.py flow (make_golden.chain_multi): 30 stars, prior, repulsion, g_ff2 and beta schedules."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym, scheduler, gen_pow_law_sample

    g = golden("chain_multi30_vc")
    nobj, niter, nsteps, dt, size = 30, 12, 30, 1e-2, 32
    gff2_list = scheduler(1 / 10.0, 4.0, 500)
    beta_list = scheduler(1e-2, 1e-12, 500)
    gym = multi_gym(dt=0.0, Nsteps=0, g_xx=0.005, g_ff=25.0, g_ff2=2.0)
    np.random.seed(77)
    gym.num_rows = gym.num_cols = size
    q_true = np.zeros((nobj, 3))
    q_model = np.zeros((nobj, 3))
    alpha = 1.5
    fmin = gym.mag2flux_converter(20.5)
    fmax = gym.mag2flux_converter(15.0)
    mag = gym.flux2mag_converter(gen_pow_law_sample(alpha, fmin, fmax, nobj))
    for i in range(nobj):
        x = np.random.random() * (size - 2.0) + 1.0
        y = np.random.random() * (size - 2.0) + 1.0
        q_true[i] = np.array([mag[i], x, y])
    gym.use_prior, gym.alpha = True, alpha
    gym.fmin, gym.fmax = fmin, fmax
    gym.use_Vc, gym.beta, gym.f_expnt, gym.Vc_r_pow = True, 1e-4, np.zeros(nobj), 4.0
    fmin_m = gym.mag2flux_converter(22.9)
    fmax_m = gym.mag2flux_converter(21.0)
    q_model[:, 0] = gym.flux2mag_converter(gen_pow_law_sample(alpha, fmin_m, fmax_m, nobj))
    q_model[:, 1] = np.random.random(size=nobj) * (size - 2.0) + 1.0
    q_model[:, 2] = np.random.random(size=nobj) * (size - 2.0) + 1.0
    gym.gen_mock_data(q_true)
    with quiet():
        gym.run_RHMC(np.copy(q_model), f_pos=True, delta=1e-6, Niter=niter, Nsteps=nsteps, dt=dt, save_traj=False,
                     schedule_g_ff2=gff2_list, schedule_beta=beta_list, N_max=nobj)
    assert np.array_equal(gym.D, g["D"])
    assert np.allclose(q_model, g["q_model"], rtol=0, atol=0)
    assert np.array_equal(gym.A_chain, g["A_chain"])
    assert relerr(gym.E_chain, g["E_chain"]) < 1e-9
    first_accept = int(np.argmax(g["A_chain"])) + 1
    assert first_divergence(gym.q_chain[: first_accept + 1], g["q_chain"][: first_accept + 1], RTOL) == -1
    # the schedules leave their last applied value on the gym, like the reference
    assert gym.g_ff2 == gff2_list[niter] and gym.beta == beta_list[niter]
    assert abs(gym.V_prior_const - float(g["V_prior_const"])) < 1e-12


@pytest.mark.gpu
def test_single_trajectory_script():
    """RHMC-single-tests.py flow (make_golden.single_traj): 16x16, implicit solver, energy drift chains."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import single_gym

    g = golden("single_traj")
    np.random.seed(5)
    gym = single_gym(dt=0.0, Nsteps=0, g_xx=1.0, g_ff=1.0)
    gym.num_rows = gym.num_cols = 16
    gym.V_prior_const = 0.0
    gym.gen_mock_data(np.array([[19.0, 8.0, 8.0]]))
    q_model = np.array([[20.0, 9.0, 8.0]])
    gym.Nsteps, gym.dt = 200, 0.1
    gym.run_single_RHMC(q_model_0=np.copy(q_model), f_pos=True, solver="implicit", delta=1e-6, p_initial=None)
    assert np.array_equal(gym.D, g["D"])
    assert relerr(gym.p_chain[0], g["p0"]) < 1e-13       # momentum draw scaled by the device metric
    assert first_divergence(gym.q_chain, g["q_chain"], 1e-9) == -1
    assert first_divergence(gym.p_chain, g["p_chain"], 1e-8) == -1
    scale = np.max(np.abs(g["V_chain"]))
    assert np.max(np.abs(gym.V_chain - g["V_chain"])) < 1e-9 * max(scale, 1.0)
    assert np.max(np.abs(gym.T_chain - g["T_chain"])) < 1e-9 * max(scale, 1.0)
    assert gym.E_chain[0] == 0.0 and gym.V_chain[0] == 0.0
    # recycle p_initial like the script does (RHMC-single-tests.py:43,61-62)
    gym.Nsteps, gym.dt = 20, 0.01
    gym.run_single_RHMC(q_model_0=np.copy(q_model), f_pos=True, solver="implicit", delta=1e-6,
                        p_initial=gym.p_chain[0, :])
    assert gym.q_chain.shape == (21, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat1", "kat2"])
def test_gym_methods_match_reference(name):
    """Every L2 method of the gym against the reference's recorded values."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym

    g = golden(name)
    if name == "kat1":
        gym = multi_gym(dt=0.2, g_xx=1.0, g_ff=1.0)
        gym.V_prior_const = 0.0
    else:
        gym = multi_gym(dt=1e-2, g_xx=0.005, g_ff=25.0, g_ff2=2.0)
        gym.fmin = gym.mag2flux_converter(20.5)
        gym.fmax = gym.mag2flux_converter(15.0)
        gym.use_prior, gym.alpha = True, 1.5
        gym.use_Vc, gym.beta, gym.Vc_r_pow, gym.f_expnt = True, 1e-4, 4.0, np.zeros(2)
    gym.num_rows = gym.num_cols = 32
    gym.D = g["D"]
    q, p = g["q"], g["p"]
    gym.Nobjs, gym.d = q.size // 3, q.size
    assert relerr(gym.V(q, f_pos=True), g["V"]) < RTOL
    assert np.allclose(gym.dVdq(q), g["dVdq"], rtol=1e-9, atol=1e-10 * np.max(np.abs(g["dVdq"])))
    assert relerr(gym.H(q), g["H"]) < RTOL
    H_val, H_grad = gym.H(q, grad=True)  # the reference returns the pair (sampler_RHMC.py:248-258)
    assert relerr(H_val, g["H"]) < RTOL and relerr(H_grad, g["dH"]) < RTOL
    assert relerr(gym.T(p, gym.H(q)), g["T"]) < RTOL
    assert np.allclose(gym.dphidq(q), g["dphidq"], rtol=1e-9, atol=1e-10 * np.max(np.abs(g["dphidq"])))
    if "dtaudq" in g:
        assert np.allclose(gym.dtaudq(q, p), g["dtaudq"], rtol=RTOL, atol=0)
        assert np.allclose(gym.dtaudp(q, p), g["dtaudp"], rtol=RTOL, atol=0)
    assert relerr(gym.H_ff(q[0]), g["H"][0]) < RTOL
    hxx, dhxx = gym.H_xx(q[0], grad=True)  # (value, grad), sampler_RHMC.py:280
    hff, dhff = gym.H_ff(q[0], grad=True)  # (value, grad), sampler_RHMC.py:292
    assert relerr(hxx, g["H"][1]) < RTOL and relerr(dhxx, g["dH"][1]) < RTOL
    assert relerr(hff, g["H"][0]) < RTOL and relerr(dhff, g["dH"][0]) < RTOL
    q1, p1 = gym.RHMC_single_step(np.copy(q), np.copy(p), 1e-6, 1000)
    assert relerr(q1, g["q_traj"][1]) < RTOL and relerr(p1, g["p_traj"][1]) < 1e-9
    assert np.isinf(gym.V(np.concatenate([[1.0], q[1:]]), f_pos=True))  # flux below f_lim


@pytest.mark.gpu
def test_reversible_jump_script():
    """RHMC-big-sim4.py flow in miniature (make_golden.rj_chain): birth/death and split/merge proposals around two
    device RHMC legs; proposal types, accept decisions, star counts and chains must follow the reference."""
    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym, gen_pow_law_sample

    g = golden("rj_chain")
    nobj, nmodel = g["q_true"].shape[0], g["q_model"].shape[0]
    niter, nsteps, dt = int(g["niter"]), int(g["nsteps"]), float(g["dt"])
    gym = multi_gym(dt=0.0, Nsteps=0, g_xx=0.05, g_ff=4.0, g_ff2=4.0)
    np.random.seed(int(g["seed"]))
    gym.num_rows = gym.num_cols = 32
    q_true = np.zeros((nobj, 3))
    q_model = np.zeros((nmodel, 3))
    alpha = 2.0
    fmin = gym.mag2flux_converter(20.0)
    fmax = gym.mag2flux_converter(15.0)
    mag = gym.flux2mag_converter(gen_pow_law_sample(alpha, fmin, fmax, nobj))
    for i in range(nobj):
        x = np.random.random() * (gym.num_rows - 2.0) + 1.0
        y = np.random.random() * (gym.num_cols - 2.0) + 1.0
        q_true[i] = np.array([mag[i], x, y])
    gym.fmin, gym.fmax = fmin, fmax
    gym.K_split, gym.beta_a, gym.beta_b = 1.0, 4.0, 4.0
    gym.use_prior, gym.alpha = True, alpha
    mag = gym.flux2mag_converter(gen_pow_law_sample(alpha, fmin, fmax, nmodel))
    q_model[:, 0] = mag
    q_model[:, 1] = np.random.random(size=nmodel) * (gym.num_rows - 2.0) + 1.0
    q_model[:, 2] = np.random.random(size=nmodel) * (gym.num_cols - 2.0) + 1.0
    gym.gen_mock_data(q_true)
    assert np.array_equal(gym.D, g["D"]) and np.array_equal(q_model, g["q_model"])
    with quiet():
        gym.run_RHMC(np.copy(q_model), f_pos=True, delta=1e-6, Niter=niter, Nsteps=nsteps, dt=dt, save_traj=False,
                     verbose=False, q_true=q_true, P_move=[0.4, 0.3, 0.3], N_max=30)
    assert np.array_equal(gym.move_chain, g["move_chain"])
    assert np.array_equal(gym.N_chain, g["N_chain"])
    assert np.array_equal(gym.A_chain, g["A_chain"])
    assert set(np.unique(g["move_chain"])) == {0, 1, 2, 3, 4} and g["N_chain"].max() > g["N_chain"].min()
    assert relerr(gym.E_chain, g["E_chain"]) < 1e-9
    assert first_divergence(gym.q_chain, g["q_chain"], 1e-7) == -1
    assert np.random.random(1)[0] == g["next_uniform"][0]
