"""Parity of the compact-table evaluation of the CTA-per-field kernel (field_kernel.cuh "v3": patch-limited PSF, all stars
in one table build, pair-wise gather) against the NumPy oracle's patch-limited restatement (oracle.patch_eval, which
follows sampler_RHMC.py:294-351 / 365-425 with every PSF cut to 25x25 pixels), against the reference's own full-image
values, and against the chunked-table path of the same kernel (SRHMC_FIELD_V3=0)."""
import numpy as np
import pytest

import stellar_oracle as so
from helpers import golden, relerr, setup_from

pytestmark = pytest.mark.gpu


def _ctx(S, n_fields, max_stars, **kw):
    from test_gpu_parity import make_ctx

    return make_ctx(S, n_fields=n_fields, max_stars=max_stars, **kw)


def _grad_close(a, b, tol, floor=1e-6):
    a, b = np.asarray(a).reshape(-1, 3), np.asarray(b).reshape(-1, 3)
    scale = np.maximum(np.abs(b), floor * np.max(np.abs(b), axis=0, keepdims=True))
    return float(np.max(np.abs(a - b) / scale))


def _crowded(rows, cols, n, seed, edge=False):
    """Setup + Poisson data + start state of a crowded field at the RHMC-big-sim4 constants."""
    rng = np.random.RandomState(seed)
    S = so.Setup(num_rows=rows, num_cols=cols, dt=5e-2, g_xx=0.05, g_ff=4.0, g_ff2=4.0, use_prior=True, alpha=2.0,
                 V_prior_const=0.7)
    fmin, fmax = S.mag2flux_converter(20.0), S.mag2flux_converter(15.0)
    f = so.pow_law_sample(2.0, fmin, fmax, rng.random_sample(n))
    x = rng.uniform(1, rows - 1, n)
    y = rng.uniform(1, cols - 1, n)
    if edge:  # stars on, next to and just outside every image edge and in the corners
        x[:8] = [0.3, rows - 0.4, 15.5, 0.0, rows - 1.0, -0.6, rows + 0.4, 0.2]
        y[:8] = [cols - 0.4, 0.4, 0.0, 10.2, cols - 1.0, 5.5, 7.7, 0.1]
    q_true = np.stack([f, x, y], axis=1)
    D = rng.poisson(_model(S, q_true, rows, cols)).astype(float)
    q0 = q_true * np.array([1.05, 1.0, 1.0]) + np.concatenate([np.zeros((n, 1)), 0.1 * rng.randn(n, 2)], axis=1)
    return S, D, q0


def _model(S, q, rows, cols):
    sig2 = (S.PSF_FWHM_pix / 2.354) ** 2
    ci, cj = np.arange(0.5, rows), np.arange(0.5, cols)
    ex = np.exp(-((ci[None, :] - q[:, 1:2]) ** 2) / (2 * sig2))
    ey = np.exp(-((cj[None, :] - q[:, 2:3]) ** 2) / (2 * sig2)) / (2 * np.pi * sig2)
    return S.B_count + np.einsum("k,ki,kj->ij", q[:, 0], ex, ey)


@pytest.mark.parametrize("rows,cols,n,edge", [(64, 64, 204, False), (64, 64, 60, True), (40, 26, 9, True), (25, 48, 9, True)])
def test_compact_eval_matches_patch_oracle(rows, cols, n, edge):
    S, D, q0 = _crowded(rows, cols, n, 3 + n, edge)
    Vo, go = so.patch_eval(S, D, q0, rad=12)
    go = go.copy()
    go[:, 0] += S.alpha / q0[:, 0]      # the context returns dV/dq including the prior term (sampler_RHMC.py:408-409)
    with _ctx(S, 1, n, patch_radius=12) as ctx:
        ctx.set_data(D)
        V, grad, _, _ = ctx.eval(q0.ravel()[None], f_pos=False, g_ff2=S.g_ff2)
    assert relerr(V[0], Vo) < 1e-12
    assert _grad_close(grad[0], go, 1e-10) < 1e-10


def test_compact_path_is_the_one_running_and_equals_chunked_path(monkeypatch):
    """Same field through the compact-table path and the chunked-table path of the kernel (SRHMC_FIELD_V3=0): identical
    physics, different summation order -> 1e-12; three leapfrog steps with identical fixed-point counts; a ragged batch
    (0, 1, odd and full star counts) in one launch."""
    S, D, q0 = _crowded(64, 64, 204, 11)
    rng = np.random.RandomState(5)
    counts = [204, 0, 1, 77, 204]
    F = len(counts)
    Db = np.stack([rng.poisson(_model(S, q0[:max(c, 1)], 64, 64)).astype(float) for c in counts])
    q = np.zeros((F, 3 * 204))
    p = np.zeros((F, 3 * 204))
    for f, c in enumerate(counts):
        q[f, : 3 * c] = q0[:c].ravel()
        H = so.metric(S, q[f, : 3 * c]) if c else np.zeros(0)
        p[f, : 3 * c] = rng.randn(3 * c) * np.sqrt(H)
    res = {}
    for tag, flag in (("compact", "1"), ("chunked", "0")):
        monkeypatch.setenv("SRHMC_FIELD_V3", flag)
        with _ctx(S, F, 204, patch_radius=12) as ctx:
            ctx.set_data(Db)
            V, grad, _, _ = ctx.eval(q, nstars=counts, f_pos=True, g_ff2=S.g_ff2)
            q1, p1, cnt = ctx.step(q, p, 3, S.dt, g_ff2=S.g_ff2, nstars=counts, return_counts=True)
        res[tag] = (V, grad, q1, p1, cnt)
    a, b = res["compact"], res["chunked"]
    assert relerr(a[0], b[0]) < 1e-13
    for f, c in enumerate(counts):
        if c:
            assert _grad_close(a[1][f, : 3 * c], b[1][f, : 3 * c], 1e-11) < 5e-10
            assert relerr(a[2][f, : 3 * c], b[2][f, : 3 * c]) < 1e-11
            assert _grad_close(a[3][f, : 3 * c], b[3][f, : 3 * c], 1e-10) < 1e-10
    assert np.array_equal(a[4], b[4])
    # and against the oracle stepping with the same truncation is covered by the eval test; here the full-image reference
    # golden bounds the truncation itself
    g = golden("field_eval_204")
    Sg = setup_from(g)
    monkeypatch.setenv("SRHMC_FIELD_V3", "1")
    with _ctx(Sg, 1, 204, patch_radius=12) as ctx:
        ctx.set_data(Sg.D)
        qs, ps, cnts = ctx.step(g["q"][None], g["p"][None], 1, float(g["dt"]), g_ff2=Sg.g_ff2, return_counts=True)
    counters = []
    so.rhmc_step(Sg, g["q"], g["p"], 1e-6, 1000, counters=counters)
    assert relerr(qs[0], g["q1"]) < 1e-10 and tuple(cnts[0]) == counters[0]


def test_compact_chain_equals_chunked_chain(monkeypatch):
    """A Philox chain of 3 crowded fields: both table paths take the same accept decisions and agree on the energies."""
    S, D, q0 = _crowded(64, 64, 120, 21)
    F = 3
    rng = np.random.RandomState(9)
    Db = np.stack([rng.poisson(_model(S, q0, 64, 64)).astype(float) for _ in range(F)])
    qb = np.repeat(q0.ravel()[None], F, axis=0)
    out = {}
    for tag, flag in (("compact", "1"), ("chunked", "0")):
        monkeypatch.setenv("SRHMC_FIELD_V3", flag)
        with _ctx(S, F, 120, patch_radius=12) as ctx:
            ctx.set_data(Db)
            out[tag] = ctx.run(qb, 8, 6, S.dt, seed=77, g_ff2=S.g_ff2, f_pos=True)
    a, b = out["compact"], out["chunked"]
    assert np.array_equal(a.A_chain, b.A_chain) and a.A_chain.sum() > 0
    assert relerr(a.E_chain, b.E_chain) < 1e-9
    first = int(np.argmax(a.A_chain[0])) + 1
    assert relerr(a.q_chain[0][: first + 1], b.q_chain[0][: first + 1]) < 1e-10


def test_compact_fp32_build_within_1e4():
    """FP32 pixels through the compact-table path: V to 1e-4, gradients to 1e-4 of the per-coordinate scale."""
    g = golden("field_eval_204")
    S = setup_from(g)
    with _ctx(S, 1, 204, patch_radius=8, precision=32) as ctx:
        ctx.set_data(S.D)
        V, grad, _, _ = ctx.eval(g["q"], f_pos=True, g_ff2=S.g_ff2)
    assert relerr(V[0], g["V"]) < 1e-4
    a, b = grad[0].reshape(-1, 3), g["dVdq"].reshape(-1, 3)
    assert np.max(np.abs(a - b) / np.max(np.abs(b), axis=0, keepdims=True)) < 1e-4
