"""FP32 build (BASELINE north_star: "<= 1e-4 in the FP32 build") of the one-star chain kernel and the large-field tile
kernel: gradient evaluations run in float (float tables, MUFU.RCP, float column sums) with FP64 star state; every
evaluation that feeds the Metropolis test stays FP64.  Checked against the NumPy oracle and against the FP64 build."""
import numpy as np
import pytest

import stellar_oracle as so
from helpers import golden, relerr, setup_from
from test_gpu_parity import make_ctx

pytestmark = pytest.mark.gpu


def _batch(F, seed):
    g = golden("chain_one_star_m19")
    S = setup_from(g)
    rng = np.random.RandomState(seed)
    mags = rng.uniform(16.0, 21.5, F)
    q = np.stack([[S.mag2flux_converter(m) for m in mags], rng.uniform(12, 20, F), rng.uniform(12, 20, F)], axis=1)
    D = np.stack([rng.poisson(so.model_image(S, qi)).astype(float) for qi in q])
    return S, D, q


def test_fp32_one_star_steps_within_1e4_of_the_oracle():
    S, D, q = _batch(24, 2)
    rng = np.random.RandomState(3)
    p = np.stack([rng.randn(3) * np.sqrt(so.metric(S, qi)) for qi in q])
    with make_ctx(S, n_fields=len(q), max_stars=1, precision=32) as ctx:
        ctx.set_data(D)
        V, grad, H, _ = ctx.eval(q, f_pos=True, g_ff2=S.g_ff2)
        q1, p1 = ctx.step(q, p, 10, 0.2, g_ff2=S.g_ff2)
    for i in range(len(q)):
        Si = S.clone(D=D[i], dt=0.2)
        assert relerr(V[i], so.potential(Si, q[i], True)) < 1e-4
        gi = so.grad_potential(Si, q[i])
        assert np.max(np.abs(grad[i] - gi)) / np.max(np.abs(gi)) < 1e-4
        qq, pp = q[i].copy(), p[i].copy()
        for _ in range(10):
            qq, pp = so.rhmc_step(Si, qq, pp)
        assert relerr(q1[i], qq) < 1e-4, i
        assert np.max(np.abs(p1[i] - pp) / np.maximum(np.abs(pp), 1e-3 * np.max(np.abs(pp)))) < 1e-3, i


def test_fp32_chains_agree_with_fp64_chains():
    """800 chains x 150 iterations with the same Philox draws: the FP32 build takes (almost) the same decisions, and the
    acceptance rate and the posterior mean / standard deviation of (f, x, y) per chain agree with the FP64 chains."""
    S, D, q = _batch(800, 5)
    out = {}
    for prec in (64, 32):
        with make_ctx(S, n_fields=len(q), max_stars=1, precision=prec) as ctx:
            ctx.set_data(D)
            out[prec] = ctx.run(q, 150, 10, 0.2, seed=11, g_ff2=S.g_ff2, f_pos=True, want=("q", "E", "A"))
    a, b = out[64], out[32]
    same = np.mean(a.A_chain == b.A_chain)
    assert same > 0.995, same
    assert abs(a.accept_rate.mean() - b.accept_rate.mean()) < 2e-3
    ma, mb = a.q_chain[:, 30:].mean(axis=1), b.q_chain[:, 30:].mean(axis=1)
    sa, sb = a.q_chain[:, 30:].std(axis=1), b.q_chain[:, 30:].std(axis=1)
    # chains whose decisions are identical track each other to FP32 rounding; the few that split stay within the posterior
    assert np.median(np.abs(ma - mb) / np.maximum(sa, 1e-12)) < 1e-3
    assert np.all(np.abs(ma - mb) <= 3.0 * np.maximum(sa, sb) + 1e-9)
    assert np.median(np.abs(sa - sb) / np.maximum(sa, 1e-12)) < 1e-2
    assert np.all(np.isfinite(b.E_chain))


def test_fp32_tile_kernel_within_1e4_of_the_fp64_build():
    """Large-field engine, 1600 x 1600 / 2000 stars (tile path): gradients of the FP32 tile kernel against the FP64 one
    (which is checked against the oracle in test_bigfield.py) to 1e-4 of the per-coordinate scale, three leapfrog steps
    to 1e-5 in q, and a Philox chain with the same decisions and energies (the energies stay FP64)."""
    import torch

    from hmc_stellar_toy_model_b200 import bigfield as bf
    from test_bigfield import _consts, _synthetic_field

    S, D, q0 = _synthetic_field(1600, 1600, 2000, 9)
    n = len(q0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    res = {}
    for prec in (64, 32):
        s = bf.BigFieldStrip(rows=1600, cols=1600, rank=0, world=1, device=0, max_stars=n, max_ghosts=1, patch_radius=12,
                             halo=20, **_consts(S))
        s.set_stream(stream.cuda_stream)
        s.set_data(D)
        s.set_precision(prec)
        s.set_stars(q0)
        eng = bf.BigFieldRHMC([s])
        eng.evaluate(want_V=False, g_ff2=4.0)
        g = eng.stars(n)[2].copy()
        rng = np.random.RandomState(3)
        s.set_momenta(rng.randn(n, 3) * np.sqrt(so.metric(S, q0.ravel()).reshape(n, 3)))
        eng.steps(3, 5e-2, g_ff2=4.0)
        q3 = eng.stars(n)[0].copy()
        s.set_stars(q0)
        chain = eng.run(6, 5, 5e-2, f_pos=True, g_ff2=4.0, seed=4)
        res[prec] = (g, q3, chain)
        s.close()
    torch.cuda.set_stream(torch.cuda.default_stream())
    g64, q64, c64 = res[64]
    g32, q32, c32 = res[32]
    assert np.max(np.abs(g32 - g64) / np.max(np.abs(g64), axis=0, keepdims=True)) < 1e-4
    assert relerr(q32, q64) < 1e-5
    assert np.array_equal(c32["A_chain"], c64["A_chain"])
    assert relerr(c32["E_chain"], c64["E_chain"]) < 1e-6
