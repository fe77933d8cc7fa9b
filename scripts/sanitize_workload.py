"""Small invocations of every kernel family for compute-sanitizer (scripts/sanitize.sh): the one-star chain kernel with
the chunked work scheduler (37 chains, spin-wait hand-over between warps), the crowded-field kernel on both table paths,
the lightsource Hessian path, the large-field engine (tile kernel with TMA staging) and a 3-strip tiling of one field
through the peer-memory exchange kernels (mailboxes, system-scope fences)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import stellar_oracle as so  # noqa: E402
from helpers import golden, setup_from  # noqa: E402
from hmc_stellar_toy_model_b200 import RHMCContext  # noqa: E402
from hmc_stellar_toy_model_b200 import bigfield as bf  # noqa: E402


def consts(S):
    return dict(psf_fwhm_pix=S.PSF_FWHM_pix, B_count=S.B_count, f_lim=S.f_lim, f_low=S.mag2flux_converter(S.mB + 2), g0=S.g0,
                g1=S.g1, g2=S.g2, g_xx=S.g_xx, g_ff=S.g_ff)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    g = golden("chain_one_star_m19")
    S = setup_from(g)
    q0 = so.format_q(S, g["q_model"])
    if which in ("all", "chain"):
        # 37 one-star chains, scheduler forced to 3 iteration chunks per chain (hand-over through global memory)
        os.environ["SRHMC_CHAIN_CHUNKS"] = "3"
        F = 37
        rng = np.random.RandomState(1)
        D = rng.poisson(np.repeat(so.model_image(S, q0)[None], F, axis=0)).astype(float)
        with RHMCContext(n_fields=F, num_rows=32, num_cols=32, max_stars=1, **consts(S)) as ctx:
            ctx.set_data(D)
            r = ctx.run(np.repeat(q0[None], F, axis=0), 8, 10, 0.2, seed=3, g_ff2=S.g_ff2)
            ctx.eval(np.repeat(q0[None], F, axis=0), f_pos=True)
            ctx.step(np.repeat(q0[None], F, axis=0), np.zeros((F, 3)), 2, 0.2)
        assert np.all(np.isfinite(r.E_chain))
        del os.environ["SRHMC_CHAIN_CHUNKS"]
        print("chain: ok", flush=True)
    g2 = golden("field_eval_204")
    S2 = setup_from(g2)
    kw2 = dict(consts(S2), use_prior=True, alpha=S2.alpha, V_prior_const=S2.V_prior_const)
    if which in ("all", "field"):
        for rad in (12, 0):   # compact-table path / chunked full-image path
            with RHMCContext(n_fields=2, num_rows=64, num_cols=64, max_stars=204, patch_radius=rad, **kw2) as ctx:
                ctx.set_data(np.repeat(S2.D[None], 2, axis=0))
                q = np.repeat(g2["q"][None], 2, axis=0)
                ctx.eval(q, f_pos=True, g_ff2=S2.g_ff2)
                ctx.run(q, 1, 2, 5e-2, seed=1, g_ff2=S2.g_ff2, want=("E", "A"))
        with RHMCContext(n_fields=1, num_rows=32, num_cols=32, max_stars=1, enable_hessian=True, **consts(S)) as ctx:
            ctx.set_data(S.D)
            ctx.hessian(q0[None], np.array([[0.02, -0.3, 0.2]]))
        print("field: ok", flush=True)
    if which in ("all", "big"):
        os.environ["SRHMC_BIG_PATH"] = "tile"
        rng = np.random.RandomState(8)
        rows, cols, n = 230, 128, 300
        Sb = so.Setup(num_rows=rows, num_cols=cols, g_xx=0.05, g_ff=4.0, g_ff2=4.0, use_prior=True, alpha=2.0, V_prior_const=0.5)
        f = so.pow_law_sample(2.0, Sb.mag2flux_converter(20.0), Sb.mag2flux_converter(15.0), rng.random_sample(n))
        q = np.stack([f, rng.uniform(1, rows - 1, n), rng.uniform(1, cols - 1, n)], axis=1)
        kb = dict(consts(Sb), use_prior=True, alpha=2.0, V_prior_const=0.5)
        one = bf.BigFieldStrip(rows=rows, cols=cols, rank=0, world=1, device=0, max_stars=n, max_ghosts=n, patch_radius=12, halo=20, **kb)
        D = one.gen_mock_data(q, seed=3, return_data=True)
        one.set_stars(q)
        a = bf.BigFieldRHMC([one]).run(2, 2, 2e-2, seed=5, g_ff2=4.0, use_graph=False)
        one.close()
        strips = []
        for r in range(3):
            s = bf.BigFieldStrip(rows=rows, cols=cols, rank=r, world=3, device=0, max_stars=n, max_ghosts=n, patch_radius=12, halo=20, **kb)
            s.set_data(D)
            s.set_stars(q)
            strips.append(s)
        b = bf.BigFieldRHMC(strips, bf.PeerComm(strips)).run(2, 2, 2e-2, seed=5, g_ff2=4.0, use_graph=False)
        assert np.array_equal(a["A_chain"], b["A_chain"]) and np.max(np.abs(a["E_chain"] - b["E_chain"]) / np.abs(a["E_chain"])) < 1e-10
        for s in strips:
            s.close()
        print("big: ok", flush=True)


if __name__ == "__main__":
    main()
