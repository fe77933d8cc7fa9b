import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import numpy as np
import stellar_oracle as so
from helpers import golden, setup_from
from test_gpu_parity import make_ctx
for name in ["chain_multi30_vc", "chain_multi100"]:
    g = golden(name); S = setup_from(g); q0 = so.format_q(S, g["q_model"]); n = q0.size // 3
    with make_ctx(S, max_stars=n) as ctx:
        ctx.set_data(S.D)
        r = ctx.run(q0[None], int(g["niter"]), int(g["nsteps"]), float(g["dt"]), normals=g["normals"][None], lnu=g["lnu"][None],
                    g_ff2=S.g_ff2, beta=S.beta, f_pos=True, schedule_g_ff2=g["schedule_g_ff2"], schedule_beta=g["schedule_beta"])
    qc = g["q_chain"][:, :3*n]
    for l in range(qc.shape[0]):
        e = np.abs(r.q_chain[0][l] - qc[l]) / np.abs(qc[l])
        print(name, l, "A", int(r.A_chain[0][l]), int(g["A_chain"][l]), "max rel q err %.2e" % e.max(), "arg", e.argmax(), "E err %.2e" % abs(r.E_chain[0][l]-g["E_chain"][l]))
