import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import stellar_oracle as so
from test_bigfield import _engine
S = so.Setup(num_rows=256, num_cols=96, g_xx=0.05, g_ff=4.0, g_ff2=4.0, use_prior=True, alpha=2.0, V_prior_const=1.0)
rng = np.random.RandomState(3); n = 700
fl = S.mag2flux_converter(rng.uniform(15.5, 20.0, n))
q = np.stack([fl, rng.uniform(1, 255, n), rng.uniform(1, 95, n)], axis=1)
sig = S.PSF_FWHM_pix / 2.354
ci, cj = np.arange(0.5, 256), np.arange(0.5, 96)
ex = np.exp(-((ci[None] - q[:, 1:2]) ** 2) / (2 * sig ** 2)); ey = np.exp(-((cj[None] - q[:, 2:3]) ** 2) / (2 * sig ** 2)) / (2 * np.pi * sig ** 2)
D = rng.poisson(S.B_count + np.einsum("k,ki,kj->ij", q[:, 0], ex, ey)).astype(float)
q0 = q * np.array([1.03, 1.0, 1.0]) + np.array([0.0, 0.05, -0.05])
for world in (1, 4):
    for ug in (False, True):
        eng = _engine(S, q0.ravel(), world=world, D=D, halo=20)
        out = eng.run(6, 5, 2e-2, f_pos=True, g_ff2=4.0, seed=11, use_graph=ug)
        print(world, ug, out["A_chain"], out["E_chain"][:4], [s.views()["counters"].tolist() for s in eng.strips][:1])
print("---- manual replay")
from hmc_stellar_toy_model_b200 import bigfield as bf
eng = _engine(S, q0.ravel(), world=1, D=D, halo=20)
s = eng.strips[0]
L = 5
s.set_draws(None, None, L); s.alloc_chains(L)
st0 = eng._step_struct(2e-2, 1e-6, 4.0, 1000, True, 0, 11)
eng._all("EVAL_V", st0); eng._all("RESET_ITER", st0)
stg = eng._step_struct(2e-2, 1e-6, 4.0, 1000, True, -1, 11)
run_stream = torch.cuda.current_stream()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, capture_error_mode="thread_local"):
    s.adopt_stream(torch.cuda.current_stream().cuda_stream)
    eng._iteration(stg, 5)
s.adopt_stream(run_stream.cuda_stream)
for k in range(L):
    graph.replay(); torch.cuda.synchronize()
    print(k, s.views()["counters"].tolist(), s.read_scalars()[:4])
print(s.read_chains(L))
