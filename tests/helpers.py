"""Shared helpers for the parity tests: load golden fixtures, build oracle setups."""
import os

import numpy as np

import stellar_oracle as so

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def setup_from(g, **over):
    """Oracle Setup carrying the gym state recorded in a golden fixture."""
    S = so.Setup()
    for k in ("B_count", "PSF_FWHM_pix", "f_lim", "mB", "flux_to_count", "g0", "g1", "g2", "g_xx", "g_ff",
              "g_ff2", "alpha", "beta", "Vc_r_pow", "V_prior_const"):
        setattr(S, k, float(g[k]))
    S.num_rows = int(g["num_rows"])
    S.num_cols = int(g["num_cols"])
    S.use_prior = bool(g["use_prior"])
    S.use_Vc = bool(g["use_Vc"])
    S.D = np.array(g["D"], dtype=float)
    if "dt" in g:
        S.dt = float(g["dt"])
    for k, v in over.items():
        setattr(S, k, v)
    return S


def relerr(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def first_divergence(a, b, rtol):
    """Index of the first row where two chains differ by more than rtol (or -1)."""
    a = np.asarray(a, dtype=float).reshape(len(a), -1)
    b = np.asarray(b, dtype=float).reshape(len(b), -1)
    bad = np.any(np.abs(a - b) > rtol * np.maximum(np.abs(b), 1e-300), axis=1)
    idx = np.nonzero(bad)[0]
    return int(idx[0]) if idx.size else -1
