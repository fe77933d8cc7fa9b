#!/usr/bin/env python
"""Benchmark of the RHMC leapfrog hot path (BASELINE.json: "RHMC star-gradient evals/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4]

One "step" = one resident pass of the hot path over one batch of synthetic input: every chain of the batch runs
(niter+1) Metropolis iterations x nsteps generalised-leapfrog steps inside ONE kernel launch.  One unit of the
metric = one star advanced by one RHMC_single_step (reference sampler_RHMC.py:522-566).

Default workload (BASELINE.json configs[1]): 11 magnitudes x 1000 independent one-star 32x32 chains
(README:92-102 of the reference; RHMC-single-full-inference-test.py constants: Nsteps=10, dt=0.2, Niter=1000,
g_xx=g_ff=g_ff2=1, delta=1e-6), each chain on its own Poisson realisation, device Philox draws.

`value`     : units/s with inputs already resident in HBM, kernel time from CUDA events on the launching stream.
`e2e`       : same metric through the public Python API with pinned HOST buffers; H2D of the images/start state and
              D2H of all chain arrays inside the timed region.
`roofline`  : FP64 (or FP32) CUDA-core pipe -- nothing on this path is HBM- or tensor-bound; peak is
              an FMA-chain microbenchmark run live on the same GPU.  `roofline_hbm` gives the HBM view.
`cpu_baseline` / `--impl reference`: the NumPy oracle port of the reference on the box's host cores (the reference
              itself is Python-2 source that cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "RHMC star-gradient evals/sec (leapfrog steps x stars)"
UNIT = "star-steps/s"
MAGS = [15, 16, 17, 18, 19, 20, 21, 21.5, 21.6, 21.7, 21.75]  # reference README:98


# ----------------------------------------------------------------------------------------------- workloads
def exp_constants():
    """default_exp_setup / compute_factors of the reference (sampler_RHMC.py:161-201), evaluated by the oracle's
    helper functions only to obtain the frozen constants."""
    flux_to_count = 1.0 / (0.00546689 * 4.62)
    B = 10 ** (0.4 * (22.5 - 23)) * flux_to_count
    return dict(flux_to_count=flux_to_count, B_count=B, f_lim=B, f_low=10 ** (0.4 * (22.5 - 25)) * flux_to_count,
                psf_fwhm_pix=1.4 / 0.4, g0=0.035997054345069765, g1=0.4523523265306124, g2=0.008141675878296745)


def psf_image(R, C, x, y, fwhm):
    sigma = fwhm / 2.354
    ci = np.arange(0.5, R)[:, None]
    cj = np.arange(0.5, C)[None, :]
    return np.exp(-((ci - x) ** 2 + (cj - y) ** 2) / (2 * sigma**2)) / (2 * np.pi * sigma**2)


def workload_c2(chains_per_mag, seed):
    """One star per 32x32 field, truth at the centre, model start = truth."""
    k = exp_constants()
    R = C = 32
    rng = np.random.RandomState(seed)
    F = chains_per_mag * len(MAGS)
    psf = psf_image(R, C, 16.0, 16.0, k["psf_fwhm_pix"])
    fl = np.repeat([10 ** (0.4 * (22.5 - m)) * k["flux_to_count"] for m in MAGS], chains_per_mag)
    D = rng.poisson(k["B_count"] + fl[:, None, None] * psf[None]).astype(np.float64)
    q0 = np.stack([fl, np.full(F, 16.0), np.full(F, 16.0)], axis=1)
    cfg = dict(n_fields=F, num_rows=R, num_cols=C, max_stars=1, g_xx=1.0, g_ff=1.0, use_prior=False,
               **{n: k[n] for n in ("psf_fwhm_pix", "B_count", "f_lim", "f_low", "g0", "g1", "g2")})
    run = dict(nsteps=10, dt=0.2, g_ff2=1.0, delta=1e-6, counter_max=1000, f_pos=True)
    flops_per_unit = 9 * R * C + 2 * R * C  # F = 9 P^2 + 2 A with P^2 = A = R*C (full-image PSF)
    # pixels the kernel actually visits for a star at the image centre: rows with |i + .5 - x| <= wcut (ex_i >= 2^-46)
    # times the 24-column window of the one-star kernel (chain_kernels.cu: use_column_window)
    sigma = k["psf_fwhm_pix"] / 2.354
    wcut = np.sqrt(46 * np.log(2.0) * 2.0) * sigma
    rows_kept = int(min(R, np.floor(15.5 + wcut) + 1) - max(0, np.ceil(15.5 - wcut)))
    cols_kept = 24 if 144.0 / (2 * sigma**2) >= 46 * np.log(2.0) else C
    flops_executed = 9 * rows_kept * cols_kept + 2 * R * C
    return dict(name="c2_one_star_32x32", D=D, q0=q0, cfg=cfg, run=run, nstars=1, flops_per_unit=flops_per_unit,
                flops_executed_per_unit=flops_executed,
                desc="%d mags x %d one-star 32x32 chains" % (len(MAGS), chains_per_mag))


def workload_c4(n_fields, seed, nstars=204, size=64):
    """Crowded 64x64 fields, 0.05 stars/px, RHMC-big-sim4.py constants (g_xx=.05, g_ff=4, g_ff2=4, dt=5e-2)."""
    k = exp_constants()
    rng = np.random.RandomState(seed)
    alpha = 2.0
    fmin = 10 ** (0.4 * (22.5 - 20)) * k["flux_to_count"]
    fmax = 10 ** (0.4 * (22.5 - 15)) * k["flux_to_count"]
    D = np.empty((n_fields, size, size))
    q0 = np.empty((n_fields, 3 * nstars))
    sigma = k["psf_fwhm_pix"] / 2.354
    ci = np.arange(0.5, size)
    for f in range(n_fields):
        u = rng.random_sample(nstars)
        fl = np.exp(np.log(fmin ** (1 - alpha) + u * (fmax ** (1 - alpha) - fmin ** (1 - alpha))) / (1 - alpha))
        x = rng.random_sample(nstars) * (size - 2.0) + 1.0
        y = rng.random_sample(nstars) * (size - 2.0) + 1.0
        ex = np.exp(-((ci[None, :] - x[:, None]) ** 2) / (2 * sigma**2))
        ey = np.exp(-((ci[None, :] - y[:, None]) ** 2) / (2 * sigma**2)) / (2 * np.pi * sigma**2)
        lam = k["B_count"] + np.einsum("k,ki,kj->ij", fl, ex, ey)
        D[f] = rng.poisson(lam)
        q0[f, 0::3] = fl * 1.05
        q0[f, 1::3] = x + 0.1 * rng.randn(nstars)
        q0[f, 2::3] = y + 0.1 * rng.randn(nstars)
    vpc = np.log(size * size) - np.log((1 - alpha) / (fmax ** (1 - alpha) - fmin ** (1 - alpha)))
    cfg = dict(n_fields=n_fields, num_rows=size, num_cols=size, max_stars=nstars, g_xx=0.05, g_ff=4.0, use_prior=True,
               alpha=alpha, V_prior_const=float(vpc), patch_radius=12,
               **{n: k[n] for n in ("psf_fwhm_pix", "B_count", "f_lim", "f_low", "g0", "g1", "g2")})
    run = dict(nsteps=10, dt=5e-2, g_ff2=4.0, delta=1e-6, counter_max=1000, f_pos=True)
    flops_per_unit = 9 * 625 + 2 * (size * size / nstars)
    return dict(name="c4_crowded_%dx%d_%dstars" % (size, size, nstars), D=D, q0=q0, cfg=cfg, run=run, nstars=nstars,
                flops_per_unit=flops_per_unit, flops_executed_per_unit=flops_per_unit, desc="%d crowded %dx%d fields x %d stars" % (n_fields, size, size, nstars))


def workload_c5(rows, cols, nstars, seed, row0, nrows, rad=12):
    """One large crowded field (BASELINE configs[4] shape: 1.49e-3 stars/px, flux power law alpha=2 in mag [15,20],
    prior on, repulsion off).  Every rank calls this with the same seed and gets the same star list; only the rows
    [row0, row0+nrows) of the data image are rendered (patch-limited) and Poisson-sampled with a per-row stream, so
    overlapping halo rows agree between ranks."""
    k = exp_constants()
    rng = np.random.RandomState(seed)
    alpha = 2.0
    fmin = 10 ** (0.4 * (22.5 - 20)) * k["flux_to_count"]
    fmax = 10 ** (0.4 * (22.5 - 15)) * k["flux_to_count"]
    u = rng.random_sample(nstars)
    fl = np.exp(np.log(fmin ** (1 - alpha) + u * (fmax ** (1 - alpha) - fmin ** (1 - alpha))) / (1 - alpha))
    x = rng.random_sample(nstars) * (rows - 2.0) + 1.0
    y = rng.random_sample(nstars) * (cols - 2.0) + 1.0
    q0 = np.stack([fl * 1.05, x + 0.1 * rng.randn(nstars), y + 0.1 * rng.randn(nstars)], axis=1)
    sigma = k["psf_fwhm_pix"] / 2.354
    lam = np.full((nrows, cols), k["B_count"])
    near = np.nonzero((x > row0 - rad - 1) & (x < row0 + nrows + rad + 1))[0]
    for s in near:
        i0, i1 = max(row0, int(x[s]) - rad), min(row0 + nrows - 1, int(x[s]) + rad)
        j0, j1 = max(0, int(y[s]) - rad), min(cols - 1, int(y[s]) + rad)
        if i0 > i1:
            continue
        ex = np.exp(-((np.arange(i0, i1 + 1) + 0.5 - x[s]) ** 2) / (2 * sigma**2))
        ey = np.exp(-((np.arange(j0, j1 + 1) + 0.5 - y[s]) ** 2) / (2 * sigma**2)) / (2 * np.pi * sigma**2)
        lam[i0 - row0:i1 - row0 + 1, j0:j1 + 1] += fl[s] * ex[:, None] * ey[None, :]
    D = np.empty_like(lam)
    for i in range(nrows):
        D[i] = np.random.RandomState((seed * 1000003 + row0 + i) % (2**32)).poisson(lam[i])
    vpc = np.log(float(rows) * cols) - np.log((1 - alpha) / (fmax ** (1 - alpha) - fmin ** (1 - alpha)))
    consts = dict(g_xx=0.05, g_ff=4.0, use_prior=True, alpha=alpha, V_prior_const=float(vpc),
                  **{n: k[n] for n in ("psf_fwhm_pix", "B_count", "f_lim", "f_low", "g0", "g1", "g2")})
    run = dict(nsteps=10, dt=5e-2, g_ff2=4.0, delta=1e-6, counter_max=1000, f_pos=True)
    A = rows * cols / float(nstars)
    return dict(D=D, q0=q0, consts=consts, run=run, flops_per_unit=9 * (2 * rad + 1) ** 2 + 2 * A,
                bytes_per_unit=8.0 * A + 96.0)


# ----------------------------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(task):
    """One oracle chain on one host core.  Returns (units, seconds)."""
    import stellar_oracle as so  # oracle/ is the checker and, here only, the CPU baseline

    wl_name, seed, niter = task
    if wl_name.startswith("c2"):
        wl = workload_c2(1, seed)
    else:
        wl = workload_c4(1, seed)
    idx = seed % wl["D"].shape[0]
    cfg, run = wl["cfg"], wl["run"]
    S = so.Setup(num_rows=cfg["num_rows"], num_cols=cfg["num_cols"], g_xx=cfg["g_xx"], g_ff=cfg["g_ff"],
                 g_ff2=run["g_ff2"], use_prior=cfg.get("use_prior", False), alpha=cfg.get("alpha", 2.0),
                 V_prior_const=cfg.get("V_prior_const", 0.0), D=wl["D"][idx])
    rng = np.random.RandomState(seed + 17)
    d = wl["q0"].shape[1]
    normals = rng.randn(niter + 1, d)
    lnu = np.log(rng.random_sample(niter + 1))
    t0 = time.perf_counter()
    so.run_rhmc(S, wl["q0"][idx], normals, lnu, niter, run["nsteps"], run["dt"], f_pos=True, delta=run["delta"],
                counter_max=run["counter_max"])
    return (niter + 1) * run["nsteps"] * wl["nstars"], time.perf_counter() - t0


def cpu_arm(wl_name, niter, cores, rounds=1):
    """All host cores, one independent chain per process.  Returns (units/s, wall seconds, units)."""
    ctx = mp.get_context("spawn")
    tasks = [(wl_name, 1000 + i, niter) for i in range(cores * rounds)]
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(wl_name, 1, 1)] * cores)  # start-up (imports) outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    units = sum(r[0] for r in res)
    return units / wall, wall, units


def cpu_niter_for(wl_name, target_seconds=12.0):
    # reference speed measured at survey time: ~2.7k star-steps/s/core (1 star), ~2.1k (204 stars 64x64)
    if wl_name.startswith("c2"):
        return max(20, int(target_seconds * 2500 / 10))
    return max(1, int(target_seconds * 2000 / (10 * 204)))


# ----------------------------------------------------------------------------------------------- GPU helpers
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def traffic_from_profile(wl_name):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh).get(wl_name)
    return None


# ----------------------------------------------------------------------------------------------- main arms
def run_ours(args):
    import torch
    import torch.distributed as dist

    from hmc_stellar_toy_model_b200 import RHMCContext, _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the RHMC path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        niter_cpu = cpu_niter_for(args.workload)
        v, wall, units = cpu_arm(args.workload, niter_cpu, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d processes x 1 chain x %d iterations x 10 steps of the same workload (%.1f s wall, NumPy "
                         "oracle port of the reference)" % (cores, niter_cpu + 1, wall)}

    wl = workload_c2(args.chains_per_mag, 77 + rank) if args.workload == "c2" else workload_c4(args.fields, 77 + rank)
    F, S = wl["D"].shape[0], wl["q0"].shape[1]
    niter = args.niter
    L = niter + 1
    units_per_step = F * L * wl["run"]["nsteps"] * wl["nstars"]
    prec = args.precision
    ctx = RHMCContext(device=local, precision=prec, **wl["cfg"])
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # pinned host buffers for the e2e path
    pin_D = _capi.PinnedBuffer(wl["D"].shape)
    pin_D.array[...] = wl["D"]
    pin_q0 = _capi.PinnedBuffer(wl["q0"].shape)
    pin_q0.array[...] = wl["q0"]
    rows = L
    outs = {"q_chain": _capi.PinnedBuffer((F, rows, S)), "p_chain": _capi.PinnedBuffer((F, rows, S)),
            "E_chain": _capi.PinnedBuffer((F, rows)), "V_chain": _capi.PinnedBuffer((F, rows)),
            "T_chain": _capi.PinnedBuffer((F, rows)), "A_chain": _capi.PinnedBuffer((F, rows), np.uint8),
            "q_final": _capi.PinnedBuffer((F, S)), "accept_rate": _capi.PinnedBuffer((F,))}
    out_arrays = {k: v.array for k, v in outs.items()}

    ctx.set_data(pin_D.array)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def make_args(seed):
        a, keep = ctx.make_run_args(pin_q0.array, niter, seed=seed + 104729 * rank, out=out_arrays, **wl["run"])
        return a, keep

    # ---- value: inputs resident, kernel-only device time
    a, keep = make_args(1)
    ctx.run_upload(a)
    for w in range(args.warmup):
        a, keep = make_args(100 + w)
        ctx.run_launch(a)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for s in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the kernel's own events)
        a, keep = make_args(1000 + s)
        ctx.run_launch(a)
        kernel_ms.append(ctx.last_kernel_ms())
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(kernel_ms))
    region_ms = ev0.elapsed_time(ev1)
    acc = None

    # ---- e2e: public API, host buffers, copies inside the timed region
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ctx.set_data(pin_D.array)
    a, keep = make_args(5)
    ctx.run_prepared(a)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ctx.set_data(pin_D.array)              # H2D images
        a, keep = make_args(2000 + s)
        ctx.run_prepared(a)                    # H2D start state, launch, D2H chains, sync
        acc = float(out_arrays["accept_rate"].mean())
    barrier()
    t_e2e = time.perf_counter() - t0
    h2d = wl["D"].nbytes + wl["q0"].nbytes
    d2h = sum(v.nbytes for v in outs.values())

    # max over ranks
    t = torch.tensor([total_ms, t_e2e, region_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, t_e2e, region_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peak_tf, _ = _capi.measure_fma_peak(local, prec)
        ms_per_step = total_ms / args.steps
        value = world * units_per_step * args.steps / (total_ms * 1e-3)
        e2e_value = world * units_per_step * e2e_steps / t_e2e
        achieved_tf = wl["flops_per_unit"] * units_per_step / (ms_per_step * 1e-3) / 1e12
        peaks, peak_src = measured_peaks()
        chain_bytes = d2h + wl["D"].nbytes * (prec / 64.0)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if prec == 64 else "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "description": wl["desc"], "fields_per_gpu": F, "stars_per_field": wl["nstars"],
                       "niter": niter, "nsteps": wl["run"]["nsteps"], "dt": wl["run"]["dt"],
                       "units_per_step_per_gpu": units_per_step, "rng": "device Philox4x32-10",
                       "l2": "256 MB flush write between timed iterations; kernel is shared-memory resident",
                       "parallelism": "independent chains sharded across %d GPU(s), no communication" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64" if prec == 64 else "fp32", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "peak_source": "FMA-chain microbenchmark on this GPU (srhmc_measure_fma_peak); nominal %s"
                                        % ("37.2" if prec == 64 else "74.4"),
                         "flops_per_unit": wl["flops_per_unit"],
                         "flops_executed_per_unit": wl["flops_executed_per_unit"],
                         "frac_executed": achieved_tf * wl["flops_executed_per_unit"] / wl["flops_per_unit"] / peak_tf,
                         "traffic": traffic_from_profile(wl["name"]),
                         "note": "CUDA-core pipe bound: images stay in shared memory for the whole launch, "
                                 "no dense contraction, so neither the HBM nor the tensor roofline applies"},
            "roofline_hbm": {"bound": "hbm", "achieved": chain_bytes / (ms_per_step * 1e-3) / 1e9,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peak_src,
                             "frac": chain_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "bytes_per_launch": int(chain_bytes)},
            "cpu_baseline": cpu,
            "accept_rate": acc,
            "timed_region_ms": region_ms, "wall_s": t_wall,
        }
        print(json.dumps(out))
    for b in list(outs.values()) + [pin_D, pin_q0]:
        b.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_big(args):
    """BASELINE configs[4]-shaped workload: ONE large field, row strips over the ranks, ghost-star all-gather and the
    energy / fixed-point all-reduces through NCCL on the compute stream."""
    import torch
    import torch.distributed as dist

    from hmc_stellar_toy_model_b200 import bigfield as bf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the RHMC path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols = args.rows * (world if args.weak else 1), args.cols
    nstars = int(args.stars * (world if args.weak else 1))
    rad, halo = 12, 24
    lo, hi = bf.strip_bounds(rows, world)[rank]
    row0, nrows = bf.data_window(rows, lo, hi, halo)
    wl = workload_c5(rows, cols, nstars, 77, row0, nrows, rad)
    strip = bf.BigFieldStrip(rows=rows, cols=cols, rank=rank, world=world, device=local, max_stars=nstars,
                             max_ghosts=max(1024, int(4 * nstars * (halo + rad + 1) / max(1, hi - lo))), patch_radius=rad,
                             halo=halo, **wl["consts"])
    stream = torch.cuda.Stream()   # kernels, NCCL collectives and the timing events all ride on this stream
    torch.cuda.set_stream(stream)
    strip.set_stream(stream.cuda_stream)
    if world == 1:
        comm = bf.NoComm()
    elif args.comm == "nccl":
        comm = bf.TorchDistComm(dist)   # NCCL collectives issued by the caller between the phases (eager launches)
    else:
        comm = bf.PeerComm([strip], dist)  # the library's own exchange kernels over NVLink peer memory; graph replay
    eng = bf.BigFieldRHMC([strip], comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    niter = args.niter
    units = nstars * (niter + 1) * wl["run"]["nsteps"]
    strip.set_data_window(wl["D"])
    launches0 = None
    for w in range(args.warmup):
        strip.set_stars(wl["q0"])
        eng.run(niter, seed=100 + w, **wl["run"])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    acc = 0.0
    launches0 = strip.launch_count
    t_e2e = 0.0
    for s in range(args.steps):
        strip.set_stars(wl["q0"])
        barrier()
        ev0.record(stream)
        out = eng.run(niter, seed=1000 + s, **wl["run"])  # read_chains synchronises
        ev1.record(stream)
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
        acc = out["accept_rate"]
    launches = strip.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: host buffers in, final stars out
    for s in range(max(1, min(args.steps, args.e2e_steps))):
        barrier()
        t0 = time.perf_counter()
        strip.set_data_window(wl["D"])
        strip.set_stars(wl["q0"])
        eng.run(niter, seed=2000 + s, **wl["run"])
        strip.get_stars()
        barrier()
        t_e2e += time.perf_counter() - t0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t = torch.tensor([sum(times), t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, t_e2e = [float(v) for v in t.tolist()]
    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = units * args.steps / (total_ms * 1e-3)
        gbs = wl["bytes_per_unit"] * units / world / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c5_tiled_field_%dx%d_%dstars" % (rows, cols, nstars), "rows": rows, "cols": cols,
                       "stars": nstars, "niter": niter, "nsteps": wl["run"]["nsteps"], "dt": wl["run"]["dt"],
                       "patch_radius": rad, "halo_rows": halo, "rng": "device Philox4x32-10",
                       "l2": "data window of %.0f MB per rank exceeds nothing smaller than L2 only when > 126 MB; "
                             "no flush between launches" % (wl["D"].nbytes / 1e6),
                       "parallelism": "row strips over %d GPU(s); per step: 1 exchange of boundary stars with the two "
                                      "neighbours, 2 max all-reduces; per iteration: 2 sum all-reduces of 8 doubles (%s)"
                                      % (world, "none: single GPU" if world == 1 else
                                         ("own kernels over NVLink peer memory, CUDA-graph replay" if args.comm == "peer"
                                          else "NCCL via torch.distributed, eager"))},
            "clocks": clocks,
            "e2e": {"value": units * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(wl["D"].nbytes + wl["q0"].nbytes),
                    "d2h_bytes_per_step": int(3 * wl["q0"].nbytes // max(1, world)), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                         "peak_source": peak_src, "bytes_per_unit": wl["bytes_per_unit"],
                         "traffic": traffic_from_profile("c5_tiled_field_%dx%d_%dstars" % (rows, cols, nstars)) if world == 1 else None,
                         "note": "per GPU, over the WHOLE leapfrog step (tile kernel + the three per-star kernels); algorithmic "
                                 "bytes = the data strip read once per gradient + star state"},
            "cpu_baseline": None, "accept_rate": acc,
        }
        print(json.dumps(out))
    strip.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path (NumPy oracle port; the Python-2 reference itself cannot
    travel to the GPU box) on all host cores, same config/metric.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    niter = max(10, cpu_niter_for(args.workload, 4.0))
    for _ in range(args.warmup):
        cpu_arm(args.workload, max(2, niter // 10), cores)
    t_total, units_total = 0.0, 0
    for _ in range(args.steps):
        v, wall, units = cpu_arm(args.workload, niter, cores)
        t_total += wall
        units_total += units
    value = units_total / t_total
    wl_name = "c2_one_star_32x32" if args.workload == "c2" else "c4_crowded_64x64_204stars"
    sample = "%d processes x 1 chain x %d iterations x 10 steps per step" % (cores, niter + 1)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": wl_name, "sample": sample},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--rows", type=int, default=8192, help="c5: image rows (per GPU with --weak); BASELINE configs[4] is 8192 x 8192")
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--stars", type=float, default=100000, help="c5: stars (per GPU with --weak); 1.49e-3 per pixel")
    ap.add_argument("--weak", action="store_true", help="c5: grow the field with the GPU count")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="c5 on several GPUs: the library's own peer-memory exchange kernels, or NCCL through torch.distributed")
    ap.add_argument("--chains-per-mag", type=int, default=1000)
    ap.add_argument("--fields", type=int, default=592)
    ap.add_argument("--niter", type=int, default=1000)
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        if args.niter == 1000:
            args.niter = 4
        run_big(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
