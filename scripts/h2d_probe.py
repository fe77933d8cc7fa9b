import time, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from hmc_stellar_toy_model_b200 import _capi, bigfield as bf
env = bench.Env()
t = bench.c5_truth(8192, 8192, 100000, 77)
strip = bench._make_strip(env, bf, 8192, 8192, 100000, t["consts"], 24, 12)
pin = _capi.PinnedBuffer((strip.nrows, 8192))
pin.array[...] = 1.0
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); strip.set_data_window(pin.array); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("set_data_window %.2f ms  %.1f GB/s" % (dt * 1e3, pin.nbytes / dt / 1e9))
x = torch.empty(pin.nbytes // 8, dtype=torch.float64, device="cuda")
h = torch.empty(pin.nbytes // 8, dtype=torch.float64).pin_memory()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("torch pinned H2D %.2f ms  %.1f GB/s" % (dt * 1e3, pin.nbytes / dt / 1e9))
for i in range(2):
    t0 = time.perf_counter(); strip.set_stars(t["q0"]); dt = time.perf_counter() - t0
    print("set_stars %.2f ms" % (dt * 1e3))
t0 = time.perf_counter(); strip.get_stars(); print("get_stars %.2f ms" % ((time.perf_counter() - t0) * 1e3))
