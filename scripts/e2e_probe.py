"""Where does the end-to-end time of the headline workload go?  (host buffers pinned, as in bench.py's e2e leg)

    python scripts/e2e_probe.py [--niter 1000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from hmc_stellar_toy_model_b200 import RHMCContext, _capi  # noqa: E402


def timed(fn, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--niter", type=int, default=1000)
    args = ap.parse_args()
    wl = bench.workload_c2(1000, 77)
    F, S = wl["D"].shape[0], wl["q0"].shape[1]
    L = args.niter + 1
    ctx = RHMCContext(device=0, precision=64, **wl["cfg"])
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    pin_D = _capi.PinnedBuffer(wl["D"].shape)
    pin_D.array[...] = wl["D"]
    pin_q0 = _capi.PinnedBuffer(wl["q0"].shape)
    pin_q0.array[...] = wl["q0"]
    outs = {"q_chain": _capi.PinnedBuffer((F, L, S)), "p_chain": _capi.PinnedBuffer((F, L, S)),
            "E_chain": _capi.PinnedBuffer((F, L)), "V_chain": _capi.PinnedBuffer((F, L)),
            "T_chain": _capi.PinnedBuffer((F, L)), "A_chain": _capi.PinnedBuffer((F, L), np.uint8),
            "q_final": _capi.PinnedBuffer((F, S)), "accept_rate": _capi.PinnedBuffer((F,))}
    out_arrays = {k: v.array for k, v in outs.items()}
    d2h = sum(v.nbytes for v in out_arrays.values())
    ctx.set_data(pin_D.array)
    a, keep = ctx.make_run_args(pin_q0.array, args.niter, seed=1, out=out_arrays, **wl["run"])
    ctx.run_prepared(a)

    print("d2h bytes %.1f MB, h2d %.1f MB" % (d2h / 1e6, (wl["D"].nbytes + wl["q0"].nbytes) / 1e6))
    print("set_data            %8.2f ms" % timed(lambda: ctx.set_data(pin_D.array))[0])
    print("make_run_args       %8.2f ms" % timed(lambda: ctx.make_run_args(pin_q0.array, args.niter, seed=1, out=out_arrays, **wl["run"]))[0])
    for parts in (4, 2, 3, 6, 8, 1):
        os.environ["SRHMC_RUN_PARTS"] = str(parts)
        print("run (%d part(s))      %8.2f ms" % (parts, timed(lambda: ctx.run_prepared(a))[0]))
    print("  upload            %8.2f ms" % timed(lambda: ctx.run_upload(a))[0])
    print("  launch            %8.2f ms   (kernel events: %.2f ms)" % (timed(lambda: ctx.run_launch(a))[0], ctx.last_kernel_ms()))
    print("  download          %8.2f ms" % timed(lambda: ctx.run_download(a))[0])
    del os.environ["SRHMC_RUN_PARTS"]
    print("accept_rate.mean()  %8.2f ms" % timed(lambda: float(out_arrays["accept_rate"].mean()))[0])
    # raw PCIe reference: one pinned copy of the same size each way
    dev = torch.empty(d2h, dtype=torch.uint8, device="cuda")
    host = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    t, _ = timed(lambda: host.copy_(dev, non_blocking=True))
    print("raw D2H %.0f MB     %8.2f ms  (%.1f GB/s)" % (d2h / 1e6, t, d2h / t / 1e6))
    t, _ = timed(lambda: dev.copy_(host, non_blocking=True))
    print("raw H2D %.0f MB     %8.2f ms  (%.1f GB/s)" % (d2h / 1e6, t, d2h / t / 1e6))
    # the library's own pinned buffer (cudaHostAlloc through the C ABI) as destination
    big = outs["q_chain"]
    tt = torch.empty(big.array.nbytes, dtype=torch.uint8, device="cuda")
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    def raw():
        rt.cudaMemcpy(ctypes.c_void_p(big.array.ctypes.data), ctypes.c_void_p(tt.data_ptr()), ctypes.c_size_t(big.array.nbytes), 2)
    try:
        t, _ = timed(raw)
        print("cudaMemcpy D2H into PinnedBuffer %.0f MB %8.2f ms (%.1f GB/s)" % (big.array.nbytes / 1e6, t, big.array.nbytes / t / 1e6))
    except OSError as e:
        print("libcudart not loadable:", e)


if __name__ == "__main__":
    main()
