"""Sharding of independent chains / fields over the GPUs of one node (BASELINE configs[1] and [3]).

Independent chains need no data-path collective: the batch is cut into blocks of `BLOCK` consecutive chains (the
one-star kernel's warp group, 8) and rank r owns blocks r, r+W, r+2W, ... (`blocks[r::W]`), runs them in ONE resident
launch on its own GPU, and the device RNG is keyed by the *global* chain id (`field_ids`), so the sharded run is
bit-identical to the same batch on one GPU.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only to gather the per-rank chain arrays; the
compute is handed in as a callable so the host logic can be exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np


def _group_size() -> int:
    """Chains per warp group of the one-star kernel (srhmc_chain_group_size); shards keep these groups intact."""
    try:
        from . import _capi

        return int(_capi.load_library().srhmc_chain_group_size())
    except Exception:  # library not built yet: the value only matters for bit-identity on a GPU
        return 8


BLOCK = _group_size()


def shard_ids(n_items: int, rank: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Global ids owned by `rank`: the chains of blocks[rank::world], blocks of `block` consecutive ids."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    ids = np.arange(n_items, dtype=np.int64)
    return ids[(ids // block) % world == rank]


def shard_counts(n_items: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Items per rank under shard_ids."""
    return np.array([len(shard_ids(n_items, r, world, block)) for r in range(world)], dtype=np.int64)


def scatter_back(parts, n_items: int, world: int) -> np.ndarray:
    """Inverse of the strided sharding: parts[r] holds the rows of ids[r::world], in order."""
    first = next(p for p in parts if p is not None and len(p))
    out = np.empty((n_items,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r, p in enumerate(parts):
        ids = shard_ids(n_items, r, world)
        if len(ids):
            out[ids] = p
    return out


def run_sharded(run_local: Callable[[np.ndarray], Dict[str, np.ndarray]], n_items: int, rank: int,
                world: int, gather: bool = True, dist=None) -> Dict[str, np.ndarray] | None:
    """Run `run_local(global_ids)` on this rank's shard and gather the named result arrays
    (leading axis = local chain) onto rank 0 in global-id order.  Returns the gathered dict on rank 0 (and the local
    one when `gather` is False), None on the other ranks."""
    ids = shard_ids(n_items, rank, world)
    local = run_local(ids)
    if not gather or world == 1:
        return local if (world == 1 or not gather) else None
    if dist is None:
        import torch.distributed as dist  # noqa: PLC0415
    import torch  # noqa: PLC0415

    counts = shard_counts(n_items, world)
    result = {} if rank == 0 else None
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    for name in sorted(local):
        arr = np.ascontiguousarray(local[name])
        tail = tuple(arr.shape[1:])
        # pad every shard to the largest so all_gather sees equal shapes
        width = int(counts.max())
        buf = np.zeros((width,) + tail, dtype=arr.dtype)
        buf[: arr.shape[0]] = arr
        view = buf.view(np.uint8) if buf.dtype == np.bool_ else buf
        t = torch.from_numpy(view).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        if rank == 0:
            parts = []
            for r in range(world):
                a = outs[r].cpu().numpy()
                if arr.dtype == np.bool_:
                    a = a.view(np.bool_)
                parts.append(a[: counts[r]])
            result[name] = scatter_back(parts, n_items, world)
    return result


def run_chains_sharded(make_context, D, q0, niter, nsteps, dt, rank, world, seed=0, gather=True, dist=None,
                       want=("q", "E", "A"), **run_kw):
    """Convenience wrapper for BASELINE configs[1]/[3]: `make_context(n_local)` builds an RHMCContext for this rank's
    GPU; `D` [F,R,C] and `q0` [F,S] are the full batch (every rank slices its own shard)."""
    D = np.asarray(D)
    q0 = np.asarray(q0)
    n_items = q0.shape[0]

    S = q0.shape[1]
    stride = int(run_kw.get("chain_stride", 1) or 1)
    rows = (niter + stride) // stride  # ceil((niter + 1) / stride) chain rows kept per field

    def run_local(ids):
        if len(ids) == 0:
            # fewer blocks than ranks: this rank owns nothing.  It still takes part in the gather, so it returns
            # zero-length arrays with the shapes and dtypes the other ranks produce (no context is created: the library
            # rejects n_fields < 1)
            out = {"q_final": np.zeros((0, S)), "accept_rate": np.zeros((0,))}
            for key, name, tail, dt_ in (("q", "q_chain", (rows, S), np.float64), ("p", "p_chain", (rows, S), np.float64),
                                         ("E", "E_chain", (rows,), np.float64), ("V", "V_chain", (rows,), np.float64),
                                         ("T", "T_chain", (rows,), np.float64), ("A", "A_chain", (rows,), np.uint8)):
                if key in want:
                    out[name] = np.zeros((0,) + tail, dtype=dt_)
            return out
        with make_context(len(ids)) as ctx:
            ctx.set_data(D[ids])
            r = ctx.run(q0[ids], niter, nsteps, dt, seed=seed, want=want, field_ids=ids, **run_kw)
            out = {"q_final": r.q_final, "accept_rate": r.accept_rate}
            for key, name in (("q", "q_chain"), ("p", "p_chain"), ("E", "E_chain"), ("V", "V_chain"),
                              ("T", "T_chain"), ("A", "A_chain")):
                if key in want:
                    out[name] = getattr(r, name)
            return out

    return run_sharded(run_local, n_items, rank, world, gather=gather, dist=dist)
