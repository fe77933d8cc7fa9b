"""Host-side multi-GPU logic on CPU: the strided sharding of independent chains, exercised with two gloo ranks, and
(on a GPU box) the bit-identity of a sharded run with the single-context run."""
import os
import socket

import numpy as np
import pytest

from hmc_stellar_toy_model_b200 import sharding


def test_shard_ids_partition_every_batch():
    for n in (0, 1, 5, 8, 11000, 8192):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_ids(n, r, world) for r in range(world)]
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
            assert [len(p) for p in parts] == list(sharding.shard_counts(n, world))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= sharding.BLOCK
            for p in parts:  # whole warp groups stay together
                assert all(len(set(p[p // sharding.BLOCK == b])) in (0, min(sharding.BLOCK, n - b * sharding.BLOCK))
                           for b in np.unique(p // sharding.BLOCK))
    with pytest.raises(ValueError):
        sharding.shard_ids(4, 2, 2)


def test_scatter_back_inverts_the_sharding():
    full = np.arange(7 * 3, dtype=float).reshape(7, 3)
    parts = [full[sharding.shard_ids(7, r, 3)] for r in range(3)]
    assert np.array_equal(sharding.scatter_back(parts, 7, 3), full)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def run_local(ids):
            # a stand-in for the resident launch: results are pure functions of the GLOBAL chain id, which is what
            # the device RNG keying (field_ids) guarantees on the GPU
            assert np.array_equal(ids, sharding.shard_ids(n_items, rank, world))
            gid = ids
            return {"q_final": np.stack([gid * 1.5, gid + 0.25, -gid], axis=1).astype(float),
                    "accept_rate": (gid % 7) / 7.0,
                    "A_chain": (gid[:, None] + np.arange(4)[None, :]) % 2 == 0}

        out = sharding.run_sharded(run_local, n_items, rank, world, dist=dist)
        if rank == 0:
            q.put({k: v for k, v in out.items()})
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [9, 16])
def test_two_rank_gloo_gather_restores_global_order(n_items):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gid = np.arange(n_items)
    assert np.array_equal(out["q_final"], np.stack([gid * 1.5, gid + 0.25, -gid], axis=1))
    assert np.array_equal(out["accept_rate"], (gid % 7) / 7.0)
    assert out["A_chain"].dtype == np.bool_
    assert np.array_equal(out["A_chain"], (gid[:, None] + np.arange(4)[None, :]) % 2 == 0)


class _StubResult:
    def __init__(self, ids, rows, S):
        gid = np.asarray(ids, dtype=float)
        self.q_final = np.repeat(gid[:, None], S, axis=1)
        self.accept_rate = gid / 10.0
        self.q_chain = np.repeat(self.q_final[:, None, :], rows, axis=1)
        self.E_chain = np.repeat(gid[:, None], rows, axis=1)
        self.A_chain = (np.repeat(gid[:, None], rows, axis=1) % 2).astype(np.uint8)


class _StubContext:
    """Stands in for RHMCContext in the CPU test of run_chains_sharded: like the library it refuses an empty batch."""

    def __init__(self, n):
        if n < 1:
            raise ValueError("n_fields must be >= 1")
        self.n = n

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def set_data(self, D):
        assert len(D) == self.n

    def run(self, q0, niter, nsteps, dt, seed=0, want=(), field_ids=None, **kw):
        return _StubResult(field_ids, niter + 1, q0.shape[1])


def _worker_empty_shard(rank, world, port, n_items, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D = np.zeros((n_items, 4, 4))
        q0 = np.zeros((n_items, 3))
        out = sharding.run_chains_sharded(_StubContext, D, q0, 6, 10, 0.2, rank, world, dist=dist)
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_rank_with_an_empty_shard_still_joins_the_gather():
    """5 chains on 2 ranks with blocks of 8: rank 1 owns nothing, must not create a context and must not leave rank 0
    waiting in the collective."""
    import torch.multiprocessing as mp

    n_items = 5
    assert sharding.BLOCK * 2 > n_items and len(sharding.shard_ids(n_items, 1, 2)) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_empty_shard, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gid = np.arange(n_items, dtype=float)
    assert out["q_chain"].shape == (n_items, 7, 3) and np.array_equal(out["q_chain"][:, 0, 0], gid)
    assert np.array_equal(out["accept_rate"], gid / 10.0)
    assert out["A_chain"].dtype == np.uint8 and np.array_equal(out["A_chain"][:, 0], gid % 2)


@pytest.mark.gpu
def test_sharded_run_is_bit_identical_to_single_context():
    """Two shards run one after the other on cuda:0 with their global chain ids reproduce the unsharded batch."""
    import stellar_oracle as so
    from helpers import golden, setup_from
    from test_gpu_parity import make_ctx

    g = golden("chain_one_star_m19")
    S = setup_from(g)
    q0 = so.format_q(S, g["q_model"])
    F, niter = 10, 25
    rng = np.random.RandomState(8)
    D = rng.poisson(np.repeat(so.model_image(S, q0)[None], F, axis=0)).astype(float)
    q = np.repeat(q0[None], F, axis=0)
    with make_ctx(S, n_fields=F, max_stars=1) as ctx:
        ctx.set_data(D)
        full = ctx.run(q, niter, 10, 0.2, seed=5, g_ff2=S.g_ff2)
    parts_q, parts_A = [], []
    for r in range(2):
        out = sharding.run_chains_sharded(lambda n: make_ctx(S, n_fields=n, max_stars=1), D, q, niter, 10, 0.2, r, 2,
                                          seed=5, gather=False, g_ff2=S.g_ff2)
        parts_q.append(out["q_chain"])
        parts_A.append(out["A_chain"])
    assert np.array_equal(sharding.scatter_back(parts_q, F, 2), full.q_chain)
    assert np.array_equal(sharding.scatter_back(parts_A, F, 2), full.A_chain)
