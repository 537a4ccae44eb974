"""Host logic of the multi-GPU paths, exercised with world_size 2 on CPU (gloo): agent / parameter-vector
partitions, the moment all-reduce, and independence of the result from the number of ranks."""
import os
import socket

import numpy as np
import pytest

from egdst_b200 import distributed as D


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10, 4096, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [D.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


class _M:  # the two model properties the host logic reads
    nt = 5

    @staticmethod
    def nsimout():
        return 4


def _fake_sim(block, agent0):
    """Stand-in for the CUDA call: a deterministic function of the *global* agent id, like the Philox stream."""
    n = block.shape[0]
    ids = agent0 + np.arange(n)
    sims = np.zeros((n, _M.nt, _M.nsimout()))
    for t in range(_M.nt):
        for j in range(_M.nsimout()):
            sims[:, t, j] = np.sin(ids * 0.1 + t) * (j + 1) + block[:, 1]
    sims[ids % 7 == 3, 3:, :] = np.nan  # some agents die
    mom = np.zeros((3, _M.nsimout(), _M.nt))
    alive = ~np.isnan(sims)
    mom[0] = np.where(alive, sims, 0).sum(axis=0).T
    mom[1] = np.where(alive, sims * sims, 0).sum(axis=0).T
    mom[2] = alive.sum(axis=0).T
    return sims, mom


def _fake_solve_sim(pblock, first):
    out = np.zeros((pblock.shape[0], 3, _M.nsimout(), _M.nt))
    for i in range(pblock.shape[0]):
        out[i] = (first + i + 1) * pblock[i].sum()
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    init = np.column_stack([np.ones(101), rng.random(101)])
    sims, mom, (lo, hi) = D.simulate_sharded(None, _M, None, init, seed=5, simulate_fn=_fake_sim)
    params = rng.random((9, 2))
    table, (plo, phi) = D.solve_batch_sharded(None, _M, params, init, seed=5, solve_sim_fn=_fake_solve_sim)
    q.put((rank, lo, hi, sims, mom, plo, phi, table))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(120)
def test_world2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    init = np.column_stack([np.ones(101), rng.random(101)])
    sims1, mom1 = _fake_sim(init, 0)
    params = rng.random((9, 2))
    table1 = _fake_solve_sim(params, 0)
    # every rank holds the global moments; the local blocks tile the single-process result
    for r in res:
        assert np.allclose(r[4], mom1, rtol=1e-13, atol=1e-12)
        assert np.allclose(r[7], table1, rtol=1e-13, atol=0)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 101
    cat = np.concatenate([res[0][3], res[1][3]], axis=0)
    assert np.array_equal(np.isnan(cat), np.isnan(sims1)) and np.allclose(np.nan_to_num(cat), np.nan_to_num(sims1))
    assert (res[0][5], res[0][6], res[1][5], res[1][6]) == (0, 5, 5, 9)
    mean, var, n = D.moments_to_stats(mom1)
    assert np.allclose(mean, np.nanmean(sims1, axis=0).T) and np.all(n <= 101)
