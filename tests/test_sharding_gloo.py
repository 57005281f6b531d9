"""world_size-2 `gloo` test of the N>1 path on CPU: partition -> local top-k -> one all-gather -> identical merge.
The local scorer is the oracle (the CUDA scorer needs a device); what is under test is the host-side sharding logic."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, k, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bopl_b200.sharding import ShardedAnchorSelector
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(N, 3))
    score = np.round(np.sin(7 * X[:, 0]) + X[:, 1], 2)           # rounded -> exact ties across ranks

    def local_top(Xb, kk, off):
        s = np.round(np.sin(7 * Xb[:, 0]) + Xb[:, 1], 2)
        order = np.argsort(-s, kind="stable")[:kk]
        return s[order], order + off

    sel = ShardedAnchorSelector(local_top)
    v, i, x = sel.select_from_full(X, k)
    q.put((rank, v, i, x))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,k", [(1000, 20), (7, 5), (3, 5)])
def test_two_rank_anchor_selection_matches_single_process(N, k):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(N, 3))
    score = np.round(np.sin(7 * X[:, 0]) + X[:, 1], 2)
    want = np.argsort(-score, kind="stable")[:min(k, N)]
    for rank, v, i, x in res:
        assert np.array_equal(i, want), (rank, i, want)
        assert np.array_equal(v, score[want]) and np.array_equal(x, X[want])
