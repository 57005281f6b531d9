"""Multi-GPU anchor selection on hardware: the merged top-k of the sharded batch must equal the single-GPU
`top_candidates` of the whole batch — what the reference's one `np.argsort(scores)[:k]` returns
(GPyOpt/optimization/anchor_points_generator.py:58-62) — bit for bit, including exact ties across ranks (lowest global index wins).

  * `test_device_merge_of_emulated_ranks...` runs on ONE GPU: the shards are scored one after the other, packed by the library's
    pack kernel, and merged by its merge kernel (`bopl_topk_pack` / `bopl_topk_merge`: the two device stages of
    `bopl_topk_allgather` without the NCCL transport).
  * `test_nccl_allgather_of_real_ranks...` needs >= 2 GPUs (skipped otherwise): one process per GPU, the library's own NCCL
    communicator, `bopl_topk_allgather` (device records, no host sync) and `bopl_topk_allgather_host`.
"""
import ctypes
import os

import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_acquisition, gpu_model

pytestmark = pytest.mark.gpu

K = 20


def _problem_with_cross_shard_ties(N=6000, world=4):
    """Config-S shaped problem (n = 300) whose best candidates are DUPLICATED into other shards: duplicated rows score bit-identically
    (per-candidate determinism), so the global top-k has exact ties across ranks."""
    from bopl_b200.sharding import shard_bounds
    p = make_problem("S", n=300, N=N, n_w=16, L=8)
    gm = gpu_model(p, gradients=False)
    acq = gpu_acquisition(p, gm)
    a = acq._compute_acq(p["Xc"])[:, 0]
    best = np.argsort(-a, kind="stable")[:6]
    X = p["Xc"].copy()
    rs = np.random.RandomState(11)
    for b in best:                                    # a copy of each of the six best rows into two other shards
        for r in rs.choice(world, size=2, replace=False):
            s, e = shard_bounds(N, world, r)
            X[rs.randint(s, e)] = X[b]
    p["Xc"] = X
    return p, gm, acq


def test_device_merge_of_emulated_ranks_equals_whole_batch_topk():
    import torch
    from bopl_b200.sharding import merge_topk, shard_bounds
    world = 4
    p, gm, acq = _problem_with_cross_shard_ties(world=world)
    eng = gm.engine
    X = p["Xc"]
    val_w, idx_w, a = acq.top_candidates(X, K, want_acq=True)
    want = np.argsort(-a, kind="stable")[:K]
    assert np.array_equal(idx_w, want)
    assert len(np.unique(a[want])) < K                 # the ties are really there
    W = 2 + p["d"]
    rec = eng.empty(world * K, W)
    locals_v, locals_i = [], []
    for r in range(world):
        s, e = shard_bounds(len(X), world, r)
        v, i = acq.top_candidates(X[s:e], K, index_offset=s)
        locals_v.append(v)
        locals_i.append(i)
        vd, idd = eng.to_dev(v), torch.from_numpy(i).to(eng.device)
        Xd = eng.to_dev(X[s:e])
        eng._check(eng.lib.bopl_topk_pack(eng.h, vd.data_ptr(), idd.data_ptr(), K, Xd.data_ptr(), s, K, rec[r * K:].data_ptr(), eng._stream()))
    val_all = eng.empty(K)
    idx_all = torch.empty(K, dtype=torch.int64, device=eng.device)
    x_all = eng.empty(K, p["d"])
    eng._check(eng.lib.bopl_topk_merge(eng.h, rec.data_ptr(), world * K, K, val_all.data_ptr(), idx_all.data_ptr(), x_all.data_ptr(),
                                       eng._stream()))
    assert np.array_equal(idx_all.cpu().numpy(), idx_w)
    assert np.array_equal(val_all.cpu().numpy(), val_w)
    assert np.array_equal(x_all.cpu().numpy(), X[idx_w])
    # and the NumPy statement of the merge (what the gloo tests run) agrees
    mv, mi, _ = merge_topk(np.stack(locals_v), np.stack(locals_i), K)
    assert np.array_equal(mi, idx_w) and np.array_equal(mv, val_w)


def test_device_merge_fillers_and_nan():
    """A rank with fewer than k candidates sends NaN / INT64_MAX fillers; NaN scores sort last; fewer real records than k overall."""
    import torch
    p = make_problem("dtlz1a_linear", N=50)
    gm = gpu_model(p, gradients=False)
    eng = gm.engine
    d, W = p["d"], 2 + p["d"]
    Xd = eng.to_dev(p["Xc"])
    rec = eng.empty(3 * 5, W)
    shards = [(np.array([3.0, 1.0, np.nan]), np.array([7, 2, 4])), (np.array([3.0, 2.0]), np.array([12, 11])), (np.empty(0), np.empty(0, dtype=np.int64))]
    for r, (v, i) in enumerate(shards):
        vd = eng.to_dev(v) if len(v) else eng.empty(1)
        idd = torch.from_numpy(np.asarray(i, dtype=np.int64)).to(eng.device) if len(i) else torch.zeros(1, dtype=torch.int64, device=eng.device)
        eng._check(eng.lib.bopl_topk_pack(eng.h, vd.data_ptr(), idd.data_ptr(), len(v), Xd.data_ptr(), 0, 5, rec[r * 5:].data_ptr(), eng._stream()))
    val_all = eng.empty(5)
    idx_all = torch.empty(5, dtype=torch.int64, device=eng.device)
    x_all = eng.empty(5, d)
    eng._check(eng.lib.bopl_topk_merge(eng.h, rec.data_ptr(), 15, 5, val_all.data_ptr(), idx_all.data_ptr(), x_all.data_ptr(), eng._stream()))
    v, i = val_all.cpu().numpy(), idx_all.cpu().numpy()
    assert np.array_equal(i[:4], [7, 12, 11, 2]) and np.array_equal(v[:4], [3.0, 3.0, 2.0, 1.0])
    assert np.isnan(v[4]) and i[4] == 4                 # the real NaN-scored record beats the fillers (lower index)
    assert np.array_equal(x_all.cpu().numpy()[:4], p["Xc"][[7, 12, 11, 2]])


def _rank_main(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from bopl_b200.sharding import ShardedAnchorSelector, TopkGatherer, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p = make_problem("S", n=300, N=6000, n_w=16, L=8)
        p["Xc"] = np.load(os.path.join(out_dir, "X.npy"))
        gm = gpu_model(p, gradients=False, device=rank)
        acq = gpu_acquisition(p, gm)
        eng = gm.engine
        X = p["Xc"]
        s, e = shard_bounds(len(X), world, rank)
        gat = TopkGatherer(eng)
        # device records, no host synchronisation inside the call
        Xd = eng.to_dev(X[s:e])
        a = eng.to_dev(acq._compute_acq(X[s:e])[:, 0])
        v, i = eng.topk_dev(a, K, s)
        va, ia, xa = gat.allgather_dev(v, i, Xd, s, K)
        # host records through the drop-in selector
        sel = ShardedAnchorSelector(lambda Xb, k, off: acq.top_candidates(Xb, k, off), gatherer=gat)
        vh, ih, xh = sel.select_from_full(X, K)
        # a rank with fewer candidates than k: 3 rows per rank, k = 5
        Xs = X[:3 * world]
        vs, is_, xs = ShardedAnchorSelector(lambda Xb, k, off: acq.top_candidates(Xb, k, off), gatherer=gat).select_from_full(Xs, 5)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), va=va.cpu().numpy(), ia=ia.cpu().numpy(), xa=xa.cpu().numpy(), vh=vh, ih=ih,
                 xh=xh, vs=vs, is_=is_, xs=xs)
        gat.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_nccl_allgather_of_real_ranks_equals_whole_batch_topk(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2); the device merge itself is covered by the emulated-ranks test")
    p, gm, acq = _problem_with_cross_shard_ties(world=world)
    X = p["Xc"]
    np.save(os.path.join(str(tmp_path), "X.npy"), X)
    val_w, idx_w, a = acq.top_candidates(X, K, want_acq=True)
    assert np.array_equal(idx_w, np.argsort(-a, kind="stable")[:K])
    vs_w, is_w = acq.top_candidates(X[:3 * world], 5)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for tag in ("a", "h"):
            assert np.array_equal(g["i" + tag], idx_w), (r, tag)
            assert np.array_equal(g["v" + tag], val_w), (r, tag)
            assert np.array_equal(g["x" + tag], X[idx_w]), (r, tag)
        assert np.array_equal(g["is_"], is_w) and np.array_equal(g["vs"], vs_w) and np.array_equal(g["xs"], X[is_w])


def test_collective_entry_points_reject_bad_arguments():
    """Error behaviour of the multi-GPU entry points: status codes + message, no exception across the ABI, nothing launched."""
    import torch
    from bopl_b200._lib import BoplError
    p = make_problem("dtlz1a_linear", N=50)
    eng = gpu_model(p, gradients=False).engine
    v, i = eng.empty(5), torch.zeros(5, dtype=torch.int64, device=eng.device)
    Xd = eng.to_dev(p["Xc"])
    out_v, out_i = eng.empty(5), torch.zeros(5, dtype=torch.int64, device=eng.device)
    with pytest.raises(BoplError) as ei:          # no communicator
        eng._check(eng.lib.bopl_topk_allgather(eng.h, None, v.data_ptr(), i.data_ptr(), 5, Xd.data_ptr(), 0, 5, out_v.data_ptr(), out_i.data_ptr(), None,
                                               eng._stream()))
    assert ei.value.code == -1 and "communicator" in str(ei.value)
    rec = eng.empty(5, 2 + p["d"])
    for k_local, k in ((6, 5), (-1, 5), (1, 0), (1, 300)):
        with pytest.raises(BoplError):
            eng._check(eng.lib.bopl_topk_pack(eng.h, v.data_ptr(), i.data_ptr(), k_local, Xd.data_ptr(), 0, k, rec.data_ptr(), eng._stream()))
    with pytest.raises(BoplError):
        eng._check(eng.lib.bopl_topk_merge(eng.h, None, 5, 5, out_v.data_ptr(), out_i.data_ptr(), None, eng._stream()))
    with pytest.raises(BoplError):                # rank outside the communicator
        comm = ctypes.c_void_p()
        ident = np.zeros(128, dtype=np.uint8)
        eng._check(eng.lib.bopl_nccl_comm_init(eng.h, 2, 2, ctypes.c_void_p(ident.ctypes.data), ctypes.byref(comm)))
