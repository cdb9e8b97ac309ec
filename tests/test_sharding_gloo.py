"""world_size-2 `gloo` test (CPU): the host-side sharding of R, the padded block layout of the
factor replicas and the per-half all-gather reproduce the single-process result bit for bit.
The per-block arithmetic here is the CPU oracle (this is a test; the product path uses the
CUDA library and NCCL with the same layout, csrc/als_abi.cu: exchange())."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gramian(M, n, rank, world, sharded):
    """Full Gramian of the n live rows of replica M; sharded: this rank's block only, summed over
    ranks with one all-reduce of the k x k fp64 partials (csrc/als_abi.cu: launch_gramian)."""
    from oracle import oracle as O
    from myrrix_recommender_b200 import sharding as S
    if not sharded:
        return O.transpose_times_self(M[:n])
    b, e = S.local_block(n, rank, world)
    k = M.shape[1]
    part = O.transpose_times_self(M[b:e]) if e > b else np.zeros((k, k))
    t = torch.from_numpy(np.ascontiguousarray(part))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    G = t.numpy().copy()
    full = O.transpose_times_self(M[:n])
    assert np.abs(G - full).max() <= 1e-12 * max(np.abs(full).max(), 1e-300)
    return G


def _worker(rank, world, port, out_dir, sharded_gramian=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import random_problem
    from oracle import oracle as O
    from myrrix_recommender_b200 import sharding as S
    k, U, I = 6, 101, 37  # deliberately not multiples of world
    ptr, idx, val, Y0 = random_problem(U, I, 7, k, seed=21, neg_fraction=0.1, empty_users=2,
                                       stale_items=1)
    tp, ti, tv = O.csr_transpose(ptr, idx, val, I)
    # this rank's shards of both orientations
    r_ptr, r_idx, r_val = S.shard_rows(ptr, idx, val, rank, world)
    c_ptr, c_idx, c_val = S.shard_rows(tp, ti, tv, rank, world)
    ub, ue = S.local_block(U, rank, world)
    ib, ie = S.local_block(I, rank, world)
    bu, bi = S.block_rows(U, world), S.block_rows(I, world)
    X = np.zeros((S.padded_rows(U, world), k), np.float32)
    Y = np.zeros((S.padded_rows(I, world), k), np.float32)
    Y[:I] = Y0
    for _ in range(3):
        G = _gramian(Y, I, rank, world, sharded_gramian)
        out = X[ub:ue].copy()
        O.als_half(r_ptr, r_idx, r_val, Y[:I], G, out)
        blk = np.zeros((bu, k), np.float32)
        blk[:ue - ub] = out
        parts = [torch.zeros(bu, k) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(blk))
        X = torch.cat(parts).numpy().copy()
        G = _gramian(X, U, rank, world, sharded_gramian)
        out = Y[ib:ie].copy()
        O.als_half(c_ptr, c_idx, c_val, X[:U], G, out)
        blk = np.zeros((bi, k), np.float32)
        blk[:ie - ib] = out
        parts = [torch.zeros(bi, k) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(blk))
        Y = torch.cat(parts).numpy().copy()
    np.save(os.path.join(out_dir, "x%d.npy" % rank), X[:U])
    np.save(os.path.join(out_dir, "y%d.npy" % rank), Y[:I])
    dist.destroy_process_group()


def test_two_rank_sharded_iterations_match_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from conftest import random_problem
    from oracle import oracle as O
    ptr, idx, val, Y0 = random_problem(101, 37, 7, 6, seed=21, neg_fraction=0.1, empty_users=2,
                                       stale_items=1)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, 37, Y0, max_iterations=3, convergence_threshold=1e-12)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("x%d.npy" % r)), Xo)
        assert np.array_equal(np.load(tmp_path / ("y%d.npy" % r)), Yo)


def test_two_rank_sharded_gramian_all_reduce(tmp_path):
    """Each rank reduces the Gramian over its own block, one all-reduce sums the partials: the
    factors agree with the single-process run (the fp64 sums associate differently, so to
    rounding of the fp32 results rather than bit for bit) and every rank holds the same bits."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), True), nprocs=world, join=True)
    from conftest import random_problem
    from oracle import oracle as O
    ptr, idx, val, Y0 = random_problem(101, 37, 7, 6, seed=21, neg_fraction=0.1, empty_users=2,
                                       stale_items=1)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, 37, Y0, max_iterations=3, convergence_threshold=1e-12)
    X0, Y0r = np.load(tmp_path / "x0.npy"), np.load(tmp_path / "y0.npy")
    assert np.array_equal(X0, np.load(tmp_path / "x1.npy")) and np.array_equal(Y0r, np.load(tmp_path / "y1.npy"))
    assert np.abs(X0 - Xo).max() <= 1e-6 * np.abs(Xo).max()
    assert np.abs(Y0r - Yo).max() <= 1e-6 * np.abs(Yo).max()


def test_block_layout_helpers():
    from myrrix_recommender_b200 import sharding as S
    assert S.block_rows(10, 4) == 3 and S.padded_rows(10, 4) == 12
    assert [S.local_block(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert S.local_block(2, 3, 4) == (2, 2)  # ranks beyond the data own nothing
    ptr = np.array([0, 2, 2, 5, 6], np.int64)
    idx = np.arange(6, dtype=np.int32)
    val = np.arange(6, dtype=np.float32)
    p, i, v = S.shard_rows(ptr, idx, val, 1, 2)
    assert list(p) == [0, 3, 4] and list(i) == [2, 3, 4, 5]


def _powerlaw_worker(rank, world, port, out_dir):
    """What als_synth_interactions_powerlaw does on a sharded handle (csrc/als_abi.cu), on CPU: every rank
    draws its own user block from the counter-based generator, sorts its entries by item (stable: users stay
    ascending), sends each item block's run to its owner, and merges what it receives -- ordered by source
    rank, i.e. by user range -- with one more stable sort on the item (build_by_item_distributed)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import synth
    from myrrix_recommender_b200 import sharding as S
    U, I = 1501, 233
    ub, ue = S.local_block(U, rank, world)
    ib, ie = S.local_block(I, rank, world)
    bi = S.block_rows(I, world)
    ptr, idx, val = synth.synth_rows_powerlaw(ub, ue - ub, I, 20, max_nnz=150, seed=5, neg_fraction=0.1)
    users = np.repeat(np.arange(ub, ue, dtype=np.int64), np.diff(ptr))
    order = np.argsort(idx, kind="stable")
    k_sorted, u_sorted, v_sorted = idx[order].astype(np.int64), users[order], val[order]
    gathered = [None] * world
    dist.all_gather_object(gathered, [(k_sorted[(k_sorted // bi) == r], u_sorted[(k_sorted // bi) == r],
                                       v_sorted[(k_sorted // bi) == r]) for r in range(world)])
    keys = np.concatenate([gathered[src][rank][0] for src in range(world)])
    usr = np.concatenate([gathered[src][rank][1] for src in range(world)])
    vals = np.concatenate([gathered[src][rank][2] for src in range(world)])
    order = np.argsort(keys, kind="stable")
    keys, usr, vals = keys[order] - ib, usr[order], vals[order]
    cptr = np.zeros(ie - ib + 1, np.int64)
    np.cumsum(np.bincount(keys, minlength=ie - ib), out=cptr[1:])
    np.savez(os.path.join(out_dir, "pl%d.npz" % rank), ptr=ptr, idx=idx, val=val, cptr=cptr,
             cidx=usr.astype(np.int32), cval=vals)
    dist.destroy_process_group()


def test_two_rank_sharded_powerlaw_generation_and_by_item_exchange(tmp_path):
    world = 2
    mp.spawn(_powerlaw_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O, synth
    from myrrix_recommender_b200 import sharding as S
    U, I = 1501, 233
    ptr, idx, val = synth.synth_rows_powerlaw(0, U, I, 20, max_nnz=150, seed=5, neg_fraction=0.1)
    tp, ti, tv = O.csr_transpose(ptr, idx, val, I)
    for r in range(world):
        z = np.load(tmp_path / ("pl%d.npz" % r))
        ep, ei, ev = S.shard_rows(ptr, idx, val, r, world)
        assert np.array_equal(z["ptr"], ep) and np.array_equal(z["idx"], ei) and np.array_equal(z["val"], ev)
        ep, ei, ev = S.shard_rows(tp, ti, tv, r, world)
        assert np.array_equal(z["cptr"], ep) and np.array_equal(z["cidx"], ei) and np.array_equal(z["cval"], ev)
