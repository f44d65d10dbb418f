"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, the single all-gather of
packed top-k lists, the df all-reduce, and that merge(shard results) == unsharded oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from polyfuzz_b200.distributed import Comm, get_comm, shard_bounds
        from polyfuzz_b200 import synth
        from oracle import native, tfidf
        comm = get_comm()
        assert isinstance(comm, Comm) and comm.world_size == world and comm.rank == rank
        to = synth.company_names(600, seed=1); frm = synth.company_names(150, seed=2)
        f, t, _ = tfidf.fit_transform_sklearn(frm, to)
        lo, hi = shard_bounds(len(to), world, rank)
        # df all-reduce: per-shard document frequencies sum to the global ones
        df_local = torch.from_numpy(np.bincount(t[lo:hi].indices, minlength=t.shape[1]).astype(np.int32))
        comm.all_reduce_sum(df_local)
        assert np.array_equal(df_local.numpy(), np.bincount(t.indices, minlength=t.shape[1]))
        # local top-k with global indices, one all-gather, canonical merge
        li, lv = native.spdot_topn(f, t[lo:hi], 5, 0.0, to_index_base=lo)
        gi, gv = comm.all_gather_topk(torch.from_numpy(li), torch.from_numpy(lv))
        assert gi.shape == (world, 150, 5) and gi.dtype == torch.int32 and gv.dtype == torch.float64
        mi, mv = native.topk_merge(gi.numpy(), gv.numpy(), 5)
        fi, fv = native.spdot_topn(f, t, 5, 0.0)
        assert np.array_equal(mi, fi) and np.array_equal(mv, fv)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from polyfuzz_b200.distributed import shard_bounds
    for n in (0, 1, 7, 100, 101):
        for g in (1, 2, 3, 8):
            cover = []
            for r in range(g):
                lo, hi = shard_bounds(n, g, r)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
            assert cover == list(range(n))


def test_pack_unpack():
    from polyfuzz_b200.distributed import pack_topk, unpack_topk
    idx = torch.tensor([[3, -1], [0, 7]], dtype=torch.int32)
    val = torch.tensor([[0.5, 0.0], [1.0, 1e-300]], dtype=torch.float64)
    i2, v2 = unpack_topk(pack_topk(idx, val))
    assert torch.equal(i2, idx) and torch.equal(v2, val)


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
